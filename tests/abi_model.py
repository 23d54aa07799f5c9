"""TEST INFRASTRUCTURE (never imported by the product package): a torch-CPU model of the subset of the C ABI (include/ffvc.h)
that the mapper engines call — ffvc_gemm with its operand modes / batch roles / epilogue, LayerNorm, SLN, softmax, the bias
reductions, casts and re-layouts.  tests/test_engine_orchestration_cpu.py swaps it in for `ops.call` / `ops.gemm` so that the
HOST logic of the engines (buffer shapes, strides, offsets, which gradient goes where) is checked against the CPU oracle
without a GPU; the kernels themselves are checked on the GPU (`-m gpu`) against torch / the oracle.  Every function here states
the documented contract of the entry point of the same name; it shares no code with the kernels."""
import torch
import torch.nn.functional as F

BF16, F32 = torch.bfloat16, torch.float32
# shape contracts of the entry points are asserted like the library rejects them; a test that trades legal shapes for speed (tiny
# images) must switch this off explicitly and say so
STRICT_SHAPES = True


def _dgelu(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


def _dsig_mul(x, c):          # d/dx [x * sigmoid(c x)]
    s = torch.sigmoid(c * x)
    return s * (1 + c * x * (1 - s))


# FFVC_ACT_*: none, exact-erf GELU, QuickGELU x*sigmoid(1.702x) (cloob.py:179-181), swish, ReLU — and their derivatives (mul_mode)
_ACT = {0: lambda v: v, 1: F.gelu, 2: lambda v: v * torch.sigmoid(1.702 * v), 3: lambda v: v * torch.sigmoid(v), 4: torch.relu}
_DACT = {1: _dgelu, 2: lambda x: _dsig_mul(x, 1.702), 3: lambda x: _dsig_mul(x, 1.0), 4: lambda x: (x > 0).float()}


def _mat(t, off, rows, cols, row_stride, col_stride):
    return torch.as_strided(t, (rows, cols), (row_stride, col_stride), t.storage_offset() + off)


def gemm_raw(a, b, out, M, N, K, *, a_mode=0, b_mode=0, a_ld=None, b_ld=None, a_role=0, b_role=0, a_bs=0, b_bs=0, batch=1,
             k_segs=1, splits=1, block_n=0, conv=None, pre_out=None, aux=None, res=None, bias=None, ldc=None, out_bs=0,
             atomic=False, bias_mode=1, act=0, mul_mode=0, alpha=1.0, a_off=0, b_off=0, out_off=0, batch_inner=1, a_bs_in=0,
             b_bs_in=0, out_bs_in=0, tile_m=0, two_cta=0, epi_warps=0, argmin_out=None):
    assert argmin_out is None, "the arg-min epilogue is modelled by k_vq_nearest_tc"
    if a_mode == 2:                                         # FFVC_OP_CONV3X3: implicit GEMM over an NHWC tensor, K = 9 * Cin
        n, h, w, c = conv
        assert M == n * h * w and K == 9 * c and batch == 1 and k_segs == 1 and not atomic
        if STRICT_SHAPES:                                   # include/ffvc.h: a 128-pixel tile of whole rows must tile one image
            assert c % 64 == 0, "conv: Cin must be a multiple of 64"
            assert (w % 128 == 0) or (128 % w == 0 and h % (128 // w) == 0), "conv: H*W must tile by the pixel tile (%dx%d)" % (h, w)
        v = alpha * _conv3x3_packed(a, b, n, h, w, c, N)
        _epilogue(v, out, N if ldc is None else ldc, bias if bias_mode == 1 else None, res, aux, mul_mode, act)
        return out
    assert a.dtype == BF16 and b.dtype == BF16
    if a_ld is None:
        a_ld = K if a_mode == 0 else M
    if b_ld is None:
        b_ld = K if b_mode == 0 else N
    ldc = N if ldc is None else ldc
    if STRICT_SHAPES:
        # both operands travel by TMA (cuTensorMapEncodeTiled): 16-byte aligned base addresses and 16-byte multiples for every
        # stride, i.e. multiples of 8 bf16 elements for row pitches, offsets and batch strides
        for what, v in (("a_ld", a_ld), ("b_ld", b_ld), ("a_off", a_off), ("b_off", b_off), ("a_bs", a_bs), ("b_bs", b_bs),
                        ("a_bs_in", a_bs_in), ("b_bs_in", b_bs_in)):
            assert v % 8 == 0, "gemm: %s = %d is not a multiple of 8 elements (TMA 16-byte rule)" % (what, v)
        assert (a.storage_offset() * a.element_size()) % 16 == 0 and (b.storage_offset() * b.element_size()) % 16 == 0, \
            "gemm: operand base not 16-byte aligned"
        assert splits == 1 or (out.dtype == F32 and atomic), "gemm: split-K needs fp32 atomic output"
    bi_n = max(1, batch_inner)
    for bidx in range(batch):
        bo, bi = bidx // bi_n, bidx % bi_n
        acc = torch.zeros(M, N)
        for seg in range(k_segs):                           # FFVC_ROLE_K_SEGMENT: one more contraction over operand dim 2
            ao = a_off + ((bo * a_bs + bi * a_bs_in) if a_role == 1 else (seg * a_bs if a_role == 2 else 0))
            bo_ = b_off + ((bo * b_bs + bi * b_bs_in) if b_role == 1 else (seg * b_bs if b_role == 2 else 0))
            A = _mat(a, ao, M, K, a_ld, 1) if a_mode == 0 else _mat(a, ao, M, K, 1, a_ld)
            Bm = _mat(b, bo_, N, K, b_ld, 1) if b_mode == 0 else _mat(b, bo_, N, K, 1, b_ld)
            acc += A.float() @ Bm.float().t()
        v = alpha * acc
        oo = out_off + bo * out_bs + bi * out_bs_in
        if bias is not None:
            v = v + (bias.float()[None, :N] if bias_mode == 1 else bias.float()[:M, None])
        if pre_out is not None:
            _mat(pre_out, oo, M, N, ldc, 1).copy_(v)
        v = _ACT[act](v)
        if mul_mode:
            v = v * _DACT[mul_mode](_mat(aux, oo, M, N, ldc, 1).float())
        if res is not None:
            v = v + _mat(res, oo, M, N, ldc, 1).float()
        O = _mat(out, oo, M, N, ldc, 1)
        if atomic:
            assert out.dtype == F32
            O.add_(v)
        else:
            O.copy_(v)
    return out


def call(name, *a):
    globals()["k_" + name](*a)


def k_cast_f32_bf16(x, y, n):
    y.view(-1)[:n] = x.reshape(-1)[:n]


def k_cast_bf16_f32(x, y, n):
    y.view(-1)[:n] = x.reshape(-1)[:n].float()


def k_transpose(src, dst, B, R, Cc, in_fp32, out_fp32):
    dst.view(-1)[:B * R * Cc] = src.reshape(-1)[:B * R * Cc].view(B, R, Cc).transpose(1, 2).reshape(-1)


def k_broadcast_rows(x, y, B, n):
    y.view(-1)[:B * n] = x.reshape(-1)[:n].repeat(B)


def k_add_bf16(a, b, y, n):
    y.view(-1)[:n] = (a.reshape(-1)[:n].float() + b.reshape(-1)[:n].float())


def k_layernorm_fwd(x, g, b, y, mean, rstd, R, D, eps):
    assert not STRICT_SHAPES or (D % 8 == 0 and D <= 2048), "layernorm: D must be a multiple of 8 and <= 2048"
    xf = x.reshape(-1)[:R * D].view(R, D).float()
    mu = xf.mean(1)
    var = xf.var(1, unbiased=False)
    rs = (var + eps).rsqrt()
    if mean is not None:
        mean.view(-1)[:R] = mu
        rstd.view(-1)[:R] = rs
    y.view(-1)[:R * D] = (((xf - mu[:, None]) * rs[:, None]) * g.view(-1)[:D] + b.view(-1)[:D]).reshape(-1)


def k_layernorm_bwd(dy, x, g, mean, rstd, add, dx, dg, db, R, D):
    assert not STRICT_SHAPES or (D % 8 == 0 and D <= 2048 and ((dg is None) == (db is None))), "layernorm_bwd: D % 8, D <= 2048, dgamma/dbeta both or none"
    xf = x.reshape(-1)[:R * D].view(R, D).float()
    dyf = dy.reshape(-1)[:R * D].view(R, D).float()
    mu, rs = mean.view(-1)[:R], rstd.view(-1)[:R]
    xh = (xf - mu[:, None]) * rs[:, None]
    if dg is not None:
        dg.view(-1)[:D].add_((dyf * xh).sum(0))
        db.view(-1)[:D].add_(dyf.sum(0))
    gy = dyf * g.view(-1)[:D]
    d = rs[:, None] * (gy - gy.mean(1, keepdim=True) - xh * (gy * xh).mean(1, keepdim=True))
    if add is not None:
        d = d + add.reshape(-1)[:R * D].view(R, D).float()
    dx.view(-1)[:R * D] = d.reshape(-1)


def k_sln_mod_fwd(n, w, gamma, beta, s, total):
    s.view(-1)[:total] = w.reshape(-1)[:total].float() * (gamma.view(-1)[0] * n.reshape(-1)[:total].float() + beta.view(-1)[0])


def k_sln_mod_bwd(ds, n, w, gamma, beta, dn, dw_acc, dgamma, dbeta, total):
    d, nv, wv = ds.reshape(-1)[:total].float(), n.reshape(-1)[:total].float(), w.reshape(-1)[:total].float()
    g, b = gamma.view(-1)[0], beta.view(-1)[0]
    dn.view(-1)[:total] = d * wv * g
    dw_acc.view(-1)[:total] += d * (g * nv + b)
    dgamma.view(-1)[0] += (d * wv * nv).sum()
    dbeta.view(-1)[0] += (d * wv).sum()


def k_softmax_fwd(s, p, rows, n, ld):
    assert ld >= n, "softmax: ld < n"
    S = s.reshape(-1)[:rows * ld].view(rows, ld)[:, :n]
    P = p.view(-1)[:rows * ld].view(rows, ld)
    P.zero_()
    P[:, :n] = torch.softmax(S, -1)


def k_softmax_bwd(p, dp, ds, rows, n, ld, scale):
    P = p.reshape(-1)[:rows * ld].view(rows, ld)[:, :n].float()
    dP = dp.reshape(-1)[:rows * ld].view(rows, ld)[:, :n]
    o = ds.view(-1)[:rows * ld].view(rows, ld)
    o.zero_()
    o[:, :n] = P * (dP - (P * dP).sum(1, keepdim=True)) * scale


def k_colsum(dy, db, rows, n):
    db.view(-1)[:n].add_(dy.reshape(-1)[:rows * n].view(rows, n).float().sum(0))


def k_vitgan_pack_qkv_weight(w, wp, H, dh, dhp, D):
    W = w.reshape(-1)[:3 * H * dh * D].view(dh, 3 * H, D)               # row d*3H + kh
    out = torch.zeros(3 * H, dhp, D)
    out[:, :dh] = W.permute(1, 0, 2)
    wp.view(-1)[:3 * H * dhp * D] = out.reshape(-1)


def k_vitgan_unpack_qkv_wgrad(dwp, dw, H, dh, dhp, D):
    G = dwp.reshape(-1)[:3 * H * dhp * D].view(3 * H, dhp, D)[:, :dh]
    dw.view(-1)[:3 * H * dh * D].add_(G.permute(1, 0, 2).reshape(-1))


def k_vitgan_pack_out_weight(w, wp, H, dh, dhp, D):
    W = w.reshape(-1)[:D * H * dh].view(D, H, dh)
    out = torch.zeros(D, H, dhp)
    out[:, :, :dh] = W
    wp.view(-1)[:D * H * dhp] = out.reshape(-1)


def k_vitgan_unpack_out_wgrad(dwp, dw, H, dh, dhp, D):
    G = dwp.reshape(-1)[:D * H * dhp].view(D, H, dhp)[:, :, :dh]
    dw.view(-1)[:D * H * dh].add_(G.reshape(-1))


def k_softmax_causal_fwd(s, p, rows, T, ld):
    S = s.reshape(-1)[:rows * ld].view(rows // T, T, ld)[:, :, :T].clone()
    mask = torch.ones(T, T, dtype=torch.bool).triu_(1)
    S.masked_fill_(mask, float("-inf"))
    P = p.view(-1)[:rows * ld].view(rows // T, T, ld)
    P.zero_()
    P[:, :, :T] = torch.softmax(S, -1)


def k_cast_f32_bf16_pitched(src, dst, rows, cols, ld):
    o = torch.zeros(rows, ld)
    o[:, :cols] = src.reshape(-1)[:rows * cols].view(rows, cols)
    dst.view(-1)[:rows * ld] = o.reshape(-1)


def k_vitgan_attn_fwd(qkv, out, probs, B, T, H, dh, ld_qkv, ld_out, scale):
    assert not STRICT_SHAPES or (T <= 32 and dh <= 256), "vitgan_attn: T <= 32, dh <= 256 (include/ffvc.h)"
    x = qkv.view(B, T, ld_qkv)[:, :, :3 * H * dh].float().view(B, T, dh, 3, H).permute(3, 0, 4, 1, 2)
    q, k, v = x[0], x[1], x[2]
    P = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    probs.view(B, H, T, T).copy_(P)
    o = (P @ v).permute(0, 2, 1, 3).reshape(B, T, H * dh)
    out.view(B, T, ld_out)[:, :, :H * dh] = o


def k_vitgan_attn_bwd(qkv, probs, dout, dqkv, B, T, H, dh, ld_qkv, ld_out, scale):
    x = qkv.view(B, T, ld_qkv)[:, :, :3 * H * dh].float().view(B, T, dh, 3, H).permute(3, 0, 4, 1, 2)
    q, k, v = x[0], x[1], x[2]
    P = probs.view(B, H, T, T)
    dO = dout.view(B, T, ld_out)[:, :, :H * dh].float().view(B, T, H, dh).permute(0, 2, 1, 3)
    dP = dO @ v.transpose(-1, -2)
    dS = P * (dP - (P * dP).sum(-1, keepdim=True)) * scale
    dq, dk, dv = dS @ k, dS.transpose(-1, -2) @ q, P.transpose(-1, -2) @ dO
    d = torch.stack([dq, dk, dv])                       # k b h t d -> b t (d k h)
    dqkv.view(B, T, ld_qkv)[:, :, :3 * H * dh] = d.permute(1, 3, 4, 0, 2).reshape(B, T, 3 * H * dh)


def k_layernorm_bwd_sums(dy, x, g, mean, rstd, add, dx, dg, db, colsum_out, rowsum_out, rowsum_T, ws, R, D):
    assert not STRICT_SHAPES or rowsum_out is None or R % rowsum_T == 0, "layernorm_bwd_sums: rows must be a multiple of rowsum_T"
    k_layernorm_bwd(dy, x, g, mean, rstd, add, dx, dg, db, R, D)
    d = dx.reshape(-1)[:R * D].view(R, D).float()
    if colsum_out is not None:
        colsum_out.view(-1)[:D].add_(d.sum(0))
    if rowsum_out is not None:
        rowsum_out.view(-1)[:rowsum_T].add_(d.view(R // rowsum_T, rowsum_T, D).sum((0, 2)))


def k_rowsum(dy, db, B, J, D):
    db.view(-1)[:J].add_(dy.reshape(-1)[:B * J * D].view(B, J, D).float().sum((0, 2)))


def k_clip_assemble(pe, cls, pos, x, N, T, W):
    X = x.view(-1)[:N * T * W].view(N, T, W)
    X[:, 0] = (cls.view(-1)[:W] + pos.view(-1)[:W])
    X[:, 1:] = pe.reshape(-1)[:N * (T - 1) * W].view(N, T - 1, W).float() + pos.view(-1)[W:T * W].view(T - 1, W)


def k_copy_rows(src, dst, rows, D, src_stride, dst_stride):
    S = torch.as_strided(src, (rows, D), (src_stride, 1), src.storage_offset())
    torch.as_strided(dst, (rows, D), (dst_stride, 1), dst.storage_offset()).copy_(S)


def _heads(qkv, N, T, H, dh):
    x = qkv.reshape(-1)[:N * T * 3 * H * dh].view(N, T, 3, H, dh).float().permute(2, 0, 3, 1, 4)      # q | k | v, head-major columns
    return x[0], x[1], x[2]


def k_mha_small_fwd(qkv, out, N, T, H, dh, scale):
    assert not STRICT_SHAPES or (T <= 64 and dh == 64), "mha_small: T <= 64, head_dim 64 (include/ffvc.h)"
    q, k, v = _heads(qkv, N, T, H, dh)
    o = torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v
    out.view(-1)[:N * T * H * dh] = o.permute(0, 2, 1, 3).reshape(-1)


def k_mha_small_bwd(qkv, dout, dqkv, N, T, H, dh, scale):
    q, k, v = _heads(qkv, N, T, H, dh)
    dO = dout.reshape(-1)[:N * T * H * dh].view(N, T, H, dh).float().permute(0, 2, 1, 3)
    P = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    dP = dO @ v.transpose(-1, -2)
    dS = P * (dP - (P * dP).sum(-1, keepdim=True)) * scale
    d = torch.stack([dS @ k, dS.transpose(-1, -2) @ q, P.transpose(-1, -2) @ dO])                    # [3][N][H][T][dh]
    dqkv.view(-1)[:N * T * 3 * H * dh] = d.permute(1, 3, 0, 2, 4).reshape(-1)


def k_embed_tokens(tok, emb, pos, x, rows, T, W):
    t = tok.reshape(-1)[:rows]
    x.view(-1)[:rows * W] = (emb.view(-1, W)[t] + pos.view(-1, W)[torch.arange(rows) % T]).reshape(-1)


def k_gather_rows(src, idx, dst, B, T, W):
    S = src.reshape(-1)[:B * T * W].view(B, T, W)
    dst.view(-1)[:B * W] = S[torch.arange(B), idx.reshape(-1)[:B]].reshape(-1)


# ------------------------------------------------------------------------------------------------ VQGAN decoder entry points
def _nhwc(x, n, h, w, c):
    return x.reshape(-1)[:n * h * w * c].view(n, h, w, c).float()


def _conv3x3_packed(x, w, n, h, wd, cin, cout):
    """x NHWC, w packed [cout][tap = ky*3+kx][cin] (include/ffvc.h: 'tap-major') -> [n*h*wd, cout] fp32"""
    W4 = w.reshape(-1)[:cout * 9 * cin].view(cout, 3, 3, cin).float().permute(0, 3, 1, 2)
    y = F.conv2d(_nhwc(x, n, h, wd, cin).permute(0, 3, 1, 2), W4, padding=1)
    return y.permute(0, 2, 3, 1).reshape(n * h * wd, cout)


def _epilogue(v, out, ldc, bias, res, aux, mul_mode, act):
    M, N = v.shape
    if bias is not None:
        v = v + bias.float()[None, :N]
    v = _ACT[act](v)
    if mul_mode:
        v = v * _DACT[mul_mode](_mat(aux, 0, M, N, ldc, 1).float())
    if res is not None:
        v = v + _mat(res, 0, M, N, ldc, 1).float()
    _mat(out, 0, M, N, ldc, 1).copy_(v)


def k_conv3x3_halo(x, w, out, n, h, wd, cin, cout, ldc, bias, res, aux, mul_mode, act, out_fp32):
    assert wd % 128 == 0 and h % 2 == 0 and cin % 64 == 0 and cout <= 128, "documented shape contract of ffvc_conv3x3_halo"
    assert (out.dtype == F32) == bool(out_fp32)
    _epilogue(_conv3x3_packed(x, w, n, h, wd, cin, cout), out, ldc, bias, res, aux, mul_mode, act)


def _group_sums(a, b, n, hw, c, groups=32):
    """per (image, group) sums of a and of a*b  ->  [n][groups][2] doubles"""
    A = a.view(n, hw, groups, c // groups).double()
    Bv = b.view(n, hw, groups, c // groups).double()
    return torch.stack([A.sum((1, 3)), (A * Bv).sum((1, 3))], -1)


def k_conv3x3_halo_gn(x, w, out, n, h, wd, cin, cout, ldc, bias, res, gn_ws):
    assert cout == 128
    k_conv3x3_halo(x, w, out, n, h, wd, cin, cout, ldc, bias, res, None, 0, 0, 0)
    o = _mat(out, 0, n * h * wd, cout, ldc, 1).float()            # statistics of the bf16-rounded output
    gn_ws.view(-1)[:n * 64] = _group_sums(o, o, n, h * wd, cout).reshape(-1)


def k_conv3x3_halo_xf(x, w, out, n, h, wd, cin, cout, ldc, bias, res, xf_mean, xf_rstd, xf_gamma, xf_beta, xf_groups, gn_ws):
    """the conv of swish(GroupNorm(x)) (statistics given), the normalised tensor rounded to bf16 like the kernel's shared-memory tile"""
    a = torch.empty(n * h * wd, cin, dtype=torch.bfloat16)
    k_groupnorm_apply(x, xf_mean, xf_rstd, xf_gamma, xf_beta, a, n, h * wd, cin, xf_groups, 1)
    if gn_ws is not None:
        k_conv3x3_halo_gn(a, w, out, n, h, wd, cin, cout, ldc, bias, res, gn_ws)
    else:
        k_conv3x3_halo(a, w, out, n, h, wd, cin, cout, ldc, bias, res, None, 0, 0, 0)


def k_groupnorm_finalize(ws, mean, rstd, N, HW, C, G, eps):
    s = ws.reshape(-1)[:N * G * 2].view(N, G, 2)
    cnt = HW * (C // G)
    mu = s[..., 0] / cnt
    var = (s[..., 1] / cnt - mu * mu).clamp_min(0)
    mean.view(-1)[:N * G] = mu.reshape(-1).float()
    rstd.view(-1)[:N * G] = (var + eps).rsqrt().reshape(-1).float()


def k_groupnorm_stats(x, ws, mean, rstd, N, HW, C, G, eps):
    _gn_check(C, G)
    xf = x.reshape(-1)[:N * HW * C].view(N, HW, G, C // G).double()
    mu = xf.mean((1, 3))
    var = xf.var((1, 3), unbiased=False)
    ws.view(-1)[:N * G * 2] = torch.stack([xf.sum((1, 3)), (xf * xf).sum((1, 3))], -1).reshape(-1)    # (sum x, sum x^2) per (image, group)
    mean.view(-1)[:N * G] = mu.reshape(-1).float()
    rstd.view(-1)[:N * G] = (var + eps).rsqrt().reshape(-1).float()


def _gn_check(C, G, backward=False):
    if STRICT_SHAPES:
        assert C % 8 == 0 and C % G == 0 and 256 % (C // 8) == 0, "groupnorm: C % 8, C % G, (C / 8) | 256"
        assert not backward or (C // G) in (1, 2, 4) or (C // G) % 8 == 0, "groupnorm_bwd: channels per group 1, 2, 4 or a multiple of 8"


def _gn_parts(x, mean, rstd, gamma, beta, N, HW, C, G):
    xf = x.reshape(-1)[:N * HW * C].view(N, HW, G, C // G).float()
    xh = (xf - mean.view(-1)[:N * G].view(N, 1, G, 1)) * rstd.view(-1)[:N * G].view(N, 1, G, 1)
    u = xh * gamma.view(-1)[:C].view(1, 1, G, C // G) + beta.view(-1)[:C].view(1, 1, G, C // G)
    return xh, u


def k_groupnorm_apply(x, mean, rstd, gamma, beta, y, N, HW, C, G, swish):
    _gn_check(C, G)
    _, u = _gn_parts(x, mean, rstd, gamma, beta, N, HW, C, G)
    y.view(-1)[:N * HW * C] = (u * torch.sigmoid(u) if swish else u).reshape(-1)


def _gn_g(dy, x, mean, rstd, gamma, beta, N, HW, C, G, swish):
    """g = dy * swish'(gamma * xhat + beta) * gamma  (gradient w.r.t. xhat), and xhat"""
    xh, u = _gn_parts(x, mean, rstd, gamma, beta, N, HW, C, G)
    d = dy.reshape(-1)[:N * HW * C].view(N, HW, G, C // G).float()
    if swish:
        d = d * _dsig_mul(u, 1.0)
    return d * gamma.view(-1)[:C].view(1, 1, G, C // G), xh


def _gn_dx(g, xh, s0, s1, rstd, add, dx, N, HW, C, G):
    cnt = HW * (C // G)
    r = rstd.view(-1)[:N * G].view(N, 1, G, 1)
    d = r * (g - (s0 / cnt).view(N, 1, G, 1).float() - xh * (s1 / cnt).view(N, 1, G, 1).float())
    if add is not None:
        d = d + add.reshape(-1)[:N * HW * C].view(N, HW, G, C // G).float()
    dx.view(-1)[:N * HW * C] = d.reshape(-1)


def k_groupnorm_bwd(dy, x, mean, rstd, gamma, beta, ws, add, dx, N, HW, C, G, swish):
    _gn_check(C, G, backward=True)
    g, xh = _gn_g(dy, x, mean, rstd, gamma, beta, N, HW, C, G, swish)
    s0, s1 = g.double().sum((1, 3)), (g * xh).double().sum((1, 3))
    ws.view(-1)[:N * G * 2] = torch.stack([s0, s1], -1).reshape(-1)     # the first pass leaves (sum g, sum g * xhat) per (image, group)
    _gn_dx(g, xh, s0, s1, rstd, add, dx, N, HW, C, G)


def k_conv3x3_halo_gnbwd(x, w, out, n, h, wd, cin, cout, ldc, res, gn_x, gn_mean, gn_rstd, gn_gamma, gn_beta, gn_ws):
    assert cout == 128
    k_conv3x3_halo(x, w, out, n, h, wd, cin, cout, ldc, None, res, None, 0, 0, 0)
    g, xh = _gn_g(_mat(out, 0, n * h * wd, cout, ldc, 1).contiguous(), gn_x, gn_mean, gn_rstd, gn_gamma, gn_beta, n, h * wd, cout, 32, 1)
    gn_ws.view(-1)[:n * 64] = torch.stack([g.double().sum((1, 3)), (g * xh).double().sum((1, 3))], -1).reshape(-1)


def k_groupnorm_bwd_apply(dy, x, mean, rstd, gamma, beta, sums, add, dx, N, HW, C, G, swish):
    _gn_check(C, G, backward=True)
    g, xh = _gn_g(dy, x, mean, rstd, gamma, beta, N, HW, C, G, swish)
    s = sums.reshape(-1)[:N * G * 2].view(N, G, 2)
    _gn_dx(g, xh, s[..., 0], s[..., 1], rstd, add, dx, N, HW, C, G)


def k_upsample2x_fwd(x, y, N, H, W, C):
    assert not STRICT_SHAPES or C % 8 == 0, "upsample: C % 8 != 0"
    X = _nhwc(x, N, H, W, C)
    y.view(-1)[:N * 4 * H * W * C] = X.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(-1)


def k_upsample2x_bwd(dy, dx, N, H, W, C):
    assert not STRICT_SHAPES or C % 8 == 0, "upsample: C % 8 != 0"
    D = dy.reshape(-1)[:N * 4 * H * W * C].view(N, H, 2, W, 2, C).float()
    dx.view(-1)[:N * H * W * C] = D.sum((2, 4)).reshape(-1)


def k_image_post_fwd(d, xr, n):
    xr.view(-1)[:n] = ((d.reshape(-1)[:n] + 1) / 2).clamp(0, 1)


def k_clamp_bwd(g, x, gx, n, lo, hi):
    """ClampWithGrad.backward (main.py:126-129): g * (g * (x - clamp(x)) >= 0)"""
    gg, xx = g.reshape(-1)[:n], x.reshape(-1)[:n]
    gx.view(-1)[:n] = gg * ((gg * (xx - xx.clamp(lo, hi))) >= 0).float()


def k_image_post_bwd(g, d, gd, n):
    t = (d.reshape(-1)[:n] + 1) / 2
    gg = g.reshape(-1)[:n]
    gd.view(-1)[:n] = 0.5 * gg * ((gg * (t - t.clamp(0, 1))) >= 0).float()


def k_im2col3x3_cin3(x, col, N, H, W):
    X = F.pad(_nhwc(x, N, H, W, 3), (0, 0, 1, 1, 1, 1))
    taps = [X[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)]           # col[p][tap*3 + c]
    c = torch.zeros(N, H, W, 32)
    c[..., :27] = torch.stack(taps, 3).reshape(N, H, W, 27)
    col.view(-1)[:N * H * W * 32] = c.reshape(-1)


def k_rownorm2(x, out, rows, C):
    out.view(-1)[:rows] = (x.reshape(-1)[:rows * C].view(rows, C).float() ** 2).sum(1)


def k_vq_nearest(z, codebook, codeT, cnorm, idx, zq_bf16, zq_f32, zc, P, C, ncodes, lo, hi):
    assert not STRICT_SHAPES or codeT is None or (ncodes % 4 == 0 and C in (64, 256)), "vq_nearest: ncodes % 4 == 0, embed dim 64 or 256"
    zz = z.reshape(-1)[:P * C].view(P, C).clamp(lo, hi)
    cb = codebook.reshape(-1)[:ncodes * C].view(ncodes, C)
    d = (zz.double() ** 2).sum(1, keepdim=True) + (cb.double() ** 2).sum(1)[None] - 2 * zz.double() @ cb.double().t()
    i = d.argmin(1)
    idx.view(-1)[:P] = i.to(idx.dtype)
    if zq_bf16 is not None:
        zq_bf16.view(-1)[:P * C] = cb[i].reshape(-1)
    if zq_f32 is not None:
        zq_f32.view(-1)[:P * C] = cb[i].reshape(-1)
    if zc is not None:
        zc.view(-1)[:P * C] = zz.reshape(-1)


def k_vq_prepare_codebook(codebook, csplit, cnorm, ncodes, C):
    cb = codebook.reshape(-1)[:ncodes * C].view(ncodes, C)
    hi_ = cb.to(BF16)
    csplit.view(ncodes, 3 * C).copy_(torch.cat([hi_, (cb - hi_.float()).to(BF16), hi_], 1))
    cnorm.view(-1)[:ncodes] = (cb ** 2).sum(1)


def k_vq_nearest_tc(z, codebook, csplit, cnorm, zsplit, keys, idx, zq_bf16, zq_f32, zc, P, C, ncodes, lo, hi):
    assert not STRICT_SHAPES or (3 * C) % 8 == 0, "vq_nearest_tc: 3*C must be a multiple of 8"
    k_vq_nearest(z, codebook, None, cnorm, idx, zq_bf16, zq_f32, zc, P, C, ncodes, lo, hi)


# ------------------------------------------------------------------------------------------------ cutouts, losses, optimizer
# MakeCutouts (main.py:212-229): the stage arithmetic is the restatement the oracle uses (oracle/cutouts.py: warp convention,
# HSV jitter); what the tests exercise here is the ENGINE's staging — image layout, per-cutout parameter routing, the
# cutout-major order k*B + j, the patch-major output — and its backward chain, obtained from autograd of each stage.
import ctypes as _C

import oracle.cutouts as _oc


def _host3(addr):
    return torch.tensor(list((_C.c_float * 3).from_address(addr)))


def _pool(x_nchw, P):
    return (F.adaptive_avg_pool2d(x_nchw, P) + F.adaptive_max_pool2d(x_nchw, P)) / 2


def k_cutout_pool_fwd(x, y, B, H, W, P):
    y.view(-1)[:B * P * P * 3] = _pool(_nhwc(x, B, H, W, 3).permute(0, 3, 1, 2), P).permute(0, 2, 3, 1).reshape(-1)


_FX = float(2 ** 40)       # fixed-point scale of the cutout backward's scatter buffers (include/ffvc.h)


def _fx_load(t, n):
    v = t.reshape(-1)[:n]
    return v.double().div(_FX).float() if v.dtype == torch.int64 else v.float()


def _fx_store(t, vals):
    n = vals.numel()
    if t.dtype == torch.int64:
        t.view(-1)[:n] = (vals.double().reshape(-1) * _FX).round().to(torch.int64)
    else:
        t.view(-1)[:n] = vals.reshape(-1)


def k_cutout_pool_bwd(x, dy, dx, B, H, W, P, dy_fixed=0):
    assert bool(dy_fixed) == (dy.dtype == torch.int64)
    with torch.enable_grad():
        xi = _nhwc(x, B, H, W, 3).permute(0, 3, 1, 2).clone().requires_grad_(True)
        g, = torch.autograd.grad(_pool(xi, P), xi, _fx_load(dy, B * P * P * 3).view(B, P, P, 3).permute(0, 3, 1, 2))
    dx.view(-1)[:B * H * W * 3] = g.permute(0, 2, 3, 1).reshape(-1)


def _warp_stage(src_nchw, hinv, N, n_src, border):
    return _oc.warp(src_nchw.repeat(N // n_src, 1, 1, 1), hinv.reshape(-1)[:N * 9].view(N, 3, 3), "border" if border else "zeros")


def k_cutout_warp_fwd(inp, hinv, out, N, n_src, P, border):
    o = _warp_stage(_nhwc(inp, n_src, P, P, 3).permute(0, 3, 1, 2), hinv, N, n_src, border)
    out.view(-1)[:N * P * P * 3] = o.permute(0, 2, 3, 1).reshape(-1)


def k_cutout_warp_bwd(dout, hinv, din, N, n_src, P, border):
    with torch.enable_grad():
        xi = torch.zeros(n_src, 3, P, P, requires_grad=True)
        g, = torch.autograd.grad(_warp_stage(xi, hinv, N, n_src, border), xi,
                                 _fx_load(dout, N * P * P * 3).view(N, P, P, 3).permute(0, 3, 1, 2))
    assert din.dtype == torch.int64 and dout.dtype == torch.int64, "fixed-point scatter buffers (include/ffvc.h)"
    _fx_store(din, g.permute(0, 2, 3, 1))


def _final_stage(c_nchw, hinv, sat, hue, erase, N, P):
    b = _oc.warp(c_nchw, hinv.reshape(-1)[:N * 9].view(N, 3, 3), "zeros")
    b = _oc.color_jitter(b, sat.reshape(-1)[:N], hue.reshape(-1)[:N])
    x0, y0, x1, y1 = [int(v) for v in erase.reshape(-1)[:4]]
    if x1 > x0 and y1 > y0:
        mask = torch.ones_like(b)
        mask[:, :, y0:y1, x0:x1] = 0
        b = b * mask
    return b


def _to_patches(img_nchw, N, P, patch):
    g = P // patch
    return img_nchw.reshape(N, 3, g, patch, g, patch).permute(0, 2, 4, 1, 3, 5).reshape(N, g * g, 3 * patch * patch)


def k_cutout_final_fwd(cut1, hinv, sat, hue, noise, facs, erase, mean, std, patches, img_out, N, P, patch):
    assert P % patch == 0, "cutout_final: cut size must be a multiple of the patch size"
    b = _final_stage(_nhwc(cut1, N, P, P, 3).permute(0, 3, 1, 2), hinv, sat, hue, erase, N, P)
    b = b + facs.reshape(-1)[:N].view(N, 1, 1, 1) * noise.reshape(-1)[:N * 3 * P * P].view(N, 3, P, P)       # main.py:223-225
    b = (b - _host3(mean).view(1, 3, 1, 1)) / _host3(std).view(1, 3, 1, 1)                                   # main.py:797
    patches.view(-1)[:N * 3 * P * P] = _to_patches(b, N, P, patch).reshape(-1)
    if img_out is not None:
        img_out.view(-1)[:N * 3 * P * P] = b.reshape(-1)


def k_cutout_final_bwd(cut1, hinv, sat, hue, erase, mean, std, dpatches, dcut1, N, P, patch):
    g = P // patch
    dp = dpatches.reshape(-1)[:N * 3 * P * P].view(N, g, g, 3, patch, patch).float().permute(0, 3, 1, 4, 2, 5).reshape(N, 3, P, P)
    dp = dp / _host3(std).view(1, 3, 1, 1)
    with torch.enable_grad():
        ci = _nhwc(cut1, N, P, P, 3).permute(0, 3, 1, 2).clone().requires_grad_(True)
        gr, = torch.autograd.grad(_final_stage(ci, hinv, sat, hue, erase, N, P), ci, dp)
    assert dcut1.dtype == torch.int64, "fixed-point scatter buffer (include/ffvc.h)"
    _fx_store(dcut1, gr.permute(0, 2, 3, 1))


def k_spherical_loss(embed, target, loss_out, dembed, dembed_bf16, N, B, D, coef):
    """main.py:801-811: mean(2 * asin(|normalize(H) - normalize(e)| / 2)^2) * coef with H = out_feats.repeat(cutn, 1)"""
    with torch.enable_grad():
        e = embed.reshape(-1)[:N * D].view(N, D).clone().requires_grad_(True)
        Hh = F.normalize(target.reshape(-1)[:B * D].view(B, D).repeat(N // B, 1), dim=-1)
        loss = ((F.normalize(e, dim=-1) - Hh).norm(dim=-1).div(2).arcsin().pow(2).mul(2)).mean() * coef
        g, = torch.autograd.grad(loss, e)
    loss_out.view(-1)[0] = loss.detach()
    if dembed is not None:
        dembed.view(-1)[:N * D] = g.reshape(-1)
    if dembed_bf16 is not None:
        dembed_bf16.view(-1)[:N * D] = g.reshape(-1)


def _flash_ref(qkv, N, T, heads, scale, causal):
    W = heads * 64
    x = qkv.reshape(-1)[:N * T * 3 * W].view(N, T, 3, heads, 64).float()
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)      # (N, heads, T, 64)
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool).triu_(1), float("-inf"))
    return q, k, v, s


def k_mha_flash_fwd(qkv, out, lse, N, T, heads, head_dim, scale, causal):
    """softmax(q k^T * scale [causal]) v per (sequence, head); lse = log2-domain log-sum-exp of the masked scaled scores"""
    assert head_dim == 64
    q, k, v, s = _flash_ref(qkv, N, T, heads, scale, causal)
    p = torch.softmax(s, dim=-1).to(torch.bfloat16).float()          # the kernel rounds P to bf16 in front of P V
    o = (p @ v).transpose(1, 2).reshape(N * T * heads * 64)
    out.view(-1)[:o.numel()] = o
    lse.view(-1)[:N * heads * T] = (torch.logsumexp(s, dim=-1) * 1.4426950408889634).reshape(-1)


def k_mha_flash_bwd(qkv, out, dout, lse, delta_ws, dqkv, N, T, heads, head_dim, scale, causal):
    W = heads * 64
    with torch.enable_grad():
        x = qkv.reshape(-1)[:N * T * 3 * W].view(N, T, 3 * W).float().clone().requires_grad_(True)
        q, k, v, s = _flash_ref(x, N, T, heads, scale, causal)
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(N, T, W)
        g, = torch.autograd.grad(o, x, dout.reshape(-1)[:N * T * W].view(N, T, W).float())
    dqkv.view(-1)[:N * T * 3 * W] = g.reshape(-1)


def k_conv_taps_gather(v, bias, y, xr, N, H, W, COUT):
    """y[p][co] = bias[co] + sum_tap v[p + off(tap)][tap * COUT + co] (zero outside the image); xr = clamp((y + 1) / 2, 0, 1)"""
    V = v.reshape(-1)[:N * H * W * 32].view(N, H, W, 32).float()
    out = torch.zeros(N, H, W, COUT)
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        ys0, ys1, xs0, xs1 = max(0, -dy), min(H, H - dy), max(0, -dx), min(W, W - dx)
        out[:, ys0:ys1, xs0:xs1] += V[:, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx, tap * COUT:(tap + 1) * COUT]
    if bias is not None:
        out = out + bias.reshape(-1)[:COUT].float()
    y.view(-1)[:N * H * W * COUT] = out.reshape(-1)
    if xr is not None:
        xr.view(-1)[:N * H * W * COUT] = ((out + 1) / 2).clamp(0, 1).reshape(-1)


def k_spherical_loss2(embed, target, target2, loss_out, dembed, dembed_bf16, N, B, D, coef, coef2):
    """main.py:801-824: the target term plus `input_loss_coef` times the same distance to the source embeddings"""
    with torch.enable_grad():
        e = embed.reshape(-1)[:N * D].view(N, D).clone().requires_grad_(True)
        en = F.normalize(e, dim=-1)
        loss = 0
        for t, c in ((target, coef), (target2, coef2)):
            Hh = F.normalize(t.reshape(-1)[:B * D].view(B, D).repeat(N // B, 1), dim=-1)
            loss = loss + ((en - Hh).norm(dim=-1).div(2).arcsin().pow(2).mul(2)).mean() * c
        g, = torch.autograd.grad(loss, e)
    loss_out.view(-1)[0] = loss.detach()
    if dembed is not None:
        dembed.view(-1)[:N * D] = g.reshape(-1)
    if dembed_bf16 is not None:
        dembed_bf16.view(-1)[:N * D] = g.reshape(-1)


def k_normalize_rows(x, y, rows, D):
    y.view(-1)[:rows * D] = F.normalize(x.reshape(-1)[:rows * D].view(rows, D), dim=1).reshape(-1)


def k_tv_loss(img, loss_accum, dimg_accum, B, H, W, C, coef):
    """main.py:423-428 on NHWC; ACCUMULATES into loss_accum / dimg_accum"""
    with torch.enable_grad():
        x = _nhwc(img, B, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
        tv = 0.5 * ((x[:, :, 1:, :] - x[:, :, :-1, :]).abs().mean() + (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().mean()) * coef
        g, = torch.autograd.grad(tv, x)
    loss_accum.view(-1)[0] += tv.detach()
    dimg_accum.view(-1)[:B * H * W * C] += g.permute(0, 2, 3, 1).reshape(-1)


def k_sumsq(x, out, n):
    out.view(-1)[0] = (x.reshape(-1)[:n].double() ** 2).sum().float()


def k_axpy_f32(x, y, a, n):
    y.view(-1)[:n] += a * x.reshape(-1)[:n]


def k_adam_tick(h):
    """include/ffvc.h: the 16-float scalar block; cosine lr of update t uses scheduler epoch t - 1 (main.py:835-837)"""
    import math
    t = float(h[8]) + 1.0
    h[8] = t
    h[4] = 1.0 - float(h[1]) ** t
    h[5] = math.sqrt(1.0 - float(h[2]) ** t)
    if float(h[13]) > 0:
        h[0] = float(h[14]) + (float(h[12]) - float(h[14])) * 0.5 * (1.0 + math.cos(math.pi * (t - 1.0) / float(h[13])))
    if float(h[9]) > 0:
        h[11] = min(1.0, float(h[9]) / (math.sqrt(float(h[10])) * float(h[6]) + 1e-6))


def _adam(p, g, m, v, shadow, ema, n, h):
    gs = float(h[6]) * (float(h[11]) if float(h[9]) > 0 else 1.0)
    gv = g.reshape(-1)[:n] * gs
    if float(h[7]) != 0:
        gv = gv + float(h[7]) * p.view(-1)[:n]
    b1, b2 = float(h[1]), float(h[2])
    m.view(-1)[:n] = b1 * m.view(-1)[:n] + (1 - b1) * gv
    v.view(-1)[:n] = b2 * v.view(-1)[:n] + (1 - b2) * gv * gv
    p.view(-1)[:n] -= (float(h[0]) / float(h[4])) * m.view(-1)[:n] / (v.view(-1)[:n].sqrt() / float(h[5]) + float(h[3]))
    if shadow is not None:
        shadow.view(-1)[:n] = p.view(-1)[:n]
    if ema is not None:                                   # torch_ema: decay = min(decay, (1 + t) / (10 + t))
        t = float(h[8])
        omd = 1.0 - min(float(h[15]), (1.0 + t) / (10.0 + t))
        ema.view(-1)[:n] -= omd * (ema.view(-1)[:n] - p.view(-1)[:n])


def k_adam_step(p, g, m, v, shadow, n, h):
    _adam(p, g, m, v, shadow, None, n, h)


def k_adam_step_ema(p, g, m, v, shadow, ema, n, h):
    _adam(p, g, m, v, shadow, ema, n, h)


# ------------------------------------------------------------------------------------------------ LPIPS-VGG16 diversity term
def k_normalize3_fwd(x, y, n, mean, std):
    y.view(-1)[:n] = ((x.reshape(-1)[:n].view(-1, 3) - _host3(mean)) / _host3(std)).reshape(-1)


def k_normalize3_bwd(dy, dx_accum, n, std):
    dx_accum.view(-1)[:n] += (dy.reshape(-1)[:n].view(-1, 3) / _host3(std)).reshape(-1)


def k_maxpool2x2_fwd(x, y, N, H, W, C):
    assert not STRICT_SHAPES or (C % 8 == 0 and H % 2 == 0 and W % 2 == 0), "maxpool2x2: C % 8, H % 2, W % 2 must be 0"
    y.view(-1)[:N * (H // 2) * (W // 2) * C] = F.max_pool2d(_nhwc(x, N, H, W, C).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1)


def k_maxpool2x2_bwd(x, dy, dx, N, H, W, C):
    with torch.enable_grad():
        xi = _nhwc(x, N, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
        g, = torch.autograd.grad(F.max_pool2d(xi, 2, 2), xi, _nhwc(dy, N, H // 2, W // 2, C).permute(0, 3, 1, 2))
    dx.view(-1)[:N * H * W * C] = g.permute(0, 2, 3, 1).reshape(-1)


def k_relu_mask(g, post, out, n):
    out.view(-1)[:n] = g.reshape(-1)[:n].float() * (post.reshape(-1)[:n].float() > 0)


def k_diversity_tap(feats, loss_accum, dfeat, R, B, HW, C, scale):
    assert not STRICT_SHAPES or C <= 512, "diversity_tap: C <= 512"
    """one VGG tap of main.py:778-782: normalize_tensor over channels, squared differences between the R samples of each prompt"""
    with torch.enable_grad():
        f = feats.reshape(-1)[:R * B * HW * C].view(R, B, HW, C).float().clone().requires_grad_(True)
        fn = f / (f.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
        div = ((fn.view(R, 1, B, HW, C) - fn.view(1, R, B, HW, C)) ** 2).sum(-1).mean() * scale
        g, = torch.autograd.grad(div, f)
    loss_accum.view(-1)[0] += div.detach()
    dfeat.view(-1)[:R * B * HW * C] = g.reshape(-1)


# ------------------------------------------------------------------------------------------------ alternates of the above
def k_groupnorm_fused_fwd(x, gamma, beta, y, mean, rstd, ws, N, HW, C, G, swish, eps):
    """single-kernel form: the same arithmetic as ffvc_groupnorm_stats + ffvc_groupnorm_apply"""
    assert not STRICT_SHAPES or (C % 8 == 0 and C % G == 0 and 512 % (C // 8) == 0), "groupnorm_fused: C % 8, C % G, (C / 8) | 512"
    xf = x.reshape(-1)[:N * HW * C].view(N, HW, G, C // G).double()
    mean.view(-1)[:N * G] = xf.mean((1, 3)).reshape(-1).float()
    rstd.view(-1)[:N * G] = (xf.var((1, 3), unbiased=False) + eps).rsqrt().reshape(-1).float()
    _, u = _gn_parts(x, mean, rstd, gamma, beta, N, HW, C, G)
    y.view(-1)[:N * HW * C] = (u * torch.sigmoid(u) if swish else u).reshape(-1)


def k_groupnorm_fused_bwd(dy, x, mean, rstd, gamma, beta, ws, add, dx, N, HW, C, G, swish):
    assert not STRICT_SHAPES or (C % 8 == 0 and C % G == 0 and 512 % (C // 8) == 0), "groupnorm_fused: C % 8, C % G, (C / 8) | 512"
    g, xh = _gn_g(dy, x, mean, rstd, gamma, beta, N, HW, C, G, swish)
    _gn_dx(g, xh, g.double().sum((1, 3)), (g * xh).double().sum((1, 3)), rstd, add, dx, N, HW, C, G)


def k_conv3x3_cin3(x, w, y, N, H, W, COUT):
    """3x3 conv with 3 input channels: fp32 NHWC in, fp32 weights [COUT][9][3] (tap-major), bf16 NHWC out"""
    assert not STRICT_SHAPES or (COUT % 8 == 0 and COUT <= 512), "conv3x3_cin3: COUT % 8 != 0 or too large"
    W4 = w.reshape(-1)[:COUT * 27].view(COUT, 3, 3, 3).float().permute(0, 3, 1, 2)
    o = F.conv2d(_nhwc(x, N, H, W, 3).permute(0, 3, 1, 2), W4, padding=1)
    y.view(-1)[:N * H * W * COUT] = o.permute(0, 2, 3, 1).reshape(-1)
