"""TEST INFRASTRUCTURE (never imported by the product package): a torch-CPU model of the subset of the C ABI (include/ffvc.h)
that the mapper engines call — ffvc_gemm with its operand modes / batch roles / epilogue, LayerNorm, SLN, softmax, the bias
reductions, casts and re-layouts.  tests/test_engine_orchestration_cpu.py swaps it in for `ops.call` / `ops.gemm` so that the
HOST logic of the engines (buffer shapes, strides, offsets, which gradient goes where) is checked against the CPU oracle
without a GPU; the kernels themselves are checked on the GPU (`-m gpu`) against torch / the oracle.  Every function here states
the documented contract of the entry point of the same name; it shares no code with the kernels."""
import torch
import torch.nn.functional as F

BF16, F32 = torch.bfloat16, torch.float32


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
    assert conv is None and argmin_out is None, "the model covers the dense forms only"
    assert a.dtype == BF16 and b.dtype == BF16
    if a_ld is None:
        a_ld = K if a_mode == 0 else M
    if b_ld is None:
        b_ld = K if b_mode == 0 else N
    ldc = N if ldc is None else ldc
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
    xf = x.reshape(-1)[:R * D].view(R, D).float()
    mu = xf.mean(1)
    var = xf.var(1, unbiased=False)
    rs = (var + eps).rsqrt()
    if mean is not None:
        mean.view(-1)[:R] = mu
        rstd.view(-1)[:R] = rs
    y.view(-1)[:R * D] = (((xf - mu[:, None]) * rs[:, None]) * g.view(-1)[:D] + b.view(-1)[:D]).reshape(-1)


def k_layernorm_bwd(dy, x, g, mean, rstd, add, dx, dg, db, R, D):
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
