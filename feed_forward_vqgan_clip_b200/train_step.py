"""One B200-native train step of the feed-forward VQGAN-CLIP pipeline — the body of the reference's hot loop
(main.py:729-837) with every device computation in libffvc_sm100.so:

    z = net(inp_feats)                       main.py:754          MixerEngine.forward
    z = clamp_with_grad(z, z_min, z_max)     main.py:763          ffvc_vq_nearest (fused clamp)
    xr = synth(vq, z)                        main.py:767,140-143  ffvc_vq_nearest + DecoderEngine.forward
    x = (make_cutouts(xr) - mean) / std      main.py:796-797      CutoutEngine.forward
    embed = perceptor.encode_image(x)        main.py:799          ClipEngine.forward
    dists = spherical distance loss          main.py:801-811      ffvc_spherical_loss (fwd + bwd fused)
    loss.backward()                          main.py:832          the engines' explicit backward passes
    opt.step()                               main.py:835          ffvc_adam_step (fused, flat arena)
    Horovod gradient averaging               main.py:627          one NCCL all-reduce of the flat gradient arena

The step can be captured into a CUDA graph (static shapes): `TrainStep.capture()`; inputs then travel through
static device buffers and one graph launch replaces ~2000 kernel launches.
"""
import math

import torch

from . import ops
from .ops import BF16, F32, call
from .cutouts import CutoutEngine, sample_params


class FusedAdam:
    """The reference's optimizer block (main.py:591,693,702-709,825-837,843-844) on the mapper's flat arena, one launch:
    torch.optim.Adam (lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0), optional `clip_grad_norm_(max_norm)`, optional
    cosine annealing (CosineAnnealingLR(T_max, eta_min=0)), optional torch_ema ExponentialMovingAverage(decay).
    Every per-step scalar lives in a device block (`hyper`, layout in include/ffvc.h) so a captured CUDA graph sees it."""

    def __init__(self, engine, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.eng = engine
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(engine.arena)
        self.v = torch.zeros_like(engine.arena)
        self.t = 0
        self.clip, self.ema = 0.0, None
        b1, b2 = betas
        self.hyper = torch.tensor([lr, b1, b2, eps, 1.0, 1.0, 1.0, weight_decay, 0.0, 0.0, 0.0, 1.0, lr, 0.0, 0.0, 0.0],
                                  dtype=F32).to(engine.dev)

    def set_lr(self, lr):
        self.lr = lr
        self.hyper[0:1].fill_(lr)
        self.hyper[12:13].fill_(lr)

    def set_grad_scale(self, s):
        self.hyper[6:7].fill_(s)

    def set_clip_grad_norm(self, max_norm):
        """main.py:693,833-834.  The norm is taken over the (rank-averaged) flat gradient arena."""
        self.clip = float(max_norm or 0.0)
        self.hyper[9:10].fill_(self.clip)

    def set_cosine(self, t_max, eta_min=0.0):
        """main.py:702-705: CosineAnnealingLR(opt, T_max=steps, eta_min=0); t_max = 0 switches the schedule off."""
        self.hyper[13:14].fill_(float(t_max))
        self.hyper[14:15].fill_(float(eta_min))

    def enable_ema(self, decay=0.995):
        """main.py:524-525,615: ExponentialMovingAverage(net.parameters(), decay) — the shadow copy is one more flat arena,
        updated inside the Adam launch."""
        if self.ema is None:
            self.ema = self.eng.arena.clone()
        self.hyper[15:16].fill_(float(decay))

    def tick(self):
        """advance the device-side step counter: bias corrections, lr of this step, clip coefficient (needs the whole gradient)"""
        e = self.eng
        self.t += 1
        if self.clip > 0:
            call("sumsq", e.grad, self.hyper[10:11], e.total)
        call("adam_tick", self.hyper)

    def step_range(self, lo, hi):
        """the update of arena elements [lo, hi) (8-element aligned slots): parameters, moments, bf16 shadow (+ EMA)"""
        e = self.eng
        if hi <= lo:
            return
        shadow = e.shadow[lo:hi] if e.shadow is not None else None
        if self.ema is not None:
            call("adam_step_ema", e.arena[lo:hi], e.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], shadow, self.ema[lo:hi],
                 hi - lo, self.hyper)
        else:
            call("adam_step", e.arena[lo:hi], e.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], shadow, hi - lo, self.hyper)
        e.ext_shadow_fresh = True                      # the kernel refreshed the bf16 shadow ...
        e._adam_ver = tuple(p._version for p in e.params)   # ... of the parameters as they are NOW (a later in-place write voids it)

    def apply(self):
        """tick the device-side step counter (bias corrections, lr, clip coefficient) and update parameters + bf16 shadow
        (+ EMA) in one pass."""
        self.tick()
        self.step_range(0, self.eng.total)

    # ---- checkpoint I/O in torch.optim.Adam's own format (main.py:593-596 loads it, :911 saves it as opt.th)
    def _slices(self):
        e = self.eng
        for i, p in enumerate(e.params):
            off = p.data_ptr() - e.arena.data_ptr()
            assert off % 4 == 0
            yield i, off // 4, p.numel(), p.shape

    def state_dict(self):
        step = int(self.hyper[8].item())
        state = {}
        if step > 0:
            for i, off, n, shape in self._slices():
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": self.m[off:off + n].view(shape).clone(),
                            "exp_avg_sq": self.v[off:off + n].view(shape).clone()}
        group = {"lr": float(self.hyper[0].item()), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self.eng.params)))}
        if float(self.hyper[13].item()) > 0:
            group["initial_lr"] = float(self.hyper[12].item())
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        state = sd["state"]
        step = 0
        for i, off, n, shape in self._slices():
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            self.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(st["step"]))
        self.t = step
        self.hyper[8:9].fill_(float(step))
        g = sd["param_groups"][0]
        self.hyper[0:1].fill_(float(g["lr"]))
        self.hyper[12:13].fill_(float(g.get("initial_lr", g["lr"])))

    # ---- EMA access (torch_ema's copy_to / average_parameters, main.py:905-910)
    def ema_state_dict(self):
        """state_dict of the mapper with the EMA weights (what the reference writes to checkpoint_ema.th)."""
        assert self.ema is not None, "EMA is not enabled"
        e = self.eng
        sd = {k: v.clone() for k, v in e.m.state_dict().items()}
        for (name, _), (i, off, n, shape) in zip(e.m.named_parameters(), self._slices()):
            sd[name] = self.ema[off:off + n].view(shape).clone()
        return sd

    def average_parameters(self):
        """context manager: parameters temporarily replaced by their EMA (torch_ema.average_parameters)."""
        import contextlib

        @contextlib.contextmanager
        def _ctx():
            e = self.eng
            keep = e.arena.clone()
            e.arena.copy_(self.ema)
            e.ext_shadow_fresh = False
            e._shadow_version = None
            try:
                yield
            finally:
                e.arena.copy_(keep)
                e.ext_shadow_fresh = False
                e._shadow_version = None
        return _ctx()


class TrainStep:
    def __init__(self, net, vq, perceptor, cutn=8, lr=1e-3, cut_size=224, target_loss_coef=1.0, world_size=1,
                 process_group=None, seed=0, l2_coef=0.0, tv_coef=0.0, diversity_coef=0.0, repeat=1, lpips_net=None,
                 clip_grad_norm=None, scheduler=None, total_steps=0, use_ema=False, ema_decay=0.995,
                 diversity_mode="between_same_prompts", input_loss=False, input_loss_coef=1.0, normalize_input=False,
                 noise_dim=0, nb_noise=None):
        self.net, self.vq, self.perceptor = net, vq, perceptor
        self.mix = net.engine()
        self.dec = vq.engine()
        self.clip = perceptor.visual.engine()
        self.dev = self.mix.dev
        self.cutn, self.cut_size = cutn, cut_size
        self.cut = CutoutEngine(cut_size, cutn, self.clip.patch, self.dev)
        self.coef = target_loss_coef
        self.l2_coef, self.tv_coef = float(l2_coef), float(tv_coef)          # main.py:690-691,758-773,831
        self.aux_loss = torch.zeros(3, device=self.dev, dtype=F32)           # [l2, tv, -diversity] (already scaled by their coefficients)
        self.diversity_coef, self.repeat = float(diversity_coef), int(repeat)  # main.py:532-537,739-740,776-791
        self.div = lpips_net.engine() if (lpips_net is not None and self.diversity_coef != 0) else None
        if diversity_mode not in ("between_same_prompts", "all"):       # main.py:695,788-789
            raise ValueError("diversity_mode should be 'between_same_prompts' lr 'all'")
        self.diversity_mode = diversity_mode
        # main.py:690-691,812-824: + input_loss_coef * the same spherical distance to the SOURCE embeddings
        self.input_loss_coef = float(input_loss_coef) if input_loss else 0.0
        self.normalize_input = bool(normalize_input)                      # main.py:696,734-735
        # main.py:457,649,680-684,741-750: a noise vector concatenated to the mapper's input — fresh N(0,1) per row, or one of
        # `nb_noise` fixed vectors per repeat group
        self.noise_dim, self.nb_noise = int(noise_dim or 0), int(nb_noise or 0)
        self.opt = FusedAdam(self.mix, lr=lr)
        self.world, self.pg = world_size, process_group
        self.opt.set_grad_scale(1.0 / world_size)
        if clip_grad_norm:                                                 # main.py:693,833-834
            self.opt.set_clip_grad_norm(clip_grad_norm)
        if scheduler is not None:                                          # main.py:702-709
            if scheduler != "cosine":
                raise ValueError(scheduler)
            if total_steps <= 0:
                raise ValueError("the cosine schedule needs total_steps (T_max = number of training steps, main.py:704-705)")
            self.opt.set_cosine(total_steps)
        if use_ema:                                                        # main.py:510,524-525,843-844
            self.opt.enable_ema(ema_decay)
        # data parallel: the flat gradient arena is all-reduced in buckets of `bucket_layers` mixer layers on a side stream
        # WHILE backward is still producing the earlier layers' gradients (the reference gets this from Horovod's hooks,
        # main.py:627); engines without per-layer completion callbacks fall back to one all-reduce after backward
        self.bucket_layers = 8
        # opt-in (not yet measured on a GPU): start Adam on the slices whose all-reduce has already been issued while the last
        # bucket (head of the arena + late ranges) is still in flight; needs the clip coefficient off (it wants the whole norm)
        self.tail_overlap = False
        # SMs left to NCCL while gradient buckets are in flight: the persistent tcgen05 grids are sized to (SMs - comm_sms) from the
        # first bucket to the end of backward (option "sm_limit"), so that NCCL's CTAs do not push a GEMM's last CTAs into a second
        # wave; pair it with NCCL_MAX_CTAS = comm_sms (bench.py --comm-sms).  0 = off.
        self.comm_sms = 0
        # the optimizer update of a bucket of layers runs on a side stream as soon as backward (and, data parallel, the bucket's
        # all-reduce) has finished it, next to the GEMMs of the remaining layers: only the last bucket's update is exposed.  Needs
        # clip_grad_norm off (the clip coefficient wants the whole gradient).  The update is the same arithmetic on the same
        # slices (FusedAdam.step_range); tests/test_dp_step_cpu.py compares it with the one-launch form.
        self.adam_overlap = True
        self.comm_stream = torch.cuda.Stream(device=self.dev) if world_size > 1 else None
        cb = self.dec.codebook
        self.z_lo, self.z_hi = float(cb.min()), float(cb.max())          # main.py:645-646,763 (global scalars)
        self.gen = torch.Generator().manual_seed(seed)
        self.noise_bank = torch.randn(self.nb_noise, self.noise_dim, generator=self.gen) if (self.noise_dim and self.nb_noise) else None
        if world_size > 1:
            self._sync_replicas()
        self.loss = torch.zeros(1, device=self.dev, dtype=F32)
        self.graph = None
        self.static = None
        self._next_prm = None           # augmentation parameters drawn ahead of the next replay()
        self.last_indices = None
        self.force_idx = None           # diagnostics: code indices to decode instead of the arg-min's
        self.debug = None               # diagnostics: a dict here receives the step's intermediates

    # ------------------------------------------------------------------ one step on device-resident inputs
    def _device_step(self, inp, out_feats, prm):
        mix, dec, clip, cut = self.mix, self.dec, self.clip, self.cut
        B = inp.shape[0]
        S, C = mix.S, mix.C
        N = self.cutn * B
        ops.fork(mix.zero_grad_arena, key="zero")      # 4 bytes per parameter of memset, off the critical path
        if self.normalize_input:                                         # main.py:734-735 (before the repeat / the loss targets)
            inp_n = torch.empty_like(inp)
            call("normalize_rows", inp, inp_n, B, inp.shape[1])
            inp = inp_n
        inp_net = inp
        if self.noise_dim:                                               # main.py:741-750: cat((inp_feats, noise), dim=1)
            E = inp.shape[1]
            inp_net = torch.empty(B, E + self.noise_dim, device=self.dev, dtype=F32)
            inp_net[:, :E].copy_(inp)
            if prm.get("mapper_noise") is not None:
                inp_net[:, E:].copy_(prm["mapper_noise"])
            else:
                inp_net[:, E:].normal_()
        z, sv_m = mix.forward(inp_net)                                   # [B*T, C] fp32
        zq, idx, zc = dec.quantize(z, self.z_lo, self.z_hi)
        if self.force_idx is not None:                                   # diagnostics only (tools/diag_fullsize.py): decode these codes
            idx = self.force_idx.to(idx.dtype).view(-1)
            zq = dec.codebook[idx.long()].to(BF16)
        self.last_indices = idx
        img, tape = dec.forward(zq.view(B, S, S, C))                     # [B, H, W, 3] fp32 in [0, 1]
        ops.join("noise")
        patches, sv_c, _ = cut.forward(img, prm)
        emb, sv_e = clip.forward(patches)
        demb = torch.empty(N, clip.E, device=self.dev, dtype=F32)
        if self.input_loss_coef != 0.0:                                  # main.py:812-824
            call("spherical_loss2", emb, out_feats, inp, self.loss, demb, None, N, B, clip.E, self.coef, self.input_loss_coef)
        else:
            call("spherical_loss", emb, out_feats, self.loss, demb, None, N, B, clip.E, self.coef)
        self.aux_loss.zero_()
        if self.debug is not None:
            self.debug.update(z=z.clone(), img=img.clone(), patches=patches.clone(), emb=emb.clone(), demb=demb.clone())
        # ---- backward
        dpatch = clip.backward(sv_e, demb)
        del sv_e
        dimg = cut.backward(sv_c, dpatch)
        del sv_c, dpatch
        if self.div is not None and self.diversity_mode == "all":        # every pair of the (repeated) batch, main.py:783-787
            self.div.forward_backward(img, B, 1, self.diversity_coef, dimg, self.aux_loss[2:3])
        elif self.div is not None and self.repeat > 1:                   # - diversity_coef * div, main.py:776-782,831
            self.div.forward_backward(img, self.repeat, B // self.repeat, self.diversity_coef, dimg, self.aux_loss[2:3])
        if self.tv_coef > 0:                                             # tv_coef * tv_loss(xr), main.py:769-773,831
            H = img.shape[1]
            call("tv_loss", img, self.aux_loss[1:2], dimg, B, H, img.shape[2], 3, self.tv_coef)
        if self.debug is not None:
            self.debug.update(dimg=dimg.clone())
        dzq = dec.backward(tape, dimg)                                   # [B*T, C] bf16 (straight-through to z)
        del tape
        if self.debug is not None:
            self.debug.update(dzq=dzq.clone())
        dzq32 = torch.empty(B * S * S, C, device=self.dev, dtype=F32)
        call("cast_bf16_f32", dzq, dzq32, dzq.numel())
        dz = torch.empty_like(dzq32)
        call("clamp_bwd", dzq32, z, dz, dz.numel(), self.z_lo, self.z_hi)
        if self.debug is not None:
            self.debug.update(dz=dz.clone())
        if self.l2_coef > 0:                                             # l2_coef * mean(z**2) on the pre-clamp z, main.py:758-762
            call("sumsq", z, self.aux_loss[0:1], z.numel())
            self.aux_loss[0:1].mul_(self.l2_coef / z.numel())
            call("axpy_f32", z, dz, 2.0 * self.l2_coef / z.numel(), z.numel())
        ops.join("zero")                      # the gradient arena was zeroed on a side stream while the forward ran (below)
        if self.world > 1 and self.bucket_layers > 0 and hasattr(mix, "layer_starts"):
            from .parallel import bucket_slices
            # bucket k (k < last): the gradients of mixer layers [L - (k+1)*n, L - k*n) (+ the final norm / output projection for
            # k = 0), complete once backward has finished the first of those layers; last bucket: whatever only completes when
            # backward has ended — the head of the arena (mixer.1) and the input projection `proj`, which the reference registers
            # AFTER the layers (mlp_mixer_pytorch.py:73-76) although its wgrad is the last GEMM of backward
            buckets = bucket_slices(mix.layer_starts(), mix.total, self.bucket_layers, late=mix.late_ranges())
            main, comm = torch.cuda.current_stream(), self.comm_stream
            first_layer_of_bucket = {mix.L - min(mix.L, (k + 1) * self.bucket_layers): k for k in range(len(buckets) - 1)}

            from . import _lib
            lib = _lib.load()

            early_adam = self.adam_overlap and self.opt.clip == 0 and len(buckets) > 1 and not self.tail_overlap
            if early_adam:
                self.opt.tick()                       # before the first bucket: step counter, bias corrections, lr (no clip coefficient)

            def reduce_bucket(slices, last=False):
                if self.comm_sms > 0:
                    lib.ffvc_set_option(b"sm_limit", max(2, ops.num_sms() - self.comm_sms))
                ev = torch.cuda.Event()
                ev.record(main)                       # the gradients of these slices are complete at this point of the main stream
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    for lo, hi in slices:
                        torch.distributed.all_reduce(mix.grad[lo:hi], group=self.pg)
                    if early_adam and not last:       # the bucket's update right behind its all-reduce, next to the remaining backward
                        for lo, hi in slices:
                            self.opt.step_range(lo, hi)

            def on_layer_done(k):
                if k in first_layer_of_bucket:
                    reduce_bucket(buckets[first_layer_of_bucket[k]])

            mix.backward(sv_m, dz, on_layer_done=on_layer_done)
            if self.tail_overlap and self.opt.clip == 0 and len(buckets) > 1:
                early = torch.cuda.Event()
                early.record(comm)                    # every bucket but the last has been issued on the side stream
                reduce_bucket(buckets[-1])
                del sv_m
                self.opt.tick()
                main.wait_event(early)
                for lo, hi in sorted(s_ for b in buckets[:-1] for s_ in b):
                    self.opt.step_range(lo, hi)       # overlaps the last bucket's all-reduce
                main.wait_stream(comm)
                lib.ffvc_set_option(b"sm_limit", 0)
                for lo, hi in buckets[-1]:
                    self.opt.step_range(lo, hi)
                return self.loss
            reduce_bucket(buckets[-1], last=True)
            main.wait_stream(comm)
            lib.ffvc_set_option(b"sm_limit", 0)
            if early_adam:
                del sv_m
                for lo, hi in buckets[-1]:
                    self.opt.step_range(lo, hi)
                return self.loss
        elif (self.world == 1 and self.adam_overlap and self.opt.clip == 0 and self.bucket_layers > 0 and hasattr(mix, "layer_starts")
              and ops.SIDE_STREAMS and mix.dev.type == "cuda"):
            from .parallel import bucket_slices
            buckets = bucket_slices(mix.layer_starts(), mix.total, self.bucket_layers, late=mix.late_ranges())
            first_layer_of_bucket = {mix.L - min(mix.L, (k + 1) * self.bucket_layers): k for k in range(len(buckets) - 1)}
            self.opt.tick()

            def on_layer_done(k):
                if k in first_layer_of_bucket:
                    sl = buckets[first_layer_of_bucket[k]]
                    ops.fork(lambda: [self.opt.step_range(lo, hi) for lo, hi in sl], key="adam")

            mix.backward(sv_m, dz, on_layer_done=on_layer_done)
            del sv_m
            ops.join("adam")
            for lo, hi in buckets[-1]:
                self.opt.step_range(lo, hi)
            return self.loss
        else:
            mix.backward(sv_m, dz)
            if self.world > 1:
                torch.distributed.all_reduce(mix.grad, group=self.pg)   # one NCCL all-reduce over NVLink (main.py:627)
        del sv_m
        self.opt.apply()
        return self.loss

    def _sync_replicas(self):
        """main.py:628-629 (hvd.broadcast_parameters / broadcast_optimizer_state): every rank starts from rank 0's weights and
        optimizer state, whatever each rank was built or resumed from; only gradients are averaged afterwards."""
        import torch.distributed as dist
        o = self.opt
        for t in (self.mix.arena, o.m, o.v, o.hyper) + ((o.ema,) if o.ema is not None else ()):
            dist.broadcast(t, src=0, group=self.pg)
        self.mix.ext_shadow_fresh = False
        self.mix._shadow_version = None

    def new_params(self, B):
        prm = sample_params(self.cutn * B, self.cut_size, self.gen, with_noise=False)
        if self.noise_bank is not None:      # main.py:742-746: `repeat` of the nb_noise fixed vectors, one per repeat group
            inds = torch.randperm(self.nb_noise, generator=self.gen)[:self.repeat]
            bs = B // self.repeat
            prm["mapper_noise"] = self.noise_bank[inds].repeat_interleave(bs, dim=0).contiguous()
        return prm

    def _stage_params(self, prm, B):
        N = self.cutn * B
        dev = self.dev
        out = {}
        for k in ("affine_inv", "persp_inv", "sat", "hue"):
            out[k] = prm[k].to(dev, non_blocking=True)
        out["erase"] = torch.tensor([int(t) for t in prm["erase"]], dtype=torch.int32).to(dev, non_blocking=True)
        out["mapper_noise"] = prm["mapper_noise"].to(dev, non_blocking=True) if prm.get("mapper_noise") is not None else None
        if "noise_raw" in prm:
            out["noise_raw"] = prm["noise_raw"].to(dev)
            out["facs"] = prm["facs"].to(dev)
        else:   # main.py:223-225: facs ~ U(0, noise_fac), noise ~ N(0,1), drawn on the device
            out["facs"] = torch.rand(N, device=dev) * 0.1
            out["noise_raw"] = torch.randn(N, 3, self.cut_size, self.cut_size, device=dev)
        return out

    def step(self, inp, out_feats=None, prm=None):
        """Eager step.  inp / out_feats: (B, clip_dim) fp32 CUDA tensors; prm: explicit cutout parameters or None."""
        if out_feats is None:
            out_feats = inp
        if self.repeat > 1:                                              # main.py:739-740
            inp, out_feats = inp.repeat(self.repeat, 1), out_feats.repeat(self.repeat, 1)
        B = inp.shape[0]
        prm = self._stage_params(prm if prm is not None else self.new_params(B), B)
        return self._device_step(inp.contiguous().float(), out_feats.contiguous().float(), prm)

    # ------------------------------------------------------------------ CUDA-graph path
    def capture(self, B, clip_dim):
        dev = self.dev
        N = self.cutn * B
        st = dict(inp=torch.zeros(B, clip_dim, device=dev), out=torch.zeros(B, clip_dim, device=dev),
                  affine_inv=torch.eye(3, device=dev).repeat(N, 1, 1).contiguous(),
                  persp_inv=torch.eye(3, device=dev).repeat(N, 1, 1).contiguous(),
                  sat=torch.ones(N, device=dev), hue=torch.zeros(N, device=dev),
                  erase=torch.zeros(4, device=dev, dtype=torch.int32))
        st["mapper_noise"] = torch.zeros(B, self.noise_dim, device=dev) if self.noise_bank is not None else None
        self.static = st

        st["facs"] = torch.empty(N, device=dev)
        st["noise_raw"] = torch.empty(N, 3, self.cut_size, self.cut_size, device=dev)

        def draw_noise():          # main.py:223-225: facs ~ U(0, noise_fac), noise ~ N(0, 1); 2.4 KB per cutout of RNG output per step
            st["facs"].uniform_(0.0, 0.1)
            st["noise_raw"].normal_()

        def body():
            ops.fork(draw_noise, key="noise")       # next to the mapper / decoder forward; joined in front of the cutouts
            prm = dict(affine_inv=st["affine_inv"], persp_inv=st["persp_inv"], sat=st["sat"], hue=st["hue"], erase=st["erase"],
                       mapper_noise=st["mapper_noise"], facs=st["facs"], noise_raw=st["noise_raw"])
            self._device_step(st["inp"], st["out"], prm)

        # warm-up on a side stream (allocator + lazy init), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may touch CUDA while we capture (world > 1)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            body()
        return self

    def load_static(self, inp, out_feats, prm):
        """copy one step's inputs (embeddings, augmentation parameters) into the captured graph's static device buffers"""
        st = self.static
        B = st["inp"].shape[0]
        if out_feats is None:
            out_feats = inp
        if self.repeat > 1 and inp.shape[0] * self.repeat == B:
            inp, out_feats = inp.repeat(self.repeat, 1), out_feats.repeat(self.repeat, 1)
        st["inp"].copy_(inp, non_blocking=True)
        st["out"].copy_(out_feats, non_blocking=True)
        for k in ("affine_inv", "persp_inv", "sat", "hue"):
            st[k].copy_(prm[k], non_blocking=True)
        st["erase"].copy_(torch.tensor([int(t) for t in prm["erase"]], dtype=torch.int32), non_blocking=True)
        if st.get("mapper_noise") is not None:
            st["mapper_noise"].copy_(prm["mapper_noise"], non_blocking=True)

    def replay(self, inp, out_feats=None, prm=None):
        """inp may be a pinned host tensor (H2D copy on the current stream) or a device tensor.  With prm=None the
        augmentation parameters of the NEXT call are drawn right after this step's graph launch, i.e. while the GPU works."""
        st = self.static
        B = st["inp"].shape[0]
        if out_feats is None:
            out_feats = inp
        if self.repeat > 1 and inp.shape[0] * self.repeat == B:          # main.py:739-740: the prompt batch is repeated, as step() does
            inp, out_feats = inp.repeat(self.repeat, 1), out_feats.repeat(self.repeat, 1)
        presample = prm is None
        if presample:
            prm = self._next_prm if self._next_prm is not None else self.new_params(B)
            self._next_prm = None
        self.load_static(inp, out_feats, prm)
        self.graph.replay()
        if presample:
            self._next_prm = self.new_params(B)
        return self.loss
