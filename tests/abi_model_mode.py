"""pytest plugin (opt-in: `python -m pytest tests -m gpu -p abi_model_mode`): run the GPU test modules on the CPU with every
`ffvc_*` launch served by tests/abi_model.py.  Two uses, both test infrastructure only:
  * the expectations the GPU tests hold for each kernel (written against torch / the oracle and green on the B200) become a check
    of the MODEL's faithfulness to the kernels' contracts;
  * a host-side change can be regression-tested against the whole GPU suite before GPU minutes are spent on it.
Variants that only select a different kernel for the same contract (tile shapes, epilogue warps, option switches) collapse onto
the one model function; tests of purely device-side behaviour (CUDA graphs, launch counters) are skipped."""
import contextlib

import pytest
import torch

import abi_model

SKIP = ("cuda_graph", "launch_count", "tma_store_epilogue", "stream_k", "argmin_epilogue", "gelu_epilogue_matches_exact_erf",
        "two_rank_step_on_one_gpu", "async_checkpoint_snapshot", "abi_model_vs_kernels", "copies_gathers", "small_arithmetic",
        "full_size_step_is_invariant")        # 64 prompts of the full architecture: minutes of CPU time (its logic runs at B = 2 below)


def pytest_configure(config):
    from feed_forward_vqgan_clip_b200 import ops
    import feed_forward_vqgan_clip_b200 as pkg
    import importlib
    import pkgutil
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: None
    ops.gemm_raw = abi_model.gemm_raw
    ops.gemm = lambda a, b, out, M, N, K, **kw: abi_model.gemm_raw(a, b, out, M, N, K, **kw)
    ops.call = abi_model.call
    ops.require_cuda = lambda dev, what: None
    ops.launch_count = lambda: 1000
    ops.reset_launch_count = lambda: None
    for m in pkgutil.iter_modules(pkg.__path__):
        if m.name in ("build",) or m.name.startswith("lib"):      # the shared object sits in the package directory too
            continue
        mod = importlib.import_module(pkg.__name__ + "." + m.name)
        if getattr(mod, "call", None) is not None and m.name != "ops":
            mod.call = abi_model.call
    _to, _empty_like = torch.Tensor.to, None

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda") else x
                  for x in a)
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return _to(self, *a, **k)
    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self
    _mto = torch.nn.Module.to
    torch.nn.Module.to = lambda self, *a, **k: self if (a and isinstance(a[0], (str, torch.device)) and str(a[0]).startswith("cuda")) \
        else _mto(self, *a, **k)
    torch.nn.Module.cuda = lambda self, *a, **k: self


@pytest.hookimpl(trylast=True)
def pytest_collection_modifyitems(config, items):
    for item in items:
        # conftest.py skips the gpu-marked tests on a machine without CUDA: that is exactly where this plugin runs them
        item.own_markers = [m for m in item.own_markers if not (m.name == "skip" and m.kwargs.get("reason") == "no CUDA device")]
        mod = item.module
        for name in ("DEV", "DEVICE"):
            if isinstance(getattr(mod, name, None), str) and getattr(mod, name).startswith("cuda"):
                setattr(mod, name, "cpu")
        if "call" in vars(mod) and getattr(mod, "call") is not abi_model.call:
            mod.call = abi_model.call
        if any(s in item.name for s in SKIP):
            item.add_marker(pytest.mark.skip(reason="device-side behaviour: not meaningful on the CPU model"))
