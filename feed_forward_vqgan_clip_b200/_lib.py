"""ctypes binding of libffvc_sm100.so.  Argument types are derived from include/ffvc.h itself, so the binding
cannot drift from the C ABI.  Fails loudly when the library is absent: there is no CPU / PyTorch fallback."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFVC_LIB") or os.path.join(_HERE, "libffvc_sm100.so")     # FFVC_LIB: another build of the same ABI (A/B runs)
HEADER_PATH = os.path.join(_HERE, "..", "include", "ffvc.h")
_lib = None
_decls = None


class GemmParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p),
        ("a_mode", C.c_int), ("b_mode", C.c_int),
        ("a_ld", C.c_int64), ("b_ld", C.c_int64),
        ("a_batch_role", C.c_int), ("b_batch_role", C.c_int),
        ("a_batch_stride", C.c_int64), ("b_batch_stride", C.c_int64),
        ("a_batch_stride_inner", C.c_int64), ("b_batch_stride_inner", C.c_int64),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("batch", C.c_int), ("batch_inner", C.c_int), ("k_segs", C.c_int), ("splits", C.c_int), ("block_n", C.c_int),
        ("tile_m", C.c_int), ("two_cta", C.c_int), ("epi_warps", C.c_int),
        ("conv_n", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_c", C.c_int),
        ("out", C.c_void_p), ("pre_out", C.c_void_p), ("aux", C.c_void_p), ("res", C.c_void_p),
        ("bias", C.c_void_p),
        ("ldc", C.c_int64), ("out_batch_stride", C.c_int64), ("out_batch_stride_inner", C.c_int64),
        ("out_fp32", C.c_int), ("atomic", C.c_int), ("bias_mode", C.c_int), ("act", C.c_int),
        ("mul_mode", C.c_int), ("alpha", C.c_float), ("argmin_out", C.c_void_p),
    ]


def _ctype(t):
    t = t.strip()
    if t == "const char*":
        return C.c_char_p
    if t.endswith("*"):
        return C.c_void_p
    if t == "int":
        return C.c_int
    if t == "long long":
        return C.c_longlong
    if t == "float":
        return C.c_float
    if t == "void":
        return None
    raise ValueError("unmapped C type in ffvc.h: %r" % t)


def header_declarations():
    """{name: (restype, [argtypes])} for every function declared in include/ffvc.h."""
    global _decls
    if _decls is not None:
        return _decls
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"typedef struct \{.*?\} \w+;", " ", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(const char\*|long long|int|void)\s+(ffvc_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("void", ""):
                continue
            a = re.sub(r"\s+", " ", a)
            mm = re.match(r"^(.*?)(\w+)$", a)           # strip the parameter name
            argtypes.append(_ctype(mm.group(1).strip().replace(" *", "*")))
        decls[name] = (_ctype(ret), argtypes)
    _decls = decls
    return decls


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libffvc_sm100.so not found at %s — build it with `python -m feed_forward_vqgan_clip_b200.build` "
            "(or __graft_entry__.build()).  There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (ret, argtypes) in header_declarations().items():
        if os.environ.get("FFVC_LIB") and not hasattr(lib, name):
            continue                     # an older build under A/B: symbols added since are simply absent
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


class FFVCError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise FFVCError("ffvc error %d: %s" % (rc, load().ffvc_last_error().decode()))
