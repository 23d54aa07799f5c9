"""ctypes binding of libffvc_sm100.so (include/ffvc.h).  Fails loudly when the library is absent:
there is no CPU or PyTorch fallback for the hot path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libffvc_sm100.so")
_lib = None


class GemmParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p),
        ("a_mode", C.c_int), ("b_mode", C.c_int),
        ("a_ld", C.c_int64), ("b_ld", C.c_int64),
        ("a_batch_role", C.c_int), ("b_batch_role", C.c_int),
        ("a_batch_stride", C.c_int64), ("b_batch_stride", C.c_int64),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("batch", C.c_int), ("k_segs", C.c_int), ("splits", C.c_int), ("block_n", C.c_int),
        ("conv_n", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_c", C.c_int),
        ("out", C.c_void_p), ("pre_out", C.c_void_p), ("aux", C.c_void_p), ("res", C.c_void_p),
        ("bias", C.c_void_p),
        ("ldc", C.c_int64), ("out_batch_stride", C.c_int64),
        ("out_fp32", C.c_int), ("atomic", C.c_int), ("bias_mode", C.c_int), ("act", C.c_int),
        ("mul_mode", C.c_int), ("alpha", C.c_float),
    ]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libffvc_sm100.so not found at %s — build it with `python -m feed_forward_vqgan_clip_b200.build` "
            "(or __graft_entry__.build()).  There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.ffvc_last_error.restype = C.c_char_p
    lib.ffvc_launch_count.restype = C.c_longlong
    lib.ffvc_sizeof.restype = C.c_int
    lib.ffvc_sizeof.argtypes = [C.c_char_p]
    _lib = lib
    return lib


class FFVCError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise FFVCError("ffvc error %d: %s" % (rc, load().ffvc_last_error().decode()))
