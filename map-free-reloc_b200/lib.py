"""ctypes binding of libmfr_b200.so. There is NO CPU fallback: if the library is missing or a
CUDA device is absent, compute entry points raise."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmfr_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mfr_b200.h")

_lib = None

c_f32p = ctypes.c_void_p
c_ll = ctypes.c_longlong
c_int = ctypes.c_int
c_float = ctypes.c_float
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mfr_b200.h one to one
SIGNATURES = {
    "mfr_version": (c_int, []),
    "mfr_device_sm_count": (c_int, []),
    "mfr_conv2d_nhwc_tf32": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_ll, c_vp, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_vp, c_vp, c_ll, c_int, c_float,
                                     c_vp, c_ll, c_vp]),
}


class MfrError(RuntimeError):
    pass


def load():
    """Loads the shared library (building is done by ``build.py`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MfrError(
            f"{LIB_PATH} not found: run `python map-free-reloc_b200/build.py` "
            "(there is no CPU fallback for the CUDA path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise MfrError(f"libmfr_b200 call failed rc={rc} {what}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
