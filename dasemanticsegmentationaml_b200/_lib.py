"""ctypes binding of libb200seg.so — the C-ABI boundary declared in include/b200seg.h.

There is deliberately no fallback: if the shared library is missing or the device is not a
compute-capability-10.x part the first call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200seg.so")

_lib = None
_checked_devices = set()


class B200Error(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first use; raises B200Error when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                "libb200seg.so is not built (%s). Run `python -m dasemanticsegmentationaml_b200.build`; "
                "this package has no CPU or PyTorch fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b200_last_error.restype = ctypes.c_char_p
        _lib.b200_version.restype = ctypes.c_int
    return _lib


def last_error():
    return lib().b200_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise B200Error("%s failed (%d): %s" % (what or "libb200seg call", rc, last_error()))


def ensure_device(index):
    """Refuse to run on anything but an sm_100 device (no multi-arch dispatch by design)."""
    if index in _checked_devices:
        return
    check(lib().b200_check_device(), "b200_check_device")
    _checked_devices.add(index)


def ptr(t):
    """Raw device pointer of a tensor (or None)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def int_array(values):
    arr = (ctypes.c_int * len(values))(*[int(v) for v in values])
    return arr


c_int = ctypes.c_int
c_float = ctypes.c_float
c_int64 = ctypes.c_int64
c_void_p = ctypes.c_void_p
