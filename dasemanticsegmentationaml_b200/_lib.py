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
_arity = None          # {entry point: parameter count}, parsed from include/b200seg.h
_arity_checked = set()


def _load_arity():
    """Parameter counts of every prototype in include/b200seg.h: ctypes cannot see C prototypes, so
    a wrapper that drifts from the header would otherwise corrupt the call frame silently."""
    import re
    global _arity
    _arity = {}
    header = os.path.join(os.path.dirname(_HERE), "include", "b200seg.h")
    if not os.path.exists(header):
        return
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    for m in re.finditer(r"\b(b200_\w+)\s*\(([^)]*)\)\s*;", text):
        params = m.group(2).strip()
        _arity[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1


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


# kernels launched per entry point when it is more than one (for the launch counter)
_MULTI = {"b200_radix_select_desc": 9, "b200_ohem_reduce": 2, "b200_fc_small_bwd": 2}
launch_count = 0      # kernels launched through this binding since import
_profile = None       # None, or a list of (name, start_event, end_event, flops, bytes)


def call(name, *args, flops=0.0, nbytes=0.0, tag=None):
    """Invoke an entry point on the current stream, raise on error, count its launches and, while
    profiling, bracket it with CUDA events on the launching stream."""
    global launch_count
    fn = getattr(lib(), name)
    if name not in _arity_checked:
        if _arity is None:
            _load_arity()
        want = _arity.get(name)
        if want is not None and want != len(args):
            raise B200Error("%s: wrapper passes %d arguments, include/b200seg.h declares %d" % (name, len(args), want))
        _arity_checked.add(name)
    if _profile is None:
        rc = fn(*args)
    else:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _profile.append((name, e0, e1, flops, nbytes, tag))
    if rc != 0:
        raise B200Error("%s failed (%d): %s" % (name, rc, last_error()))
    launch_count += _MULTI.get(name, 1)


def profile_start(host_ahead_ms=0.0):
    """Start bracketing every entry-point call with CUDA events.  An eager step is HOST bound (Python +
    ctypes + tensor-map encoding per launch), so a start event recorded on an idle GPU would charge the
    host's launch latency to the kernel; host_ahead_ms > 0 first parks the stream behind a spin kernel of
    about that length so that the events and launches queue up and the events bracket device time only."""
    global _profile
    if host_ahead_ms > 0:
        import torch
        torch.cuda._sleep(int(host_ahead_ms * 1e-3 * 1.9e9))
    _profile = []


def profile_stop():
    """-> {name: dict(calls, ms, flops, bytes)} for the calls made since profile_start()."""
    global _profile
    import torch
    torch.cuda.synchronize()
    agg = {}
    detail = {}
    for name, e0, e1, flops, nbytes, tag in _profile or []:
        ms = e0.elapsed_time(e1)
        for key, table in ((name, agg), ("%s %s" % (name, tag), detail)):
            a = table.setdefault(key, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            a["calls"] += 1
            a["ms"] += ms
            a["flops"] += flops
            a["bytes"] += nbytes
    _profile = None
    global last_profile_detail
    last_profile_detail = detail
    return agg


last_profile_detail = {}


def ensure_device(index):
    """Refuse to run on anything but an sm_100 device (no multi-arch dispatch by design)."""
    if index in _checked_devices:
        return
    check(lib().b200_check_device(), "b200_check_device")
    _checked_devices.add(index)


def set_device_index(index):
    """Device whose current stream the launches go to (one process drives one GPU)."""
    global _device_index
    _device_index = int(index)


def ptr(t):
    """Raw device pointer of a tensor (or None)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


_raw_stream = None
_device_index = 0


def stream():
    """Raw cudaStream_t of torch's current stream (the capture stream while a CUDA graph records)."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        if _raw_stream is None:
            _raw_stream = lambda dev: torch.cuda.current_stream(dev).cuda_stream  # noqa: E731
    return ctypes.c_void_p(_raw_stream(_device_index))


def int_array(values):
    arr = (ctypes.c_int * len(values))(*[int(v) for v in values])
    return arr


c_int = ctypes.c_int
c_float = ctypes.c_float
c_int64 = ctypes.c_int64
c_void_p = ctypes.c_void_p
