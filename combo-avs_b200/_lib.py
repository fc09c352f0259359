"""ctypes binding of libmsda_b200.so (C ABI declared in include/msda_b200.h).

There is deliberately no fallback: if the shared library is missing or fails to load, importing the op raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsda_b200.so")

F32, BF16, F64 = 0, 1, 2

# every symbol include/msda_b200.h declares
EXPORTS = (
    "msda_forward",
    "msda_backward",
    "msda_backward_workspace_bytes",
    "msda_select_variant",
    "msda_variant_names",
    "msda_last_variant",
    "msda_launch_count",
    "msda_last_error",
    "msda_version",
    "msda_host_pipeline_create",
    "msda_host_pipeline_destroy",
    "msda_host_pipeline_error",
    "msda_forward_backward_host",
    "msda_fused_supported",
    "msda_fused_forward",
    "msda_fused_backward",
    "msda_linear_supported",
    "msda_linear_forward",
    "msda_linear_wgrad_supported",
    "msda_linear_wgrad",
    "msda_layernorm_supported",
    "msda_layernorm_backward_workspace_bytes",
    "msda_add_layernorm_forward",
    "msda_layernorm_backward",
    "msda_colsum_supported",
    "msda_colsum_workspace_bytes",
    "msda_colsum",
)

_lib = None


class MSDALibraryError(ImportError):
    pass


def load():
    """Load libmsda_b200.so once; raise MSDALibraryError with build instructions if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MSDALibraryError(
            "\n\n%s is missing. Build the sm_100a kernels first:\n\t`python combo-avs_b200/build.py`\n"
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU or PyTorch fallback.\n" % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the host
        raise MSDALibraryError("cannot load %s: %s" % (LIB_PATH, e)) from e
    vp, ip, i, sz = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.msda_forward.argtypes = [i, i, vp, ip, ip, vp, vp, vp, i, i, i, i, i, i, i, i, ip, ip, i, vp]
    lib.msda_forward.restype = i
    lib.msda_backward.argtypes = [i, i, vp, ip, ip, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, ip, ip, vp, sz, i, vp]
    lib.msda_backward.restype = i
    lib.msda_backward_workspace_bytes.argtypes = [i] * 9
    lib.msda_backward_workspace_bytes.restype = sz
    lib.msda_select_variant.argtypes = [i, ctypes.c_char_p]
    lib.msda_select_variant.restype = i
    lib.msda_variant_names.argtypes = [i]
    lib.msda_variant_names.restype = ctypes.c_char_p
    lib.msda_last_variant.argtypes = [i]
    lib.msda_last_variant.restype = ctypes.c_char_p
    lib.msda_launch_count.argtypes = []
    lib.msda_launch_count.restype = ctypes.c_uint64
    lib.msda_last_error.argtypes = []
    lib.msda_last_error.restype = ctypes.c_char_p
    lib.msda_version.argtypes = []
    lib.msda_version.restype = i
    lib.msda_linear_supported.argtypes = [i, i, i]
    lib.msda_linear_supported.restype = i
    lib.msda_linear_forward.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp]
    lib.msda_linear_forward.restype = i
    lib.msda_linear_wgrad_supported.argtypes = [i, i, i]
    lib.msda_linear_wgrad_supported.restype = i
    lib.msda_linear_wgrad.argtypes = [vp, vp, vp, i, i, i, i, vp]
    lib.msda_linear_wgrad.restype = i
    lib.msda_layernorm_supported.argtypes = [i]
    lib.msda_layernorm_supported.restype = i
    lib.msda_layernorm_backward_workspace_bytes.argtypes = [i, i, i]
    lib.msda_layernorm_backward_workspace_bytes.restype = sz
    lib.msda_add_layernorm_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, ctypes.c_float, i, vp]
    lib.msda_add_layernorm_forward.restype = i
    lib.msda_layernorm_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, i, i, i, vp]
    lib.msda_layernorm_backward.restype = i
    lib.msda_colsum_supported.argtypes = [i]
    lib.msda_colsum_supported.restype = i
    lib.msda_colsum_workspace_bytes.argtypes = [i, i, i]
    lib.msda_colsum_workspace_bytes.restype = sz
    lib.msda_colsum.argtypes = [vp, vp, vp, sz, i, i, i, vp]
    lib.msda_colsum.restype = i
    lib.msda_host_pipeline_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), i, i, i]
    lib.msda_host_pipeline_create.restype = i
    lib.msda_host_pipeline_destroy.argtypes = [vp]
    lib.msda_host_pipeline_destroy.restype = i
    lib.msda_host_pipeline_error.argtypes = [vp]
    lib.msda_host_pipeline_error.restype = ctypes.c_char_p
    lib.msda_forward_backward_host.argtypes = [vp, i, i, vp, ip, ip, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i,
                                               ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    lib.msda_forward_backward_host.restype = i
    lib.msda_fused_supported.argtypes = [i, i, i, i, i, i, i, i, ip, ip, i]
    lib.msda_fused_supported.restype = i
    lib.msda_fused_forward.argtypes = [i, vp, ip, ip, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, ip, ip, i, vp]
    lib.msda_fused_forward.restype = i
    lib.msda_fused_backward.argtypes = [i, vp, ip, ip, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, ip, ip, i, vp]
    lib.msda_fused_backward.restype = i
    _lib = lib
    return lib


def last_error():
    return load().msda_last_error().decode("utf-8", "replace")


def check(rc):
    """Translate a status code of the C ABI into the exception the reference would have raised."""
    if rc == 0:
        return
    msg = last_error()
    if rc in (-2, -3, -5):  # dimension / dtype / variant problems
        raise ValueError("libmsda_b200: %s (status %d)" % (msg, rc))
    raise RuntimeError("libmsda_b200: %s (status %d)" % (msg, rc))


def select_variant(pass_name, name):
    """Force a kernel variant ('auto' to undo). pass_name: 'forward' | 'backward'."""
    idx = {"forward": 0, "fwd": 0, "backward": 1, "bwd": 1}[pass_name]
    check(load().msda_select_variant(idx, name.encode()))


def variant_names(pass_name):
    idx = {"forward": 0, "fwd": 0, "backward": 1, "bwd": 1}[pass_name]
    return load().msda_variant_names(idx).decode().split(",")


def last_variant(pass_name):
    idx = {"forward": 0, "fwd": 0, "backward": 1, "bwd": 1}[pass_name]
    return load().msda_last_variant(idx).decode()


def launch_count():
    return int(load().msda_launch_count())
