"""CPU: the C-ABI library loads and exports every symbol include/msda_b200.h declares; argument errors are reported
through status codes + msda_last_error() without touching a GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import combo_avs_b200  # noqa: F401  (loads the library or raises)
    from combo_avs_b200 import _lib
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert set(declared) == set(lib.EXPORTS), (declared, lib.EXPORTS)
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), "libmsda_b200.so does not export %s" % name


def test_version_and_variants(lib):
    L = lib.load()
    assert L.msda_version() == 102
    assert "auto" in lib.variant_names("forward") and "generic" in lib.variant_names("backward")
    assert L.msda_select_variant(0, b"no-such-kernel") == -5
    assert b"unknown forward variant" in L.msda_last_error()
    assert L.msda_select_variant(7, b"auto") == -5
    assert L.msda_select_variant(0, b"auto") == 0
    with pytest.raises(ValueError):
        lib.select_variant("backward", "bogus")


def test_argument_errors_need_no_gpu(lib):
    L = lib.load()
    null = None
    # dtype 9 does not exist
    rc = L.msda_forward(9, 9, null, null, null, null, null, null, 1, 4, 1, 4, 1, 1, 1, 64, null, null, 0, null)
    assert rc == -3 and b"dtype" in L.msda_last_error()
    # f32 values with bf16 locations is not a supported combination
    rc = L.msda_forward(0, 1, null, null, null, null, null, null, 1, 4, 1, 4, 1, 1, 1, 64, null, null, 0, null)
    assert rc == -3
    # non-positive head count
    rc = L.msda_forward(0, 0, null, null, null, null, null, null, 1, 4, 0, 4, 1, 1, 1, 64, null, null, 0, null)
    assert rc == -2
    # the reference's im2col_step rule (ms_deform_attn_cuda.cu:55-57): N % min(N, step) == 0
    rc = L.msda_forward(0, 0, null, null, null, null, null, null, 160, 4, 1, 4, 1, 1, 1, 128, null, null, 0, null)
    assert rc == -6 and b"must divide" in L.msda_last_error()
    rc = L.msda_backward(0, 0, null, null, null, null, null, null, null, null, null, 3, 4, 1, 4, 1, 1, 1, 2, null, null, null, 0, 0, null)
    assert rc == -6
    # NULL tensors
    rc = L.msda_forward(0, 0, null, null, null, null, null, null, 1, 4, 1, 4, 1, 1, 1, 64, null, null, 0, null)
    assert rc == -1
    with pytest.raises(RuntimeError):
        lib.check(rc)


def test_workspace_query(lib):
    L = lib.load()
    assert L.msda_backward_workspace_bytes(0, 0, 40, 1029, 8, 32, 3, 1029, 4) == 0
    # no fp32 variant needs scratch, whichever is forced (the experimental pipe32 backward of round 1, which did, is gone)
    assert L.msda_select_variant(1, b"sorted32") == 0
    assert L.msda_backward_workspace_bytes(0, 0, 40, 1029, 8, 32, 3, 1029, 4) == 0
    assert L.msda_select_variant(1, b"pipe32") != 0
    assert L.msda_select_variant(1, b"auto") == 0
    assert L.msda_backward_workspace_bytes(1, 0, 2, 10, 2, 32, 1, 5, 4) == 4 * 2 * 10 * 2 * 32
    assert L.msda_backward_workspace_bytes(1, 1, 0, 10, 2, 32, 1, 5, 4) == 0


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under combo-avs_b200/ or include/ may reference it
    for base in ("combo-avs_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in text.lower(), "%s mentions the oracle" % os.path.join(dirpath, f)
                    assert "grid_sample" not in text, "%s holds a PyTorch fallback" % f


def test_fused_and_host_entry_points_validate_without_a_gpu(lib):
    """Argument errors of the newer entry points are reported before anything touches the device."""
    L = lib.load()
    null = None
    # fused forward / backward: the reference's batch rule and dtype checks come first (status < 0), as for msda_forward
    rc = L.msda_fused_forward(0, null, null, null, null, null, null, null, null, null, 160, 4, 1, 32, 1, 1, 4, 128, null, null, 0, null)
    assert rc == -6 and b"must divide" in L.msda_last_error()
    rc = L.msda_fused_forward(7, null, null, null, null, null, null, null, null, null, 1, 4, 1, 32, 1, 1, 4, 128, null, null, 0, null)
    assert rc == -3
    rc = L.msda_fused_backward(0, null, null, null, null, null, null, null, null, null, null, null, 3, 4, 1, 32, 1, 1, 4, 2, null, null, 0, null)
    assert rc == -6
    # shapes the fused path does not take are a plain "no" (no host geometry, wrong dtype)
    assert L.msda_fused_supported(0, 40, 1029, 8, 32, 3, 1029, 4, null, null, 0) == 0
    assert L.msda_fused_supported(1, 40, 1029, 8, 32, 3, 1029, 4, null, null, 0) == 0
    # host pipeline: bad construction arguments, NULL handle
    import ctypes
    h = ctypes.c_void_p()
    assert L.msda_host_pipeline_create(ctypes.byref(h), 0, 0, 3) == -2      # chunk_frames < 1
    assert L.msda_host_pipeline_create(ctypes.byref(h), 0, 4, 9) == -2      # slots out of range
    assert L.msda_host_pipeline_create(None, 0, 4, 3) == -1
    assert L.msda_host_pipeline_destroy(None) == 0
    assert L.msda_forward_backward_host(None, 0, 0, null, null, null, null, null, null, null, null, null, null,
                                        1, 4, 1, 32, 1, 1, 4, 128, None, None) == -1
    assert L.msda_host_pipeline_error(None) == b"NULL pipeline"
