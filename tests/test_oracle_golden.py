"""CPU: pin the oracle (C restatement + torch restatement of the reference's CPU path) to golden vectors that
oracle/gen_golden.py produced from the reference's own ms_deform_attn_core_pytorch."""
import numpy as np
import pytest
import torch

import oracle
from oracle.core_pytorch import msda_core_pytorch
from conftest import load_golden, make_inputs, rel_err, S4_SHAPES

CASES = ["testpy_fwd_double", "testpy_fwd_float", "testpy_grad_D30", "testpy_grad_D32", "testpy_grad_D64",
         "testpy_grad_D71", "s4pyr_unit", "s4pyr_edge", "ragged_levels"]


def _grad_out(g):
    return g["grad_out"].astype(np.float64) if "grad_out" in g else np.ones_like(g["out"])


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_f64_matches_reference(name):
    g = load_golden(name)
    v, loc, a = (g[k].astype(np.float64) for k in ("value", "loc", "attn"))
    out = oracle.forward(v, g["shapes"], g["start"], loc, a)
    assert out.shape == g["out"].shape
    assert rel_err(out, g["out"]) < 1e-13
    gv, gl, ga = oracle.backward(v, g["shapes"], g["start"], loc, a, _grad_out(g))
    assert rel_err(gv, g["grad_value"]) < 1e-13
    assert rel_err(gl, g["grad_loc"]) < 1e-12
    assert rel_err(ga, g["grad_attn"]) < 1e-13


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_f32_within_fp32_tolerance(name):
    g = load_golden(name)
    v, loc, a = (g[k].astype(np.float32) for k in ("value", "loc", "attn"))
    out = oracle.forward(v, g["shapes"], g["start"], loc, a)
    assert out.dtype == np.float32
    assert rel_err(out, g["out"]) < 1e-5  # north_star forward tolerance
    gv, gl, ga = oracle.backward(v, g["shapes"], g["start"], loc, a, _grad_out(g).astype(np.float32))
    assert rel_err(gv, g["grad_value"]) < 1e-4  # north_star gradient tolerance
    assert rel_err(gl, g["grad_loc"]) < 1e-4
    assert rel_err(ga, g["grad_attn"]) < 1e-4


def test_testpy_float_path_matches_reference_float_output():
    # what the reference's test.py:62 compares the CUDA op against, in float
    g = load_golden("testpy_fwd_float")
    out = oracle.forward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"])
    assert np.allclose(out, g["out_f32path"], rtol=1e-2, atol=1e-3)  # test.py:67 tolerance
    assert rel_err(out, g["out_f32path"]) < 1e-6


@pytest.mark.parametrize("name", ["testpy_fwd_double", "s4pyr_edge", "ragged_levels"])
def test_torch_restatement_matches_reference(name):
    g = load_golden(name)
    v = torch.from_numpy(g["value"]).double().requires_grad_(True)
    loc = torch.from_numpy(g["loc"]).double().requires_grad_(True)
    a = torch.from_numpy(g["attn"]).double().requires_grad_(True)
    out = msda_core_pytorch(v, torch.from_numpy(g["shapes"]), loc, a)
    assert rel_err(out.detach().numpy(), g["out"]) < 1e-13
    out.backward(torch.from_numpy(_grad_out(g)))
    assert rel_err(v.grad.numpy(), g["grad_value"]) < 1e-13
    assert rel_err(loc.grad.numpy(), g["grad_loc"]) < 1e-12
    assert rel_err(a.grad.numpy(), g["grad_attn"]) < 1e-13


def test_c_oracle_equals_torch_restatement_on_s4_frames():
    # a full S4 frame pair (all 8 heads, 1029 queries), locations beyond the map included
    d = make_inputs(2, S4_SHAPES, 8, 32, 1029, 4, seed=5, lo=-0.2, hi=1.2, dtype=np.float64)
    out = oracle.forward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"])
    ref = msda_core_pytorch(torch.from_numpy(d["value"]), torch.from_numpy(d["shapes"]), torch.from_numpy(d["loc"]),
                            torch.from_numpy(d["attn"])).numpy()
    assert rel_err(out, ref) < 1e-13


def test_oracle_properties():
    d = make_inputs(2, [(5, 3), (2, 4)], 3, 6, 11, 2, seed=9, lo=-0.3, hi=1.3, dtype=np.float64)
    args = (d["shapes"], d["start"], d["loc"], d["attn"])
    out = oracle.forward(d["value"], *args)
    # linear in value
    assert rel_err(oracle.forward(d["value"] * 3.0, *args), out * 3.0) < 1e-14
    # adjoint: <out(v), g> == <v, grad_value(g)>
    gv, gl, ga = oracle.backward(d["value"], *args, d["grad_out"])
    lhs = float((out * d["grad_out"]).sum())
    rhs = float((d["value"] * gv).sum())
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)
    # grad_attn is the un-weighted sample: sum_lp attn*grad_attn == <out, g> per (n,q,m)
    per_qm = (d["attn"] * ga).sum(axis=(-1, -2))
    N, Lq, M = per_qm.shape
    ref = (out * d["grad_out"]).reshape(N, Lq, M, -1).sum(-1)
    assert rel_err(per_qm, ref) < 1e-12
    # samples entirely outside contribute nothing
    far = d["loc"].copy()
    far[..., 0] = 5.0
    assert np.abs(oracle.forward(d["value"], d["shapes"], d["start"], far, d["attn"])).max() == 0.0
    gv2, gl2, ga2 = oracle.backward(d["value"], d["shapes"], d["start"], far, d["attn"], d["grad_out"])
    assert np.abs(gv2).max() == 0.0 and np.abs(gl2).max() == 0.0 and np.abs(ga2).max() == 0.0


def test_oracle_empty_inputs():
    d = make_inputs(0, [(2, 2)], 2, 4, 3, 2, dtype=np.float64)
    assert oracle.forward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"]).shape == (0, 3, 8)
    d = make_inputs(2, [(2, 2)], 2, 4, 0, 2, dtype=np.float64)
    assert oracle.forward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"]).shape == (2, 0, 8)
    gv, gl, ga = oracle.backward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"], d["grad_out"])
    assert np.abs(gv).max() == 0.0 and gl.size == 0 and ga.size == 0
