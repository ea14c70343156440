"""GPU parity for the whole-field kernels behind the interpreter: box / edge stencils (post_processing.py:552-623) and
the vector-field modifier pipeline (vector_modification_functions.py), against the oracle and the reference's own
outputs (tests/golden/fields.npz)."""
import os

import numpy as np
import pytest

from oracle import fields_np as F
from field_cases import AVG_CASES, EDGE_CASES, RES, vector_inputs, pipelines

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "fields.npz"))


@pytest.mark.parametrize("k", range(len(AVG_CASES)))
def test_conv_averaging_matches_reference_outputs(gold, k):
    import aegolius_b200 as ab
    shape, ks, it = AVG_CASES[k]
    u = gold[f"avg{k}_in"]
    got = ab.conv_averaging(u, ks, it)
    assert got.shape == tuple(shape) and got.dtype == np.float64
    assert np.max(np.abs(got - gold[f"avg{k}_out"])) <= 1e-13
    got32 = ab.conv_averaging(u.astype(np.float32), ks, it)
    assert got32.dtype == np.float32 and np.max(np.abs(got32 - gold[f"avg{k}_out"])) <= 2e-6


@pytest.mark.parametrize("k", range(len(EDGE_CASES)))
def test_conv_edge_detection_matches_reference_outputs(gold, k):
    import aegolius_b200 as ab
    got = ab.conv_edge_detection(gold[f"edge{k}_in"])
    assert np.max(np.abs(got - gold[f"edge{k}_out"])) <= 1e-12


def test_conv_averaging_larger_field_against_oracle():
    import aegolius_b200 as ab
    rng = np.random.default_rng(5)
    u = rng.normal(size=(65, 40, 33))
    for ks, it in (((5, 5, 1), 1), ((3, 4, 6), 2), (9, 1)):
        assert np.max(np.abs(ab.conv_averaging(u, ks, it) - F.conv_averaging(u, ks, it))) <= 1e-13
    assert ab.conv_averaging(u, 3, 0) is u  # the reference returns its argument untouched
    with pytest.raises(ValueError):
        ab.conv_averaging(u, (3, 3), 1)


@pytest.mark.parametrize("name", ["example", "everything"])
def test_vector_modifier_pipeline_matches_reference_outputs(gold, name):
    import aegolius_b200 as ab
    inp = vector_inputs()
    vf = ab.VectorFieldFromSDF(RES)
    for op, *a in pipelines(inp)[name]:
        getattr(vf, op)(*a)
    got = vf.create(inp["sdf"])
    assert got.shape == (3, int(np.prod(RES)))
    assert np.max(np.abs(got - gold[f"vec_{name}"])) <= 1e-12
    comps = vf.components(inp["sdf"])
    for c in ("x", "y", "z", "phi", "theta", "length"):
        ref = gold[f"vec_{name}_{c}"]
        ok = np.isfinite(ref)
        assert np.array_equal(ok, np.isfinite(comps[c])), c
        # acos is ill-conditioned at |z| -> 1: compare through the cosine there
        if c == "theta":
            assert np.max(np.abs(np.cos(comps[c][ok]) - np.cos(ref[ok]))) <= 1e-12
        else:
            assert np.max(np.abs(comps[c][ok] - ref[ok])) <= 1e-11, c
    assert np.array_equal(vf.x(inp["sdf"]), comps["x"])
    got32 = vf.create(inp["sdf"].astype(np.float32))
    assert got32.dtype == np.float32
    # fp32 central differences of an O(10) field carry ~1e-6 relative error before normalisation
    assert np.max(np.abs(got32 - gold[f"vec_{name}"])) <= 2e-4


def test_plain_vector_field_and_zero_vectors():
    import aegolius_b200 as ab
    inp = vector_inputs()
    got = ab.VectorFieldFromSDF(RES).create(inp["sdf"])
    gold_ = np.load(os.path.join(os.path.dirname(__file__), "golden", "fields.npz"))["vec_plain"]
    assert np.max(np.abs(got - gold_)) <= 1e-13
    assert np.any(np.all(got == 0, axis=0))  # flat planes of the input give zero vectors, kept zero by normalize


def test_vector_ops_reject_bad_operands():
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    inp = vector_inputs()
    vf = ab.VectorFieldFromSDF(RES)
    vf.rotate_z(np.zeros(7))
    with pytest.raises(ValueError):
        vf.create(inp["sdf"])
    import ctypes as C
    op = (cabi.ab_vec_op * 1)()
    op[0].opcode = 99
    d = C.c_void_p()
    cabi.check(cabi.lib().ab_device_alloc(3 * 64 * 8, 0, C.byref(d)))
    try:
        rc = cabi.lib().ab_vec_apply(d, 64, 64, op, 1, cabi.AB_F64, 0, None)
        assert rc == -2 and b"unknown opcode" in cabi.lib().ab_last_error()
        op[0].opcode = cabi.AB_VOP_REVOLVE_X  # needs a coordinate array
        rc = cabi.lib().ab_vec_apply(d, 64, 64, op, 1, cabi.AB_F64, 0, None)
        assert rc == -1
    finally:
        cabi.lib().ab_device_free(d, 0)
