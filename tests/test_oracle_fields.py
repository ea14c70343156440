"""The stencil / vector-field oracle (oracle/fields_np.py) against outputs of the reference itself
(tests/golden/fields.npz, made by tests/golden/make_golden_fields.py)."""
import os

import numpy as np
import pytest

from oracle import fields_np as F
from field_cases import AVG_CASES, EDGE_CASES, RES, vector_inputs, pipelines


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "fields.npz"))


@pytest.mark.parametrize("k", range(len(AVG_CASES)))
def test_conv_averaging_matches_reference(gold, k):
    shape, ks, it = AVG_CASES[k]
    got = F.conv_averaging(gold[f"avg{k}_in"], ks, it)
    assert got.shape == tuple(shape)
    assert np.max(np.abs(got - gold[f"avg{k}_out"])) <= 1e-14


@pytest.mark.parametrize("k", range(len(EDGE_CASES)))
def test_conv_edge_detection_matches_reference(gold, k):
    got = F.conv_edge_detection(gold[f"edge{k}_in"])
    assert np.max(np.abs(got - gold[f"edge{k}_out"])) <= 1e-13


def test_conv_averaging_argument_handling():
    u = np.arange(12.0).reshape(3, 4)
    assert F.conv_averaging(u, 3, 0) is not None and np.array_equal(F.conv_averaging(u, 3, 0), u)
    with pytest.raises(ValueError):
        F.conv_averaging(u, (3, 3, 3), 1)


@pytest.mark.parametrize("name", ["example", "everything"])
def test_vector_modifier_pipeline_matches_reference(gold, name):
    from oracle import interp_np
    inp = vector_inputs()
    base = interp_np.from_sdf(inp["sdf"], RES)
    assert np.max(np.abs(base - gold["vec_plain"])) <= 1e-14
    got = F.apply_ops(base, pipelines(inp)[name])
    assert np.max(np.abs(got - gold[f"vec_{name}"])) <= 1e-13
    comps = F.components(got)
    for c in ("x", "y", "z", "phi", "theta", "length"):
        ref = gold[f"vec_{name}_{c}"]
        ok = np.isfinite(ref)
        assert np.array_equal(ok, np.isfinite(comps[c]))
        assert np.max(np.abs(comps[c][ok] - ref[ok])) <= 1e-12
