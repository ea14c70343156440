"""GPU parity: the CUDA interpreter (through the C ABI) against (a) outputs of the reference itself (golden
fixtures) and (b) the oracle on the same inputs. Tolerances are BASELINE.json's: fp32 |d| <= 1e-5 * extent and
identical sign wherever |ref| > 1e-6 * extent; fp64 <= 1e-12 * extent."""
import numpy as np
import pytest

from conftest import golden_names, load_case
from oracle import interp_np

pytestmark = pytest.mark.gpu

F32_TOL, F32_SIGN_BAND, F64_TOL = 1e-5, 1e-6, 1e-12


def _spec(c):
    from aegolius_b200.grid import GridSpec
    dims3 = c["res"][2] > 1
    return GridSpec(c["size"][:3] if dims3 else c["size"][:2], c["res"][:3] if dims3 else c["res"][:2])


def _compare(got, exp, margin, extent, tol, band, name):
    assert got.shape == exp.shape
    keep = margin > band * extent
    assert keep.mean() > 0.97, f"{name}: too many points on branch boundaries"
    g, e = got[keep].astype(np.float64), exp[keep]
    assert np.array_equal(np.isnan(g), np.isnan(e)), f"{name}: NaN pattern differs"
    ok = ~np.isnan(e)
    err = np.max(np.abs(g[ok] - e[ok])) if ok.any() else 0.0
    assert err <= tol * extent, f"{name}: max |cuda - reference| = {err:.3e} > {tol * extent:.1e}"
    far = ok & (np.abs(e) > F32_SIGN_BAND * extent)
    assert np.array_equal(np.sign(g[far]), np.sign(e[far])), f"{name}: inside/outside mask differs"
    return err


@pytest.mark.parametrize("name", golden_names())
def test_grid_f32_matches_reference(golden, name):
    import aegolius_b200 as ab
    c = load_case(golden, name)
    _, margin = interp_np.run_grid(c["prog"], c["size"], c["res"], return_margin=True)
    got = ab.create(c["prog"], _spec(c), dtype="f32")
    assert got.dtype == np.float32
    _compare(got, c["expected"], margin, c["extent"], F32_TOL, 2e-6, name)


@pytest.mark.parametrize("name", golden_names())
def test_grid_f64_matches_reference(golden, name):
    import aegolius_b200 as ab
    c = load_case(golden, name)
    _, margin = interp_np.run_grid(c["prog"], c["size"], c["res"], return_margin=True)
    got = ab.create(c["prog"], _spec(c), dtype="f64")
    assert got.dtype == np.float64
    _compare(got, c["expected"], margin, c["extent"], F64_TOL, 1e-9, name)


@pytest.mark.parametrize("name", ["c1_sphere_box_smooth_union", "c3_deep_tree", "struct_deep_combines",
                                  "c2_composite_2d_all13"])
def test_points_mode_matches_oracle(golden, name):
    import aegolius_b200 as ab
    c = load_case(golden, name)
    rng = np.random.default_rng(7)
    n = 10007  # ragged: not a multiple of the tile
    co = rng.uniform(-0.5, 0.5, size=(3, n)) * np.asarray(c["size"]).reshape(3, 1)
    exp, margin = interp_np.run(c["prog"], co, return_margin=True)
    got64 = ab.create(c["prog"], co, dtype="f64")
    _compare(got64, exp, margin, c["extent"], F64_TOL, 1e-9, name)
    got32 = ab.create(c["prog"], co, dtype="f32")
    _compare(got32, exp, margin, c["extent"], F32_TOL, 2e-6, name)


def test_slabs_concatenate_bit_identically(golden):
    import aegolius_b200 as ab
    c = load_case(golden, "c3_deep_tree")
    spec = _spec(c)
    whole = ab.create(c["prog"], spec, dtype="f32")
    for parts in (2, 3, 8):
        pieces = [ab.create(c["prog"], spec, dtype="f32", slab=s) for s in ab.engine.slab_ranges(spec.res[0], parts)]
        assert np.array_equal(np.concatenate(pieces), whole)


def test_plain_spomso_style_array_is_detected_as_grid(golden):
    import aegolius_b200 as ab
    c = load_case(golden, "c1_sphere_box_smooth_union")
    spec = _spec(c)
    co = np.array(spec.materialize())  # plain ndarray, as SPOMSO's generate_grid returns
    assert ab.detect_grid(co) is not None
    a = ab.create(c["prog"], co, dtype="f64")
    b = ab.create(c["prog"], spec, dtype="f64")
    assert np.array_equal(a, b)
    # a perturbed array is NOT a grid and goes through points mode, same values to rounding
    co2 = co + 0.0
    co2[0, 5] += 1e-3
    assert ab.detect_grid(co2) is None
    d = ab.create(c["prog"], co2, dtype="f64")
    assert np.max(np.abs(np.delete(d, 5) - np.delete(a, 5))) < 1e-12


@pytest.mark.parametrize("name", ["c1_sphere_box_smooth_union", "c3_deep_tree", "mod_twist", "mod_bend",
                                  "prim3_torus", "comb_SMOOTH_INTERSECT2_BOLTZMANN", "struct_extruded_combo"])
def test_spatial_gradient_matches_central_differences_of_oracle(golden, name):
    """The dual-number gradient against fp64 central differences of the oracle itself, away from kinks."""
    import aegolius_b200 as ab
    c = load_case(golden, name)
    rng = np.random.default_rng(11)
    n = 4001
    co = rng.uniform(-0.45, 0.45, size=(3, n)) * np.asarray(c["size"]).reshape(3, 1)
    h = 1e-6 * c["extent"]
    fd = np.empty((3, n))
    kink = np.zeros(n, dtype=bool)
    f0, margin = interp_np.run(c["prog"], co, return_margin=True)
    for k in range(3):
        e = np.zeros((3, 1))
        e[k] = h
        fp, fm = interp_np.run(c["prog"], co + e), interp_np.run(c["prog"], co - e)
        fd[k] = (fp - fm) / (2 * h)
        kink |= np.abs((fp - f0) / h - (f0 - fm) / h) > 1e-3  # one-sided slopes disagree: a kink
    val, grad = ab.create(c["prog"], co, dtype="f64", grad="spatial")
    assert np.max(np.abs(val - f0)[margin > 1e-9]) <= F64_TOL * c["extent"]
    ok = ~kink & (margin > 1e-4 * c["extent"])
    assert ok.mean() > 0.8
    assert np.max(np.abs(grad[:, ok] - fd[:, ok])) < 2e-5
    val32, grad32 = ab.create(c["prog"], co, dtype="f32", grad="spatial")
    assert np.max(np.abs(grad32[:, ok] - fd[:, ok])) < 5e-3
    assert np.max(np.abs(val32 - f0)[margin > 2e-6 * c["extent"]]) <= F32_TOL * c["extent"]
