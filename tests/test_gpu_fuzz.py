"""Differential fuzzing of the flattener + CUDA interpreter against the oracle: seeded random trees (random primitives,
Euclidean transforms, modifications and the 13 combine ops, nesting depth <= 3) on a small grid. The golden scenarios pin
every op on its own; this pins their COMBINATIONS (stack slots, affine folding, superinstruction fusion, tier choice)."""
import numpy as np
import pytest

from oracle import interp_np

pytestmark = pytest.mark.gpu

NONPARAM = ["UNION2", "UNION", "SUBTRACT2", "INTERSECT2", "INTERSECT", "SUM", "DIFFERENCE"]
PARAM = ["SMOOTH_UNION2_2", "SMOOTH_UNION2", "SMOOTH_INTERSECT2", "SMOOTH_INTERSECT2_BOLTZMANN", "SMOOTH_SUBTRACT2",
         "SMOOTH_SUBTRACT2_BOLTZMANN"]


def _leaf(ab, rng):
    u = rng.uniform
    k = rng.integers(0, 12)
    if k == 0:
        return ab.Sphere(u(0.3, 1.0))
    if k == 1:
        return ab.Box(u(0.4, 1.6), u(0.4, 1.6), u(0.4, 1.6))
    if k == 2:
        return ab.Cylinder(u(0.2, 0.8), u(0.4, 1.8))
    if k == 3:
        return ab.Torus(u(0.5, 1.0), u(0.1, 0.3))
    if k == 4:
        return ab.Cone(u(0.6, 1.5), u(0.3, 1.0))
    if k == 5:
        return ab.ChainLink(u(0.3, 0.6), u(0.05, 0.15), u(0.4, 1.0))
    if k == 6:
        return ab.Line(tuple(u(-1, 1, 3)), tuple(u(-1, 1, 3)))
    if k == 7:
        return ab.Triangle3D(tuple(u(-1, 1, 3)), tuple(u(-1, 1, 3)), tuple(u(-1, 1, 3)))
    if k == 8:
        return ab.SolidAngle(u(0.5, 1.2), u(0.2, 0.8), u(1.0, 2.0))
    if k == 9:
        o = ab.NGon(u(0.4, 0.9), int(rng.integers(3, 8)))
        o.extrusion(u(0.3, 1.2))
        return o
    if k == 10:
        o = ab.Rectangle(u(0.3, 0.8), u(0.2, 0.6))
        o.revolution(u(0.5, 1.0))
        return o
    return ab.Arc3D(u(0.5, 1.0), u(0.1, 0.25), u(0.2, 1.2), u(1.5, 2.8))


def _modify(o, rng):
    u = rng.uniform
    for _ in range(int(rng.integers(0, 3))):
        k = rng.integers(0, 12)
        if k == 0:
            o.elongation(tuple(u(0.0, 0.6, 3)))
        elif k == 1:
            o.rounding(u(0.02, 0.15))
        elif k == 2:
            o.onion(u(0.03, 0.12))
        elif k == 3:
            o.twist(u(-1.5, 1.5))
        elif k == 4:
            o.symmetry(int(rng.integers(0, 3)))
        elif k == 5:
            o.mirror(tuple(u(-1, 0, 3)), tuple(u(0.2, 1, 3)))
        elif k == 6:
            o.rotational_symmetry(int(rng.integers(3, 8)), u(0.6, 1.4), u(0.0, 0.5))
        elif k == 7:
            o.infinite_repetition(tuple(u(1.5, 2.5, 3)))
        elif k == 8:
            o.scale_sdf(u(0.6, 1.6))
        elif k == 9:
            o.shear_xz(u(-0.4, 0.4))
        elif k == 10:
            o.bend(u(1.0, 2.0), u(0.5, 1.5))
        else:
            o.concentric(u(0.1, 0.3))
    return o


def _transform(o, rng):
    u = rng.uniform
    if rng.random() < 0.7:
        o.rotate(u(-3, 3), tuple(u(-1, 1, 3) + np.array([0, 0, 1.5])))
    if rng.random() < 0.5:
        o.rescale(u(0.6, 1.5))
    if rng.random() < 0.8:
        o.move(tuple(u(-0.8, 0.8, 3)))
    return o


def _tree(ab, rng, depth):
    if depth == 0 or rng.random() < 0.25:
        return _transform(_modify(_leaf(ab, rng), rng), rng)
    if rng.random() < 0.55:
        op = NONPARAM[rng.integers(0, len(NONPARAM))]
        n = int(rng.integers(2, 5)) if op in ("UNION", "INTERSECT") else 2
        node = ab.CombineGeometry(op).combine(*[_tree(ab, rng, depth - 1) for _ in range(n)])
    else:
        op = PARAM[rng.integers(0, len(PARAM))]
        node = ab.CombineGeometry(op).combine_parametric(_tree(ab, rng, depth - 1), _tree(ab, rng, depth - 1),
                                                         parameters=float(rng.uniform(0.15, 0.5)))
    if rng.random() < 0.3:  # nesting through propagate, as basics_3D.py:111 does
        node = ab.GenericGeometry(node.propagate, ())
    return _transform(_modify(node, rng), rng)


@pytest.mark.parametrize("seed", range(60))
def test_random_tree_matches_oracle(seed):
    import aegolius_b200 as ab
    rng = np.random.default_rng(1000 + seed)
    tree = _tree(ab, rng, 3)
    prog = ab.flatten(tree)
    spec = ab.GridSpec((5.0, 5.0, 5.0), (20, 18, 22))
    exp, margin = interp_np.run_grid(prog, spec.size, spec.res, return_margin=True)
    ok = np.isfinite(exp)
    assert ok.mean() > 0.99
    scale = max(1.0, float(np.max(np.abs(exp[ok]))))  # SUM / repeated scaling can grow the field beyond the extent
    for dt, tol, band in (("f64", 1e-11, 1e-9), ("f32", 3e-5, 2e-6)):
        got = ab.create(prog, spec, dtype=dt).astype(np.float64)
        keep = ok & (margin > band * 5.0)
        assert keep.mean() > 0.9, f"seed {seed}: {1 - keep.mean():.2%} of the points sit on branch boundaries"
        err = np.max(np.abs(got[keep] - exp[keep]))
        assert err <= tol * 5.0 * scale, f"seed {seed} {dt}: max |cuda - oracle| = {err:.3e}\\n{prog.disassemble()}"
    # the optimiser must not change the result: unfused / unfolded program against the optimised one
    raw = ab.flatten(tree, optimize=False)
    a = ab.create(raw, spec, dtype="f64")
    b = ab.create(prog, spec, dtype="f64")
    keep = ok & (margin > 1e-9 * 5.0)
    assert np.max(np.abs(a[keep] - b[keep])) <= 1e-11 * 5.0 * scale, f"seed {seed}: optimised program diverges"


@pytest.mark.parametrize("seed", range(20))
def test_random_tree_gradient_and_points_mode(seed):
    """Same trees at scattered points: points mode against the oracle, and the dual-number gradient against fp64 central
    differences of the oracle away from kinks and branch boundaries."""
    import aegolius_b200 as ab
    rng = np.random.default_rng(1000 + seed)
    prog = ab.flatten(_tree(ab, rng, 3))
    n = 3001
    co = np.random.default_rng(seed).uniform(-2.4, 2.4, size=(3, n))
    f0, margin = interp_np.run(prog, co, return_margin=True)
    scale = max(1.0, float(np.max(np.abs(f0[np.isfinite(f0)]))))
    h = 5e-6
    fd = np.empty((3, n))
    kink = ~np.isfinite(f0)
    for k in range(3):
        e = np.zeros((3, 1))
        e[k] = h
        fp, fm = interp_np.run(prog, co + e), interp_np.run(prog, co - e)
        fd[k] = (fp - fm) / (2 * h)
        kink |= ~(np.abs((fp - f0) / h - (f0 - fm) / h) <= 1e-3 * scale)
    val, grad = ab.create(prog, co, dtype="f64", grad="spatial")
    keep = np.isfinite(f0) & (margin > 1e-9 * 5.0)
    assert np.max(np.abs(val - f0)[keep]) <= 1e-11 * 5.0 * scale, f"seed {seed}: points mode diverges"
    ok = ~kink & (margin > 1e-4 * 5.0)
    assert ok.mean() > 0.6
    err = np.max(np.abs(grad[:, ok] - fd[:, ok]))
    assert err < 1e-4 * scale, f"seed {seed}: gradient off by {err:.2e}\\n{prog.disassemble()}"


def _leaf2(ab, rng):
    u = rng.uniform
    k = rng.integers(0, 9)
    if k == 0:
        return ab.Circle(u(0.3, 1.2))
    if k == 1:
        return ab.Rectangle(u(0.5, 2.0), u(0.4, 1.5))
    if k == 2:
        return ab.RoundedRectangle(u(1.0, 2.0), u(0.8, 1.5), tuple(u(0.05, 0.3, 4)))
    if k == 3:
        return ab.NGon(u(0.5, 1.2), int(rng.integers(3, 9)))
    if k == 4:
        return ab.Triangle(tuple(u(-1.2, 1.2, 2)), tuple(u(-1.2, 1.2, 2)), tuple(u(-1.2, 1.2, 2)))
    if k == 5:
        return ab.Sector(u(0.6, 1.5), u(0.1, 0.8), u(1.0, 2.6))
    if k == 6:
        return ab.Arc(u(0.6, 1.3), u(0.1, 0.3), u(1.2, 2.8))
    if k == 7:
        return ab.Segment(tuple(u(-1.5, 1.5, 2)), tuple(u(-1.5, 1.5, 2)))
    return ab.NEUCircle(u(0.5, 1.2), float(rng.choice([1.0, 3.0, 4.0])))


def _tree2(ab, rng, depth):
    u = rng.uniform

    def place(o):
        for _ in range(int(rng.integers(0, 2))):
            j = rng.integers(0, 4)
            if j == 0:
                o.rounding(u(0.02, 0.15))
            elif j == 1:
                o.onion(u(0.03, 0.1))
            elif j == 2:
                o.symmetry(int(rng.integers(0, 2)))
            else:
                o.boundary()
        o.rotate(u(0, 2 * np.pi), (0, 0, 1))
        if rng.random() < 0.5:
            o.rescale(u(0.6, 1.6))
        o.move((u(-2.0, 2.0), u(-2.0, 2.0), 0.0))
        return o
    if depth == 0 or rng.random() < 0.2:
        return place(_leaf2(ab, rng))
    if rng.random() < 0.5:
        op = NONPARAM[rng.integers(0, len(NONPARAM))]
        n = int(rng.integers(2, 6)) if op in ("UNION", "INTERSECT") else 2
        node = ab.CombineGeometry(op).combine(*[_tree2(ab, rng, depth - 1) for _ in range(n)])
    else:
        op = PARAM[rng.integers(0, len(PARAM))]
        node = ab.CombineGeometry(op).combine_parametric(_tree2(ab, rng, depth - 1), _tree2(ab, rng, depth - 1),
                                                         parameters=float(u(0.15, 0.5)))
    return place(node) if rng.random() < 0.5 else node


@pytest.mark.parametrize("seed", range(30))
def test_random_2d_tree_matches_oracle(seed):
    import aegolius_b200 as ab
    rng = np.random.default_rng(5000 + seed)
    prog = ab.flatten(_tree2(ab, rng, 3))
    spec = ab.GridSpec((8.0, 8.0), (70, 58))
    exp, margin = interp_np.run_grid(prog, spec.size + (0.0,), spec.res, return_margin=True)
    ok = np.isfinite(exp)
    assert ok.mean() > 0.99
    scale = max(1.0, float(np.max(np.abs(exp[ok]))))
    for dt, tol, band in (("f64", 1e-11, 1e-9), ("f32", 3e-5, 2e-6)):
        got = ab.create(prog, spec, dtype=dt).astype(np.float64)
        keep = ok & (margin > band * 8.0)
        assert keep.mean() > 0.9
        err = np.max(np.abs(got[keep] - exp[keep]))
        assert err <= tol * 8.0 * scale, f"seed {seed} {dt}: max |cuda - oracle| = {err:.3e}\\n{prog.disassemble()}"
