"""GPU parity for the two non-interpreter kernels (point cloud -> SDF, from_sdf) and the C-ABI error paths."""
import ctypes as C

import numpy as np
import pytest

from oracle import interp_np

pytestmark = pytest.mark.gpu


def _cloud(m, seed=0):
    rng = np.random.default_rng(seed)
    x, y = rng.uniform(-1, 1, m), rng.uniform(-1, 1, m)
    z = 0.2 * np.sin(3 * x) * np.cos(2 * y) + 0.01 * rng.normal(size=m)
    return np.stack([x, y, z])


@pytest.mark.parametrize("m", [1, 37, 1024, 5000])
def test_point_cloud_grid_matches_exhaustive_oracle(m):
    import aegolius_b200 as ab
    pts = _cloud(m)
    spec = ab.GridSpec((2.5, 2.5, 1.5), (20, 16, 12))
    exp = interp_np.point_cloud_distance(spec.materialize(), pts)
    got64 = ab.point_cloud_sdf(spec, pts, dtype="f64")
    assert np.max(np.abs(got64 - exp)) <= 1e-12 * 2.5
    got32 = ab.point_cloud_sdf(spec, pts, dtype="f32")
    assert np.max(np.abs(got32 - exp)) <= 1e-5 * 2.5


def test_point_cloud_2d_and_points_mode():
    import aegolius_b200 as ab
    pts = _cloud(777, seed=3)
    rng = np.random.default_rng(5)
    co = rng.uniform(-1.2, 1.2, size=(3, 3001))
    exp3 = interp_np.point_cloud_distance(co, pts, dim=3)
    exp2 = interp_np.point_cloud_distance(co, pts, dim=2)
    assert np.max(np.abs(ab.point_cloud_sdf(co, pts, dim=3, dtype="f64") - exp3)) <= 1e-12 * 2.4
    assert np.max(np.abs(ab.point_cloud_sdf(co, pts, dim=2, dtype="f64") - exp2)) <= 1e-12 * 2.4
    assert np.max(np.abs(ab.point_cloud_sdf(co, pts, dim=3, dtype="f32") - exp3)) <= 1e-5 * 2.4


@pytest.mark.parametrize("m", [2000, 60000])
def test_point_cloud_leaf_inside_a_tree_matches_oracle(m):
    """PointCloud3D(points).onion(t) under a transform, the pointcloud_terrain_3D.py:39-47 pattern. Clouds of >= 2048
    points are searched through the octree inside the interpreter, smaller ones are scanned."""
    import aegolius_b200 as ab
    pts = _cloud(m, seed=9)
    pc = ab.PointCloud3D(pts)
    pc.onion(0.03)
    pc.rotate(0.3, (0, 0, 1))
    pc.move((0.1, 0.0, 0.05))
    prog = ab.flatten(pc)
    spec = ab.GridSpec((2.5, 2.5, 1.5), (16, 16, 12))
    exp = interp_np.run_grid(prog, spec.size, spec.res)
    assert np.max(np.abs(ab.create(prog, spec, dtype="f64") - exp)) <= 1e-12 * 2.5
    assert np.max(np.abs(ab.create(prog, spec, dtype="f32") - exp)) <= 1e-5 * 2.5
    # analytic gradient of the leaf: unit vector from the nearest point (away from ties), through the rotation
    f, g = ab.create(prog, spec, dtype="f64", grad="spatial")
    assert np.max(np.abs(f - exp)) <= 1e-12 * 2.5
    nrm = np.sqrt((g * g).sum(axis=0))
    assert np.max(np.abs(nrm - 1.0)) <= 1e-9


def test_bare_point_cloud_object_takes_the_dedicated_path():
    """PointCloud3D(points).create(grid) with no transform or modification is routed to ab_nn_grid (packet walk);
    same field as the interpreter leaf to fp32 rounding, and identical to point_cloud_sdf."""
    import aegolius_b200 as ab
    pts = _cloud(30000, seed=2)
    spec = ab.GridSpec((2.5, 2.5, 1.5), (24, 20, 16))
    bare = ab.PointCloud3D(pts)
    assert ab.flatten(bare).n_ops == 2
    a = ab.create(bare, spec, dtype="f32")
    assert np.array_equal(a, ab.point_cloud_sdf(spec, pts, dtype="f32"))
    moved = ab.PointCloud3D(pts)
    moved.move((0.0, 0.0, 1e-300))  # a no-op translation keeps the interpreter path (3 ops)
    b = ab.create(moved, spec, dtype="f32")
    assert np.max(np.abs(a - b)) <= 1e-6 * 2.5
    exp = interp_np.point_cloud_distance_kdtree(spec.materialize(), pts)
    assert np.max(np.abs(a - exp)) <= 1e-5 * 2.5


def test_point_cloud_2d_leaf_with_octree_matches_oracle():
    import aegolius_b200 as ab
    rng = np.random.default_rng(3)
    t = rng.uniform(0, 2 * np.pi, 30000)
    pts = np.stack([np.cos(t) * (1 + 0.2 * np.sin(5 * t)), np.sin(t) * (1 + 0.2 * np.sin(5 * t))])
    pc = ab.PointCloud2D(pts)
    pc.rescale(0.8)
    pc.move((0.2, -0.1, 0.0))
    u = ab.CombineGeometry("UNION2").combine(pc, ab.Circle(0.3))
    prog = ab.flatten(u)
    spec = ab.GridSpec((4, 4), (96, 80))
    exp = interp_np.run_grid(prog, spec.size, spec.res)
    assert np.max(np.abs(ab.create(prog, spec, dtype="f64") - exp)) <= 1e-12 * 4
    assert np.max(np.abs(ab.create(prog, spec, dtype="f32") - exp)) <= 1e-5 * 4


@pytest.mark.parametrize("res", [(9, 7, 11), (13, 5), (2, 2, 2), (33, 2, 3)])
def test_from_sdf_matches_np_gradient_semantics(res):
    import aegolius_b200 as ab
    rng = np.random.default_rng(2)
    f = rng.normal(size=int(np.prod(res)))
    f[:3] = 0.0
    exp = interp_np.from_sdf(f, res)
    got = ab.from_sdf(f, res)
    assert got.shape == exp.shape and got.dtype == np.float64
    assert np.max(np.abs(got - exp)) < 1e-13
    got32 = ab.from_sdf(f.astype(np.float32), res)
    assert got32.dtype == np.float32
    assert np.max(np.abs(got32 - exp)) < 5e-5  # fp32 field: differences of O(1) values
    flat = np.zeros(int(np.prod(res)))
    assert not np.any(ab.from_sdf(flat, res))  # zero vectors stay zero (batch_normalize mask)


def test_c_abi_rejects_bad_programs_with_messages():
    from aegolius_b200 import cabi, flatten, Sphere, GridSpec
    import aegolius_b200 as ab
    prog = flatten(Sphere(1.0))
    bad = ab.Program(prog.ops.copy(), prog.args.copy(), [], 1, 1)
    bad.ops["opcode"][0] = 159  # unknown opcode
    with pytest.raises(cabi.AegoliusError) as e:
        ab.create(bad, GridSpec((2, 2, 2), (4, 4, 4)))
    assert e.value.code == cabi.AB_EUNSUPPORTED_OP and "opcode" in str(e.value)
    bad2 = ab.Program(prog.ops.copy(), prog.args[:0].copy(), [], 1, 1)
    with pytest.raises(cabi.AegoliusError) as e:
        ab.create(bad2, GridSpec((2, 2, 2), (4, 4, 4)))
    assert e.value.code == cabi.AB_EINVAL
    with pytest.raises(cabi.AegoliusError):
        ab.create(prog, GridSpec((2, 2, 2), (4, 4, 4)), device=99)
    # empty input: no launch, empty result
    out = ab.create(prog, np.zeros((3, 0)))
    assert out.shape == (0,)


def test_launch_counter_counts_kernels():
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    n0 = cabi.launch_count()
    ab.create(ab.flatten(ab.Sphere(1.0)), ab.GridSpec((2, 2, 2), (8, 8, 8)))
    assert cabi.launch_count() == n0 + 1


def test_parameter_gradient_map_matches_oracle_differences():
    """The gradient_map_3D.py:55-84 geometry (arc, concentric, elongation, two rotated copies, union, smooth union, onion)
    differentiated with respect to each of its four parameters."""
    import aegolius_b200 as ab

    def geometry(r, a, w, s):
        def part(sign):
            arc = ab.Arc3D(r, 0.0, np.pi * 5 / 6, -np.pi * 5 / 6)
            arc.concentric(w)
            arc.elongation((0.0, 0.0, 0.75 / 2))
            if sign:
                arc.rotate(float(np.deg2rad(sign * a)), (0, 0, 1))
                arc.rescale(1.2)
                arc.move((0, 0, 1.5 * sign))
            return arc
        u = ab.CombineGeometry("UNION2").combine(part(+1), part(-1))
        v = ab.CombineGeometry("SMOOTH_UNION2").combine_parametric(u, part(0), parameters=s)
        v.onion(0.1)
        return v

    p = (2.0, 30.0, 0.5, 1.8)
    rng = np.random.default_rng(4)
    co = rng.uniform(-3, 3, size=(3, 5003))
    for k in range(4):
        got = ab.jacfwd(geometry, argnums=k, mode="fd")(co, *p)
        dual64 = ab.jacfwd(geometry, argnums=k, mode="dual", dtype="f64")(co, *p)
        dual32 = ab.jacfwd(geometry, argnums=k, mode="dual", dtype="f32")(co, *p)
        h = 1e-5
        lo, hi = list(p), list(p)
        lo[k] -= h
        hi[k] += h
        exp = (interp_np.run(ab.flatten(geometry(*hi)), co) - interp_np.run(ab.flatten(geometry(*lo)), co)) / (2 * h)
        kink = np.abs(exp - got) > 1e-3  # points whose active branch flips inside the stencil
        assert kink.mean() < 0.02
        assert np.max(np.abs(exp - got)[~kink]) < 1e-5
        # forward-mode dual numbers through the interpreter (AB_GRAD_PARAM) agree with the differences
        kd = np.abs(exp - dual64) > 1e-3
        assert kd.mean() < 0.02
        assert np.max(np.abs(exp - dual64)[~kd]) < 1e-5
        assert np.max(np.abs(dual32 - dual64)[~kd]) < 2e-3


def test_parameter_tangent_through_a_table_is_refused():
    import aegolius_b200 as ab
    from aegolius_b200 import cabi

    def geometry(r):
        b = ab.Box(0.3, 0.2, 0.2)
        b.curve_instancing(lambda t, rr: np.asarray((rr * np.cos(2 * np.pi * t), rr * np.sin(2 * np.pi * t), 0 * t)),
                           (r,), (0, 0.9, 7))
        return b
    co = np.random.default_rng(1).uniform(-1, 1, size=(3, 100))
    with pytest.raises(cabi.AegoliusError) as e:
        ab.jacfwd(geometry, mode="dual")(co, 1.0)
    assert e.value.code == cabi.AB_EUNSUPPORTED_OP and "table" in str(e.value)
    assert np.isfinite(ab.jacfwd(geometry, mode="fd")(co, 1.0)).all()  # the difference mode handles every op


def test_point_cloud_few_queries_many_points_split_path():
    """Short query lists take the split-cloud kernel (warp-shuffle min + atomic min); ragged sizes on purpose."""
    import aegolius_b200 as ab
    pts = _cloud(70_001, seed=21)
    rng = np.random.default_rng(8)
    for n in (1, 17, 1003):
        co = rng.uniform(-1.3, 1.3, size=(3, n))
        exp = interp_np.point_cloud_distance(co, pts, dim=3)
        got = ab.point_cloud_sdf(co, pts, dim=3, dtype="f32")
        assert got.shape == (n,) and np.max(np.abs(got - exp)) <= 1e-5 * 2.6
        got2 = ab.point_cloud_sdf(co, pts, dim=2, dtype="f32")
        assert np.max(np.abs(got2 - interp_np.point_cloud_distance(co, pts, dim=2))) <= 1e-5 * 2.6


def test_position_optimisation_converges_like_the_jax_example():
    """position_optimization.py:60-179 without JAX/optax: a circle is moved onto a target circle by Adam on the
    least-squares field mismatch; value and gradient come from aegolius_b200.value_and_grad (AB_GRAD_PARAM passes +
    device reductions)."""
    import aegolius_b200 as ab
    spec = ab.GridSpec((4, 4), (128, 128))
    radius, target_xy = 0.5, (0.7, -0.4)

    def circle_at(x0, y0):
        c = ab.Circle(radius)
        c.move((x0, y0, 0.0))
        return c

    target = ab.create(circle_at(*target_xy), spec, dtype="f32")
    vg = ab.value_and_grad(circle_at, spec, target)
    params = np.array([-0.6, 0.5])
    loss0, g0 = vg(params)
    # gradient against a central difference of the loss itself
    h = 1e-3
    fd = np.array([(vg(params + h * e)[0] - vg(params - h * e)[0]) / (2 * h) for e in np.eye(2)])
    assert np.allclose(g0, fd, rtol=2e-2, atol=1e-2 * abs(fd).max())
    m, v, lr = np.zeros(2), np.zeros(2), 0.05
    for it in range(1, 200):
        loss, g = vg(params)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        params = params - lr * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
        if loss < 1e-3 * loss0:
            break
    assert loss < 1e-2 * loss0 and np.max(np.abs(params - np.array(target_xy))) < 0.05


def test_fused_loss_reduction_matches_unfused_and_oracle():
    """ab_eval_grid_loss (sums reduced inside the AB_GRAD_PARAM kernel) against the same sums formed from the stored
    (F, dF/dtheta) maps and against the fp64 oracle's field; 3D tree, both precisions."""
    import torch
    import aegolius_b200 as ab
    from aegolius_b200 import engine
    spec = ab.GridSpec((4, 4, 4), (48, 40, 36))

    def geometry(r, w, x0):
        s = ab.Sphere(r)
        s.move((x0, 0.1, -0.2))
        b = ab.Box(1.5, 1.0, 0.8)
        b.rotate(0.6, (0, 0, 1))
        return ab.CombineGeometry("SMOOTH_UNION2").combine_parametric(s, b, parameters=w)

    params = [0.9, 0.3, 0.4]
    rng = np.random.default_rng(4)
    target64 = interp_np.run_grid(ab.flatten(geometry(1.0, 0.25, 0.5)), spec.size, spec.res) + 0.01 * rng.normal(size=spec.n_points)
    exp_field = interp_np.run_grid(ab.flatten(geometry(*params)), spec.size, spec.res)
    exp_loss = float(np.sum((exp_field - target64) ** 2))
    for dt, rtol in (("f64", 1e-11), ("f32", 2e-5)):
        loss, grads = ab.value_and_grad(geometry, spec, target64, dtype=dt)(params)
        assert abs(loss - exp_loss) <= rtol * exp_loss
        # the unfused path: identity `post` makes value_and_grad store the maps and reduce them with torch
        loss_u, grads_u = ab.value_and_grad(geometry, spec, target64, dtype=dt, post=lambda f, d: (f, d))(params)
        assert abs(loss - loss_u) <= 10 * rtol * abs(loss_u)
        assert np.allclose(grads, grads_u, rtol=50 * rtol, atol=50 * rtol * np.max(np.abs(grads_u)))
    # slabs add up: two half-grid calls through the C ABI
    import ctypes as C
    from aegolius_b200 import cabi
    prog = engine.program_tangent(geometry, params, 0)
    cp = cabi.CProgram(prog)
    tgt = torch.as_tensor(target64, dtype=torch.float64, device="cuda")
    acc = torch.zeros(2, dtype=torch.float64, device="cuda")
    total = np.zeros(2)
    plane = spec.res[1] * spec.res[2]
    for x0, x1 in ((0, 20), (20, spec.res[0])):
        g = cabi.make_grid(spec.size, spec.res, (x0, x1))
        cabi.check(cabi.lib().ab_eval_grid_loss(cp.ref(), C.byref(g), cabi.AB_F64, tgt.data_ptr() + 8 * x0 * plane,
                                                acc.data_ptr(), 0, None))
        total += acc.cpu().numpy()
    full, gfull = ab.value_and_grad(geometry, spec, target64, dtype="f64")(params)
    assert abs(total[0] - full) <= 1e-11 * full and abs(total[1] - gfull[0]) <= 1e-9 * abs(gfull[0])


def _cloud_cases():
    rng = np.random.default_rng(1)
    from aegolius_b200 import workloads
    return {
        "uniform": rng.uniform(-1, 1, (3, 20000)),
        "surface": workloads.c4_cloud(50000, 3),
        "clusters": np.concatenate([rng.normal(0, 0.01, (3, 5000)), rng.normal(1.5, 0.3, (3, 5000))], axis=1),
        "single": np.array([[0.3], [0.2], [-0.1]]),
        "duplicates": np.repeat(rng.uniform(-1, 1, (3, 10)), 300, axis=1),
        "collinear": np.stack([np.linspace(-1, 1, 3000), np.zeros(3000), np.zeros(3000)]),
        "far_from_origin": rng.uniform(100, 100.5, (3, 4000)),
    }


@pytest.mark.parametrize("name", ["uniform", "surface", "clusters", "single", "duplicates", "collinear", "far_from_origin"])
def test_point_cloud_octree_and_brute_force_return_identical_bits(name, monkeypatch):
    """The octree path prunes, it never approximates: same bits as the exhaustive kernel (grid + point-list queries,
    2D + 3D, both precisions), including degenerate clouds."""
    import aegolius_b200 as ab
    pts = _cloud_cases()[name]
    rng = np.random.default_rng(7)
    co = rng.uniform(-3, 3, (3, 20011))
    for dt in ("f32", "f64"):
        for dim, target in ((3, ab.GridSpec((4, 4, 4), (40, 38, 42))), (2, ab.GridSpec((4, 4), (150, 131))), (3, co), (2, co)):
            monkeypatch.setenv("AB_NN_ALGO", "brute")
            a = ab.point_cloud_sdf(target, pts, dim=dim, dtype=dt)
            monkeypatch.setenv("AB_NN_ALGO", "tree")
            b = ab.point_cloud_sdf(target, pts, dim=dim, dtype=dt)
            assert np.array_equal(a, b), (name, dt, dim, float(np.max(np.abs(a - b))))


def test_point_cloud_octree_default_dispatch_matches_oracle(monkeypatch):
    """n*m >= 2^30 takes the octree without being asked; checked against the k-d tree oracle."""
    monkeypatch.delenv("AB_NN_ALGO", raising=False)
    import aegolius_b200 as ab
    from aegolius_b200 import workloads
    pts = workloads.c4_cloud(300000, 5)
    spec = ab.GridSpec((2.5, 2.5, 1.5), (32, 32, 32))
    from aegolius_b200 import cabi
    before = cabi.launch_count()
    got32 = ab.point_cloud_sdf(spec, pts, dtype="f32")
    assert cabi.launch_count() - before == 9  # octree build (8 launches) + packet walk
    exp = interp_np.point_cloud_distance_kdtree(spec.materialize(), pts)
    assert np.max(np.abs(got32 - exp)) <= 1e-5 * 2.5
    got64 = ab.point_cloud_sdf(spec, pts, dtype="f64")
    assert np.max(np.abs(got64 - exp)) <= 1e-12 * 2.5
    spec2 = ab.GridSpec((2.5, 2.5), (128, 128))
    exp2 = interp_np.point_cloud_distance_kdtree(spec2.materialize(), pts, dim=2)
    assert np.max(np.abs(ab.point_cloud_sdf(spec2, pts, dim=2, dtype="f64") - exp2)) <= 1e-12 * 2.5
    # slabs of the grid see the same tree
    part = ab.point_cloud_sdf(spec, pts, dtype="f32", slab=(7, 19))
    assert np.array_equal(part, got32.reshape(33, -1)[7:19].ravel())
