"""Host logic without a GPU: the SPOMSO-mirroring front end, the flattener and grid helpers."""
import os
import sys

import numpy as np
import pytest

import aegolius_b200 as ab
from aegolius_b200 import opcodes as oc
from oracle import interp_np
from scenarios import SCENARIOS, make_namespace

REF = os.path.isdir("/root/reference/Code/spomso")
if REF and "/root/reference/Code/spomso" not in sys.path:
    sys.path.insert(0, "/root/reference/Code/spomso")


def test_rodrigues_matches_scipy_rotation():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = rng.normal(size=3) * rng.uniform(0, 3)
        assert np.allclose(ab.frontend.rotation_matrix_from_rotvec(v), Rotation.from_rotvec(v).as_matrix(), atol=1e-15)
    assert np.array_equal(ab.frontend.rotation_matrix_from_rotvec(np.zeros(3)), np.eye(3))


def test_transform_state_follows_the_reference_rules():
    s = ab.Sphere(1.0)
    s.move((1, 2, 3))
    s.move((1, 0, 0))
    assert np.array_equal(s.center, (2, 2, 3))
    s.move(0.5)  # scalar broadcasts, as `self._center += vector` does in transformations.py:104
    assert np.array_equal(s.center, (2.5, 2.5, 3.5))
    with pytest.raises(ValueError):
        s.move((1, 2))  # a 2-vector cannot be added to the 3-vector (same failure as the reference)
    s.set_location((7, 8))
    assert np.array_equal(s.center, (7, 8, 3.5))
    s.rescale(2)
    s.rescale(1.5)
    assert s.scale == 3.0
    with pytest.raises(TypeError):
        s.rescale("2")
    with pytest.raises(SyntaxError):
        s.move((1, 2, 3, 4))
    s.rotate(np.pi / 2, (0, 0, 2))   # rotate() normalises the axis ...
    r1 = s.rotation_matrix.copy()
    s.set_rotation(np.pi / 2, (0, 0, 2))  # ... set_rotation() does not (transformations.py:163-173): angle * |axis|
    assert np.allclose(r1 @ np.array([1, 0, 0]), (0, 1, 0))
    assert np.allclose(s.rotation_matrix @ np.array([1, 0, 0]), (-1, 0, 0))
    with pytest.raises(ValueError):
        s.rotate(0.3, (0, 0, 0))
    a = ab.Box(1, 1, 1)
    a.rotate(0.3, (1, 0, 0))
    a.rotate(0.4, (0, 1, 0))  # accumulates by LEFT multiplication
    ra = ab.frontend.rotation_matrix_from_rotvec
    assert np.allclose(a.rotation_matrix, ra((0, 0.4, 0)) @ ra((0.3, 0, 0)))


def test_combine_api_errors_match_the_reference():
    with pytest.raises(SyntaxError):
        ab.CombineGeometry("NOPE").combine(ab.Sphere(1), ab.Sphere(2))
    with pytest.raises(SyntaxError):
        ab.CombineGeometry("UNION2").combine_parametric(ab.Sphere(1), ab.Sphere(2), parameters=0.1)
    with pytest.raises(TypeError):  # the reference's 2-argument lambda raises TypeError at create()
        ab.flatten(ab.CombineGeometry("UNION2").combine(ab.Sphere(1), ab.Sphere(2), ab.Sphere(3)))
    u = ab.CombineGeometry("UNION").combine(ab.Sphere(1), ab.Sphere(2), ab.Sphere(3))
    assert [int(o["opcode"]) for o in ab.flatten(u).ops].count(oc.C_UNION) == 2


def test_unflattenable_trees_raise_not_implemented_naming_the_culprit():
    s = ab.Sphere(1.0)
    for name in ("custom_modification", "displacement", "define_volume", "custom_post_process"):
        with pytest.raises(NotImplementedError, match=name):
            getattr(s, name)(lambda *a: 0, ())
    with pytest.raises(NotImplementedError):
        ab.GenericGeometry(lambda co: co[0])


def test_signed_becomes_a_stage_and_needs_a_3d_resolution():
    """signed (modifications.py:220-275): a stage of kind 2 fed by the ops before it; a 2-entry co_resolution fails like the
    reference's three-index slicing of the reshaped field does."""
    s = ab.Sphere(1.0)
    s.boundary()
    s.signed((16, 12, 20))
    s.rounding(0.1)
    prog = ab.flatten(s)
    assert [oc.NAMES[int(o["opcode"])] for o in prog.ops] == ["P_SPHERE", "ABS", "P_FIELD", "ROUND", "END"]
    assert [st["kind"] for st in prog.stages] == [2] and prog.stages[0]["res"] == (16, 12, 20)
    c = ab.Circle(1.0)
    c.signed((16, 12))
    with pytest.raises(IndexError):
        ab.flatten(c)


def test_grid_stencil_modifications_become_stages():
    """conv_averaging / conv_edge_detection as modifications: a P_FIELD op in the list + a stage record; the prefix
    program of a stage ends right before its P_FIELD; programs with stages survive serialisation."""
    s = ab.Sphere(1.0)
    s.conv_averaging((3, 3, 3), 2, (16, 12, 20))
    s.rounding(0.1)
    s.conv_edge_detection((16, 12, 20))
    prog = ab.flatten(s)
    names = [oc.NAMES[int(o["opcode"])] for o in prog.ops]
    assert names == ["P_SPHERE", "P_FIELD", "ROUND", "P_FIELD", "END"]
    assert [st["kind"] for st in prog.stages] == [0, 1] and prog.stages[0]["ksize"] == (3, 3, 3)
    assert prog.stages[0]["iterations"] == 2 and prog.stages[1]["res"] == (16, 12, 20)
    pre = prog.prefix(prog.stage_op_index(prog.stages[1]))
    assert [oc.NAMES[int(o["opcode"])] for o in pre.ops] == ["P_SPHERE", "P_FIELD", "ROUND", "END"]
    assert len(pre.stages) == 1
    back = ab.Program.from_arrays(prog.to_arrays())
    assert back.stages == prog.stages and np.array_equal(back.ops, prog.ops)
    # zero iterations leave the field untouched (post_processing.py:573-574): no stage at all
    z = ab.Sphere(1.0)
    z.conv_averaging(3, 0, (8, 8, 8))
    assert not ab.flatten(z).stages
    with pytest.raises(ValueError):
        bad = ab.Sphere(1.0)
        bad.conv_averaging((3, 3), 1, (8, 8, 8))
        ab.flatten(bad)


def test_late_binding_of_children_like_the_reference_closures():
    a, b = ab.Sphere(0.5), ab.Box(1, 1, 1)
    u = ab.CombineGeometry("UNION2").combine(a, b)
    p0 = ab.flatten(u)
    a.move((1, 0, 0))  # mutated AFTER combine(): seen at create() time (combine.py:129-135 is late-bound)
    p1 = ab.flatten(u)
    assert not np.array_equal(p0.args, p1.args)


def test_program_evaluation_order_and_slot_sharing():
    u = ab.workloads.build_c1()
    names = [oc.NAMES[int(o["opcode"])] for o in ab.flatten(u).ops]
    assert names == ["SAVE_P", "TRANSLATE", "P_SPHERE", "NEXT_AFFINE", "P_BOX", "C_SMIN3", "END"]  # fused child entry
    raw = [oc.NAMES[int(o["opcode"])] for o in ab.flatten(u, optimize=False).ops]
    assert raw == ["SAVE_P", "AFFINE", "P_SPHERE", "PUSH_V", "LOAD_P", "AFFINE", "P_BOX", "C_SMIN3", "END"]
    # last-called modification is the outermost wrapper: coordinate parts run in reverse call order
    b = ab.Box(1, 1, 1)
    b.rounding(0.1)
    b.elongation((0.2, 0, 0))
    b.twist(1.0)
    names = [oc.NAMES[int(o["opcode"])] for o in ab.flatten(b).ops]
    assert names == ["TWIST", "ELONGATE", "P_BOX", "ROUND", "END"]
    # a left-deep chain of identity-transform combines shares ONE saved-coordinate slot
    c2 = ab.flatten(ab.workloads.build_c2())
    assert c2.n_pslots == 1 and c2.n_ops > 80


def test_peephole_composes_affine_runs():
    s = ab.Sphere(1.0)
    s.move_sdf((1, 0, 0))
    s.rotate_sdf(ab.frontend.rotation_matrix_from_rotvec((0, 0, 0.3)))
    s.shear_xz(0.2)
    s.rotate(0.5, (0, 1, 0))
    s.rescale(1.5)
    s.move((0.1, 0.2, 0.3))
    raw = ab.flatten(s, optimize=False)
    opt = ab.flatten(s, optimize=True)
    assert raw.n_ops > opt.n_ops and [oc.NAMES[int(o["opcode"])] for o in opt.ops] == ["AFFINE", "P_SPHERE", "SCALE_V", "END"]
    co = np.random.default_rng(3).uniform(-2, 2, size=(3, 500))
    assert np.max(np.abs(interp_np.run(raw, co) - interp_np.run(opt, co))) < 1e-13


def test_generate_grid_matches_reference_layout_and_detection():
    co, res = ab.generate_grid((4, 2, 6), (8, 5, 6))
    assert res == (9, 5, 7) and co.shape == (3, 9 * 5 * 7) and co.dtype == np.float64
    assert np.array_equal(co[2, :7], np.linspace(-3, 3, 7))           # z is the fastest axis
    assert np.array_equal(co[0, ::5 * 7], np.linspace(-2, 2, 9))      # x the slowest
    spec = ab.detect_grid(np.array(co))
    assert spec is not None and spec.res == (9, 5, 7) and spec.size == (4.0, 2.0, 6.0)
    co2, res2 = ab.generate_grid((8, 8), (16, 12))
    assert res2 == (17, 13, 17) and co2.shape == (3, 17 * 13) and not co2[2].any()
    spec2 = ab.detect_grid(np.array(co2))
    assert spec2 is not None and spec2.res == (17, 13, 1) and spec2.dims == 2
    assert ab.detect_grid(np.random.default_rng(0).normal(size=(3, 100))) is None
    assert ab.detect_grid(np.array(co)[:, ::-1].copy()) is None
    assert ab.smarter_reshape(np.zeros(9 * 5 * 7), (8, 5, 6)).shape == (9, 5, 7)
    with pytest.raises(IndexError):
        ab.generate_grid((4,), (8,))


@pytest.mark.reference
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_introspected_spomso_objects_flatten_like_the_mirror_front_end(name):
    """The same construction code run against real SPOMSO classes (flattened by closure introspection) and against the
    mirror front end must give the identical program."""
    sc = SCENARIOS[name]
    p_ref = ab.flatten(sc["build"](make_namespace("reference")))
    p_own = ab.flatten(sc["build"](make_namespace("frontend")))
    assert np.array_equal(p_ref.ops, p_own.ops)
    assert np.array_equal(p_ref.args, p_own.args)
    assert (p_ref.n_pslots, p_ref.n_vslots) == (p_own.n_pslots, p_own.n_vslots)


@pytest.mark.reference
def test_generate_grid_is_bit_identical_to_spomso():
    from spomso.cores.helper_functions import generate_grid
    for size, res in (((4, 4, 4), (16, 12, 20)), ((8, 8), (64, 48)), ((2.5, 2.5, 1.5), (7, 9, 5))):
        a, ra = ab.generate_grid(size, res)
        b, rb = generate_grid(size, res)
        assert ra == rb and np.array_equal(np.asarray(a), b)


@pytest.mark.reference
def test_introspection_rejects_custom_closures_by_name():
    ns = make_namespace("reference")
    s = ns.Sphere(1.0)
    s.custom_modification(lambda f, co, p, mp: f(co, *p), (), "wobble")
    with pytest.raises(NotImplementedError, match="custom_modification"):
        ab.flatten(s)


@pytest.mark.reference
def test_patch_rebinds_spomso_create_and_has_no_silent_fallback():
    """aegolius_b200.patch() routes an unmodified SPOMSO object's create() to the GPU library; in this GPU-less container
    that must raise NoDeviceError (there is no CPU fallback), and unpatch() restores SPOMSO's own NumPy path."""
    from spomso.cores.geom_3d import Sphere
    from spomso.cores.helper_functions import generate_grid
    from aegolius_b200 import cabi, engine
    co, _ = generate_grid((2, 2, 2), (4, 4, 4))
    s = Sphere(0.5)
    ref = s.create(co)
    engine.patch(dtype="f64")
    try:
        if cabi.device_count() == 0:
            with pytest.raises(cabi.NoDeviceError):
                s.create(co)
        else:
            assert np.max(np.abs(s.create(co) - ref)) < 1e-12
    finally:
        engine.unpatch()
    assert np.array_equal(s.create(co), ref)


def test_affine_after_aligned_instancing_is_folded_into_the_frames():
    """program._peephole(fold_frames): p' = M R_j (p - o_j) + b keeps only the translation as an op. Same field as the
    unfolded program (oracle, fp64 rounding), one op less in C3; program_tangent keeps the affine (tables carry no
    parameter tangents)."""
    import numpy as np
    import aegolius_b200 as ab
    from aegolius_b200 import opcodes as oc, engine
    from aegolius_b200.program import flatten
    from oracle import interp_np
    obj = ab.workloads.build_c3()
    folded, plain = flatten(obj), flatten(obj, fold_frames=False)
    codes = lambda p: [int(c) for c in p.ops["opcode"]]
    assert len(codes(folded)) == len(codes(plain)) - 1
    i = codes(plain).index(oc.CURVE_INST)
    assert codes(plain)[i + 1] == oc.AFFINE and codes(folded)[i + 1] != oc.AFFINE
    co = np.random.default_rng(3).uniform(-3, 3, size=(3, 5000))
    assert np.max(np.abs(interp_np.run(folded, co) - interp_np.run(plain, co))) <= 1e-13

    def build(angle):
        t = ab.Torus(0.4, 0.1)
        t.shear_xz(angle)  # modifications run outermost first: the instancing, then the shear (an AFFINE op)
        t.aligned_curve_instancing(lambda s: np.stack([np.cos(s), np.sin(s), 0 * s]), (), (0.0, 2 * np.pi, 9))
        return t

    assert oc.AFFINE not in codes(flatten(build(0.3)))[1:]  # folded on the default path

    prog = engine.program_tangent(build, [0.3], 0)
    k = codes(prog).index(oc.CURVE_INST)
    assert codes(prog)[k + 1] == oc.AFFINE and np.any(prog.dargs != 0)


def test_pruned_drops_exactly_the_dead_ops(golden):
    """Program.pruned(): backward liveness over (p, acc, P slots, V slots). No golden program without a stencil stage
    holds dead code; a staged one loses the subtree whose value its P_FIELD op overwrites; and a program with a dead
    branch spliced in evaluates (oracle) to the same field after pruning."""
    import numpy as np
    import aegolius_b200 as ab
    from aegolius_b200 import opcodes as oc
    from aegolius_b200.program import Program, OP_DTYPE
    from conftest import golden_names, load_case
    from oracle import interp_np
    names = lambda p: [oc.NAMES[int(c)] for c in p.ops["opcode"]]
    for name in golden_names():
        prog = load_case(golden, name)["prog"]
        if not prog.stages:
            assert prog.pruned() is prog, name
    tree = load_case(golden, "stencil_conv_averaging_tree3d")["prog"]
    assert names(tree.pruned()) == ["P_FIELD", "END"]
    second = tree.prefix(tree.stage_op_index(sorted(tree.stages, key=tree.stage_op_index)[1]))
    assert names(second.pruned()) == ["SAVE_P", "P_FIELD", "ROUND", "NEXT_AFFINE", "P_BOX", "C_SMIN3", "END"]
    # C1 (a smooth union that uses P and V slots), then a leaf that overwrites its value: only the leaf's inputs survive
    c1 = ab.flatten(ab.workloads.build_c1())
    sph = ab.Sphere(0.7)
    sph.move((0.2, -0.1, 0.3))
    tail = ab.flatten(sph)  # TRANSLATE, P_SPHERE, END with its own argument pool
    shift = len(c1.args)
    tail_ops = tail.ops[:-1].copy()
    tail_ops["arg"] += shift
    load = np.array([(oc.LOAD_P, 0, 0, 0)], dtype=OP_DTYPE)
    ops = np.concatenate([c1.ops[:-1], load, tail_ops, c1.ops[-1:]])
    spliced = Program(ops, np.concatenate([c1.args, tail.args]), [], c1.n_pslots, c1.n_vslots, [])
    assert int(c1.ops["opcode"][0]) == oc.SAVE_P and int(c1.ops["a"][0]) == 0  # slot 0 holds the grid point
    pruned = spliced.pruned()
    assert names(pruned) == ["SAVE_P", "LOAD_P", "TRANSLATE", "P_SPHERE", "END"]
    co = np.random.default_rng(1).uniform(-2, 2, size=(3, 4000))
    assert np.array_equal(interp_np.run(pruned, co), interp_np.run(spliced, co))
    assert np.array_equal(interp_np.run(pruned, co), interp_np.run(tail, co))
