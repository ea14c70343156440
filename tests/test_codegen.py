"""Program compiler (aegolius_b200/codegen.py), host side: the signature / hash / argument layout the generated kernels
hard-code must be the library's own, the emitted source must cover every op of every golden program, and a built binary
must register. Nothing here needs a GPU (nvcc cross-compiles)."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from conftest import golden_names, load_case

HAVE_NVCC = shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")


def _some_programs(golden):
    import aegolius_b200 as ab
    progs = [ab.flatten(ab.workloads.build_c1()), ab.flatten(ab.workloads.build_c2()), ab.flatten(ab.workloads.build_c3())]
    for name in ("prim3_segmented_line", "mod_rotational_symmetry", "prim2_polygon_concave", "mod_curve_instancing"):
        if name in golden_names():
            progs.append(load_case(golden, name)["prog"])
    return progs


def test_signature_hash_matches_the_library(golden):
    from aegolius_b200 import cabi, codegen as cg
    lib = cabi.lib()
    for prog in _some_programs(golden):
        sig = np.ascontiguousarray(cg.signature(prog))
        assert len(sig) == prog.n_ops - 1
        for dt, gm in ((0, 0), (1, 1), (0, 2 | (1 << 8))):
            assert cg.signature_hash(sig, dt, gm) == lib.ab_prog_signature_hash(sig.ctypes.data, len(sig), dt, gm)


@pytest.mark.parametrize("name", golden_names())
def test_fixed_argument_offsets_match_the_library_layout(golden, name):
    """The compile-time offsets of the generated kernels == where run_program puts the arguments (tables last)."""
    from aegolius_b200 import cabi, codegen as cg, opcodes as oc
    prog = load_case(golden, name)["prog"]
    if prog.stages:
        pytest.skip("staged program: evaluated through prefixes")
    cp = cabi.CProgram(prog)
    n = prog.n_ops - 1
    offs = (C.c_uint32 * max(1, n))()
    total = C.c_uint32()
    cabi.check(cabi.lib().ab_prog_arg_layout(cp.ref(), offs, C.byref(total)))
    sig = cg.signature(prog)
    mine = cg.fixed_arg_offsets(sig)
    for i, w in enumerate(sig):
        if (int(w) & 0xffff) in cg.TABLE_OPS:
            assert offs[i] >= max(mine.values(), default=0), "tables are packed after the fixed arguments"
        else:
            assert mine[i] == offs[i], (name, i, oc.NAMES[int(w) & 0xffff])


@pytest.mark.parametrize("name", golden_names())
def test_source_is_generated_for_every_golden_program(golden, name):
    from aegolius_b200 import codegen as cg
    c = load_case(golden, name)
    prog = c["prog"]
    sig = cg.signature(prog)
    for dtype, grad in (("f32", "none"), ("f64", "none"), ("f32", "spatial"), ("f64", "param")):
        src = cg.generate(sig, dtype, grad, is2d=not c["res"][2] > 1)
        assert src.count("\n    // ") >= len(sig)  # one commented block per op
        assert "ab_prog_launch" in src and "ab_prog_kernel" in src
    # the structure alone decides the source: other argument values give the same text
    p2 = type(prog)(prog.ops, prog.args * 1.5 + 0.25, prog.blobs, prog.n_pslots, prog.n_vslots, prog.stages)
    assert cg.generate(cg.signature(p2), "f32", "none") == cg.generate(sig, "f32", "none")


@pytest.mark.skipif(not HAVE_NVCC, reason="needs nvcc")
def test_built_binary_registers_and_modes(tmp_path, monkeypatch):
    import aegolius_b200 as ab
    from aegolius_b200 import cabi, codegen as cg
    monkeypatch.setattr(cg, "JIT_DIR", str(tmp_path))
    s = ab.Sphere(0.7)
    s.onion(0.125)
    s.move((0.1, 0.2, 0.3))
    prog = ab.flatten(s)
    monkeypatch.setenv("AB_JIT", "off")
    assert cg.ensure(prog, "f32", None) is False
    monkeypatch.setenv("AB_JIT", "cache")
    assert cg.ensure(prog, "f32", None) is False  # nothing on disk yet, and cache mode never runs nvcc
    assert not os.listdir(tmp_path)
    assert cg.ensure(prog, "f32", None, how="sync") is True
    built = [f for f in os.listdir(tmp_path) if f.endswith(".so")]
    assert len(built) == 1 and not [f for f in os.listdir(tmp_path) if f.endswith((".cu", ".tmp"))]
    lib = C.CDLL(os.path.join(tmp_path, built[0]))
    lib.ab_prog_hash.restype = C.c_uint64
    sig = cg.signature(prog)
    assert lib.ab_prog_kind() == 0 and lib.ab_prog_flavor() == 0
    assert lib.ab_prog_hash() == cg.signature_hash(sig, cabi.AB_F32, cabi.AB_GRAD_NONE)
    # a second process / call finds it on disk
    cg._registered.clear()
    assert cg.ensure(prog, "f32", None) is True
    # a binary built against another KParams layout is refused
    sigc = np.ascontiguousarray(sig)
    rc = cabi.lib().ab_prog_register(sigc.ctypes.data, len(sigc), cabi.AB_F32, cabi.AB_GRAD_NONE, 0,
                                     C.cast(lib.ab_prog_launch, C.c_void_p), 12345)
    assert rc == cabi.AB_EINVAL and b"another library version" in cabi.lib().ab_last_error()
    cg._registered.clear()
    cabi.lib().ab_prog_clear()


def test_long_programs_stay_on_the_interpreter():
    import aegolius_b200 as ab
    from aegolius_b200 import codegen as cg
    objs = [ab.Sphere(0.1 + 0.01 * i) for i in range(140)]
    for i, o in enumerate(objs):
        o.move((0.01 * i, 0, 0))
    u = ab.CombineGeometry("UNION").combine(*objs)
    prog = ab.flatten(u)
    assert len(cg.signature(prog)) > cg.MAX_COMPILED_OPS
    assert cg.ensure(prog, "f32", None, how="sync") is False


from aegolius_b200 import codegen as cg  # noqa: E402


def _body(src):
    return src.split("namespace ab {", 1)[1].split("#define AB_PROG_EXPORT", 1)[0]


def test_gradient_kernels_pull_back_through_the_coordinate_ops_in_reverse_order(golden):
    """_AdjointEmitter on the C3 tree: coordinate ops on plain points with tapes, the deep branch pulled back to the frame
    of the saved point (where the second child lives) right before the combine, the mirror's flip last."""
    import aegolius_b200 as ab
    prog = ab.flatten(ab.workloads.build_c3())
    body = _body(cg.generate(cg.signature(prog), "f32", "spatial"))
    assert "Pt<P> p;" in body and "Pt<S> p;" not in body
    order = [body.index(k) for k in ("fwd_curve_inst(", "fwd_rotsym(", "fwd_bend(", "fwd_twist(", "fwd_elongate(", "grad_torus(",
                                     "grad_sphere(", "pb_elongate(V0", "pb_affine(V0", "pb_twist(V0", "pb_bend(V0", "pb_rot(V0",
                                     "pb_curve_inst(V0", "smin_poly3(V0, acc", "pb_flip<0>(acc")]
    assert order == sorted(order)
    # forward tangents on request, and for programs with an op that has no pull-back
    assert "pb_" not in _body(cg.generate(cg.signature(prog), "f32", "spatial", adjoint=False))
    staged = load_case(golden, "stencil_signed_union3d")["prog"]  # P_FIELD (a grid stencil's output) has no derivative
    fallback = _body(cg.generate(cg.signature(staged), "f32", "spatial"))
    assert "Pt<S> p;" in fallback and "pb_" not in fallback
    polygon = load_case(golden, "shape_closed_segmented_line_polygon")["prog"]  # the interior sign multiplies the value
    assert "op_poly_sign(acc, S(P" in _body(cg.generate(cg.signature(polygon), "f32", "spatial"))
    # a point cloud is a leaf like any other: evaluated on the identity-seeded copy of the point
    cloud = ab.PointCloud3D(np.random.default_rng(0).uniform(-1, 1, size=(3, 50)))
    cloud.twist(0.4)
    body = _body(cg.generate(cg.signature(ab.flatten(cloud)), "f32", "spatial"))
    assert "prim_point_cloud<S, T>(q, " in body and body.index("fwd_twist(") < body.index("pb_twist(acc")
    # value kernels and parameter-tangent kernels are untouched
    assert "pb_" not in _body(cg.generate(cg.signature(prog), "f32", "none"))
    assert "pb_" not in _body(cg.generate(cg.signature(prog), "f64", "param"))


def test_extrusion_and_translations_in_the_pull_back_emitter():
    """EXTRUDE_BEGIN stores |z| - h in the frame before z is zeroed; EXTRUDE_END pulls both operands to their common frame.
    Translations and repetitions have an identity Jacobian: no pull-back statement, no frame."""
    import aegolius_b200 as ab
    c = ab.Circle(0.5)
    c.extrusion(0.4)
    c.twist(0.7)
    body = _body(cg.generate(cg.signature(ab.flatten(c)), "f64", "spatial"))
    assert body.index("fwd_twist(") < body.index("seed_local(q, p); V0 = abs_(q.z)") < body.index("pb_zero_z(acc)") \
        < body.index("op_extrude_end<S, T>(acc, V0)") < body.index("pb_twist(acc")
    s = ab.Sphere(0.3)
    s.infinite_repetition((1.0, 1.0, 1.0))
    s.move((0.2, 0.1, 0.0))
    body = _body(cg.generate(cg.signature(ab.flatten(s)), "f32", "spatial"))
    assert "op_rep_inf(p" in body and "pb_" not in body.split("// gradient back to grid coordinates")[1].split("emit")[0]
