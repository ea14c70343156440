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
