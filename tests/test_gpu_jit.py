"""Program-compiled kernels (aegolius_b200/codegen.py) — the default evaluation path — against the interpreter and the
reference's own outputs, on every golden scenario, through the C ABI.

What is asserted:
  * the compiled kernel really ran (ab_prog_hits moves) and meets the same parity bounds against the reference's outputs
    as the interpreter does (tests/test_gpu_parity.py: fp32 1e-5 * extent + sign mask, fp64 1e-12 * extent);
  * fp64: bit-identical to the interpreter (scalar mul.rn / add.rn are never contracted);
  * fp32: identical except where ptxas contracts the packed multiply that ends one op with the packed add that starts the
    next (CUDA 12.9 ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false; the contracted form has one
    rounding LESS; in a tree that scales every child, like C2, about a fifth of the points differ in the last bit).
    Bound: 2e-6 * extent away from the branch boundaries the oracle reports.
"""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from conftest import GRADIENT_CASES, golden_names, load_case
from oracle import interp_np
from test_gpu_parity import _compare, _spec, F32_TOL, F64_TOL

pytestmark = pytest.mark.gpu
HAVE_NVCC = shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")


def _both(prog, spec, dtype, **kw):
    """(compiled result, interpreter result, compiled launches)"""
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    lib = cabi.lib()
    h0 = lib.ab_prog_hits()
    a = ab.create(prog, spec, dtype=dtype, **kw)
    hits = lib.ab_prog_hits() - h0
    old = lib.ab_prog_enable(0)
    try:
        b = ab.create(prog, spec, dtype=dtype, **kw)
    finally:
        lib.ab_prog_enable(old)
    return a, b, hits


@pytest.mark.parametrize("name", golden_names())
def test_compiled_kernels_match_reference_and_interpreter(golden, name):
    c = load_case(golden, name)
    spec = _spec(c)
    _, margin = interp_np.run_grid(c["prog"], c["size"], c["res"], return_margin=True)
    ext = c["extent"]
    # fp64
    a, b, hits = _both(c["prog"], spec, "f64")
    assert hits >= 1, f"{name}: no compiled fp64 kernel was used (is aegolius_b200/jit/ prebuilt? python tools/prebuild_jit.py)"
    _compare(a, c["expected"], margin, ext, F64_TOL, 1e-9, name)
    assert np.array_equal(a, b, equal_nan=True), f"{name}: fp64 compiled != interpreter, max |d| = {np.nanmax(np.abs(a - b)):.3e}"
    # fp32
    a, b, hits = _both(c["prog"], spec, "f32")
    assert hits >= 1, f"{name}: no compiled fp32 kernel was used"
    _compare(a, c["expected"], margin, ext, F32_TOL, 2e-6, name)
    keep = margin > 2e-6 * ext
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))[keep & ~np.isnan(a)]
    assert d.size == 0 or d.max() <= 2e-6 * ext, f"{name}: fp32 compiled vs interpreter max |d| = {d.max():.3e}"


@pytest.mark.parametrize("name", GRADIENT_CASES)
def test_compiled_gradient_kernels_on_point_lists(golden, name):
    """Field + gradient kernels, points mode, ragged size: compiled vs interpreter. The compiled kernels pull the gradient
    back through the coordinate ops (csrc/ab_adjoint.cuh) where the interpreter pushes three tangents forward: the FIELD is
    the same arithmetic (fp64 to the bit), the GRADIENT the same derivative in another association order (fp64 within
    1e-11, fp32 within 2e-5, both relative to the largest gradient component of the scenario, away from the kinks the
    oracle reports). Measured over all 134 single-stage golden scenarios (tools/check_adjoint.py): 2.8e-14 / 3.8e-6."""
    c = load_case(golden, name)
    rng = np.random.default_rng(5)
    size3 = np.ones(3)
    size3[:len(c["size"])] = c["size"]
    co = rng.uniform(-0.45, 0.45, size=(3, 10007)) * size3.reshape(3, 1)
    _, margin = interp_np.run(c["prog"], co, return_margin=True)
    keep = margin > 1e-4 * c["extent"]
    assert keep.mean() > 0.5
    (fa, ga), (fb, gb), hits = _both(c["prog"], co, "f64", grad="spatial")
    assert hits >= 1, f"{name}: no compiled fp64 gradient kernel (python tools/prebuild_jit.py)"
    assert np.array_equal(fa, fb, equal_nan=True)
    scale = max(1.0, float(np.max(np.abs(gb[:, keep]))))
    assert np.max(np.abs(ga - gb)[:, keep]) <= 1e-11 * scale
    (fa, ga), (fb, gb), hits = _both(c["prog"], co, "f32", grad="spatial")
    assert hits >= 1
    assert np.max(np.abs(fa - fb)[keep]) <= 2e-6 * c["extent"]
    assert np.max(np.abs(ga - gb)[:, keep]) <= 2e-5 * scale


@pytest.mark.parametrize("name", ["c3_deep_tree", "c1_sphere_box_smooth_union"])
def test_forward_tangent_build_matches_the_interpreter_gradient_to_the_bit(golden, name):
    """AB_JIT_ADJOINT=0 / generate(adjoint=False): the straight-line kernel that carries three tangents forward like the
    interpreter. fp64: field AND gradient bit-identical to the interpreter (the pull-back build is only rounding-close)."""
    from aegolius_b200 import codegen as cg
    c = load_case(golden, name)
    sig = cg.signature(c["prog"])
    forward = cg.binary_path(cg.generate(sig, "f64", "spatial", adjoint=False))
    default = cg.binary_path(cg.generate(sig, "f64", "spatial"))
    if not (os.path.exists(forward) and os.path.exists(default)):
        pytest.skip("kernels not prebuilt (python tools/prebuild_jit.py)")
    rng = np.random.default_rng(7)
    co = rng.uniform(-0.45, 0.45, size=(3, 10007)) * np.asarray(c["size"]).reshape(3, 1)
    cg.ensure(c["prog"], "f64", "spatial")  # the default build first, so that create() does not register it over ours
    cg._register(forward, sig, "f64", "spatial", 0)  # re-registration replaces the launcher of this structure
    try:
        (fa, ga), (fb, gb), hits = _both(c["prog"], co, "f64", grad="spatial")
    finally:
        cg._register(default, sig, "f64", "spatial", 0)
    assert hits >= 1
    assert np.array_equal(fa, fb, equal_nan=True) and np.array_equal(ga, gb, equal_nan=True)


@pytest.mark.parametrize("name", ["struct_2d_mirror_rotsym", "c2"])
def test_compiled_gradient_kernels_on_2d_grids(golden, name):
    """Field + gradient on a 2D GRID (flavour bit 0: in-kernel coordinates of a (nx, ny) grid, z = 0): the pull-back build
    against the interpreter, fp64 and fp32, odd row length."""
    import aegolius_b200 as ab
    if name == "c2":
        prog, size, ext = ab.flatten(ab.workloads.build_c2()), (8.0, 8.0), 8.0
    else:
        c = load_case(golden, name)
        prog, size, ext = c["prog"], tuple(float(v) for v in c["size"][:2]), c["extent"]
    spec = ab.GridSpec(size, (300, 212))  # 301 x 213 samples
    _, margin = interp_np.run_grid(prog, spec.size, spec.res, return_margin=True)
    keep = margin > 1e-4 * ext
    for dt, tol_f, tol_g in (("f64", 0.0, 1e-11), ("f32", 2e-6 * ext, 2e-5)):
        (fa, ga), (fb, gb), hits = _both(prog, spec, dt, grad="spatial")
        assert hits >= 1, f"{name}: no compiled {dt} gradient kernel for 2D grids (python tools/prebuild_jit.py)"
        scale = max(1.0, float(np.max(np.abs(gb[:, keep]))))
        assert np.max(np.abs(fa - fb)[keep]) <= tol_f
        assert np.max(np.abs(ga - gb)[:, keep]) <= tol_g * scale


def test_compiled_slabs_concatenate_bit_identically(golden):
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    c = load_case(golden, "c3_deep_tree")
    spec = _spec(c)
    h0 = cabi.lib().ab_prog_hits()
    whole = ab.create(c["prog"], spec, dtype="f32")
    for parts in (2, 3, 8):
        pieces = [ab.create(c["prog"], spec, dtype="f32", slab=s) for s in ab.engine.slab_ranges(spec.res[0], parts)]
        assert np.array_equal(np.concatenate(pieces), whole)
    assert cabi.lib().ab_prog_hits() - h0 == 1 + 2 + 3 + 8


def test_one_binary_serves_every_argument_value():
    """The structure is the key: moving / resizing the shapes reuses the registered kernel (no rebuild, same hits path)."""
    import aegolius_b200 as ab
    from aegolius_b200 import cabi, codegen as cg
    spec = ab.GridSpec((4, 4, 4), (40, 36, 44))

    def tree(r, t):
        s = ab.Sphere(r)
        s.move((t, 0.1, -0.2))
        b = ab.Box(1.5 * r, 1.0, 0.8)
        b.rotate(0.3 + t, (0, 0, 1))
        b.move((-0.4, 0.2, 0.1))
        return ab.CombineGeometry("SMOOTH_UNION2").combine_parametric(s, b, parameters=0.3)

    p0 = ab.flatten(tree(1.0, 0.5))
    assert np.array_equal(cg.signature(p0), cg.signature(ab.flatten(ab.workloads.build_c1())))
    n_reg = len(cg._registered)
    for r, t in ((1.0, 0.5), (0.8, 0.1), (1.3, -0.4)):
        a, b, hits = _both(tree(r, t), spec, "f32")
        assert hits == 1
        assert np.max(np.abs(a - b)) <= 2e-6 * 4
        exp = interp_np.run_grid(ab.flatten(tree(r, t)), spec.size, spec.res)
        assert np.max(np.abs(a - exp)) <= F32_TOL * 4
    assert len(cg._registered) <= n_reg + 1


@pytest.mark.skipif(not HAVE_NVCC, reason="needs nvcc at run time")
def test_background_build_switches_over_without_changing_results(tmp_path, monkeypatch):
    """AB_JIT=on (the shipped default): the first evaluations of a new structure run on the interpreter while nvcc works in
    the background; once the binary is registered the same call runs on it. Same numbers before and after."""
    import aegolius_b200 as ab
    from aegolius_b200 import cabi, codegen as cg
    monkeypatch.setattr(cg, "JIT_DIR", str(tmp_path))
    monkeypatch.setenv("AB_JIT", "on")
    t = ab.Torus(0.9, 0.2)
    t.twist(0.7)
    t.onion(0.03)
    t.rotate(0.4, (1, 0, 0))
    s = ab.Cylinder(0.3, 1.1)
    u = ab.CombineGeometry("SMOOTH_SUBTRACT2").combine_parametric(t, s, parameters=0.15)
    prog = ab.flatten(u)
    spec = ab.GridSpec((3, 3, 3), (48, 40, 56))
    lib = cabi.lib()
    h0 = lib.ab_prog_hits()
    first = ab.create(prog, spec, dtype="f64")
    assert lib.ab_prog_hits() == h0, "the build cannot be ready yet: the interpreter serves the first call"
    ab.wait_for_compilations(120)
    assert not cg._failed, cg._failed
    second = ab.create(prog, spec, dtype="f64")
    assert lib.ab_prog_hits() == h0 + 1
    assert np.array_equal(first, second)
    exp = interp_np.run_grid(prog, spec.size, spec.res)
    assert np.max(np.abs(second - exp)) <= F64_TOL * 3


def test_compiled_parameter_tangent_and_fused_loss(golden):
    """AB_GRAD_PARAM kernels (jacfwd / value_and_grad): compiled vs interpreter, stored maps and the fused loss reduction."""
    import torch
    import aegolius_b200 as ab
    from aegolius_b200 import cabi, codegen as cg

    def geometry(r, w):
        s = ab.Sphere(r)
        s.move((0.3, 0.0, 0.1))
        b = ab.Box(1.2, 0.9, 0.7)
        return ab.CombineGeometry("SMOOTH_UNION2").combine_parametric(s, b, parameters=w)

    spec = ab.GridSpec((4, 4, 4), (40, 40, 40))
    prog = ab.engine.program_tangent(geometry, (0.9, 0.3), 0)
    assert cg.ensure(prog, "f64", "param", how="sync" if HAVE_NVCC else None) or not HAVE_NVCC
    (fa, da), (fb, db), hits = _both(prog, spec, "f64", grad="param")
    if HAVE_NVCC:
        assert hits == 1
    assert np.array_equal(fa, fb) and np.array_equal(da, db)
    target = torch.as_tensor(fa + 0.01, device="cuda")
    f = ab.value_and_grad(geometry, spec, target, dtype="f64")
    la, ga = f((0.9, 0.3))
    old = cabi.lib().ab_prog_enable(0)
    try:
        lb, gb = f((0.9, 0.3))
    finally:
        cabi.lib().ab_prog_enable(old)
    assert abs(la - lb) <= 1e-12 * abs(lb) and np.allclose(ga, gb, rtol=1e-12, atol=1e-14)
