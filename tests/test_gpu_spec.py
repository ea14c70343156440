"""Program-specialised kernels (engine.specialize): the interpreter compiled with only the ops of one program must return
the same bits as the general tiers, be picked only for programs it covers, and be refused when built against another
library layout."""
import ctypes as C
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _interpreter_only():
    """These tests are about the interpreter builds: keep the program-compiled kernels (the default path) out of the way."""
    from aegolius_b200 import cabi
    old = cabi.lib().ab_prog_enable(0)
    yield
    cabi.lib().ab_prog_enable(old)


@pytest.mark.skipif(shutil.which("nvcc") is None and not __import__("os").path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs nvcc at run time")
def test_specialised_kernel_is_bit_identical_and_selected_by_coverage():
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    lib = cabi.lib()
    lib.ab_spec_clear()
    tree = ab.workloads.build_c3()
    spec = ab.GridSpec((6, 6, 6), (96, 80, 72))
    f0, g0 = ab.create(tree, spec, dtype="f32", grad="spatial")
    v0 = ab.create(tree, spec, dtype="f32")
    hits = lib.ab_spec_hits()
    ab.specialize(tree, dtype="f32", grad="spatial")
    f1, g1 = ab.create(tree, spec, dtype="f32", grad="spatial")
    assert lib.ab_spec_hits() > hits, "the specialised kernel was not used"
    assert np.array_equal(f0, f1) and np.array_equal(g0, g1)
    # other modes / programs with other ops keep the general kernels
    hits = lib.ab_spec_hits()
    assert np.array_equal(ab.create(tree, spec, dtype="f32"), v0)
    c1 = ab.workloads.build_c1()  # contains P_BOX: not covered
    ab.create(c1, spec, dtype="f32", grad="spatial")
    assert lib.ab_spec_hits() == hits
    # a sub-program of the covered op set is covered too
    s = ab.Sphere(0.7)
    s.onion(0.1)
    s.twist(0.5)
    a = ab.create(s, spec, dtype="f32", grad="spatial")
    assert lib.ab_spec_hits() > hits
    lib.ab_spec_clear()
    b = ab.create(s, spec, dtype="f32", grad="spatial")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # values-only and fp64 specialisations
    ab.specialize(tree, dtype="f32")
    ab.specialize(tree, dtype="f64", grad="spatial")
    assert np.array_equal(ab.create(tree, spec, dtype="f32"), v0)
    f64 = ab.create(tree, spec, dtype="f64", grad="spatial")
    lib.ab_spec_clear()
    g64 = ab.create(tree, spec, dtype="f64", grad="spatial")
    assert np.array_equal(f64[0], g64[0]) and np.array_equal(f64[1], g64[1])


def test_specialised_tangent_kernel_in_the_optimisation_loop():
    """value_and_grad(..., specialized=True): the fused loss reduction on a specialised AB_GRAD_PARAM kernel gives the same
    loss and gradient as the general kernel."""
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    lib = cabi.lib()
    lib.ab_spec_clear()
    spec = ab.GridSpec((4, 4, 4), (40, 36, 32))

    def geometry(r, w):
        s = ab.Sphere(r)
        s.twist(0.4)
        b = ab.Box(1.5, 1.0, 0.8)
        return ab.CombineGeometry("SMOOTH_UNION2").combine_parametric(s, b, parameters=w)

    target = ab.create(geometry(1.0, 0.25), spec, dtype="f32")
    l0, g0 = ab.value_and_grad(geometry, spec, target)([0.8, 0.3])
    hits = lib.ab_spec_hits()
    l1, g1 = ab.value_and_grad(geometry, spec, target, specialized=True)([0.8, 0.3])
    assert lib.ab_spec_hits() >= hits + 2
    assert abs(l0 - l1) <= 1e-9 * abs(l0) and np.allclose(g0, g1, rtol=1e-9, atol=0)
    lib.ab_spec_clear()


def test_auto_specialisation_switches_over_in_the_background():
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    lib = cabi.lib()
    lib.ab_spec_clear()
    tree = ab.workloads.build_c2()
    spec = ab.GridSpec((8, 8), (300, 260))
    ab.set_auto_specialize(True)
    try:
        a = ab.create(tree, spec, dtype="f32")          # general kernel; the build starts in the background
        ab.wait_for_specializations(timeout=300)
        hits = lib.ab_spec_hits()
        b = ab.create(tree, spec, dtype="f32")          # specialised kernel from now on
        assert lib.ab_spec_hits() > hits
        assert np.array_equal(a, b)
    finally:
        ab.set_auto_specialize(False)
        lib.ab_spec_clear()


def test_specialisation_registry_rejects_foreign_layouts():
    from aegolius_b200 import cabi, opcodes as oc
    lib = cabi.lib()
    mask = (C.c_uint8 * oc.OP_COUNT)()
    fn = C.cast(lib.ab_version, C.c_void_p)
    assert lib.ab_spec_register(cabi.AB_F32, cabi.AB_GRAD_NONE, mask, oc.OP_COUNT, fn, 12345) == -1
    assert b"KParams" in lib.ab_last_error()
    assert lib.ab_spec_register(cabi.AB_F32, 7, mask, oc.OP_COUNT, fn, 12345) == -1
    assert lib.ab_spec_register(cabi.AB_F32, cabi.AB_GRAD_NONE, mask, 7, fn, 12345) == -1
