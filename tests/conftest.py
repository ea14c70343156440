import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


# Kernel builds in tests: by default only binaries already in aegolius_b200/jit/ are used (python tools/prebuild_jit.py,
# run by __graft_entry__.build()), nvcc is never started behind a test's back. tests/test_gpu_jit.py switches modes itself.
os.environ.setdefault("AB_JIT", "cache")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the SPOMSO sources under /root/reference (build container only)")


def pytest_sessionstart(session):
    """The library is a build artefact (git-ignored): build it once if a fresh checkout has none (nvcc cross-compiles
    without a GPU; ~2 min). Tests that need it fail loudly if this fails."""
    from aegolius_b200 import cabi
    if not os.path.exists(cabi.LIB_PATH):
        try:
            from aegolius_b200 import build
            build.build(verbose=False)
        except Exception as e:  # pragma: no cover
            print(f"[conftest] could not build libaegolius_b200.so: {e}")
    # Safety net for the kernel cache (git-ignored build output, normally shipped with the tree by __graft_entry__.build()):
    # on a GPU box whose snapshot lacks it, build it once here with that box's nvcc instead of letting every test that
    # expects the program-compiled kernels fail.
    try:
        if cabi.device_count() > 0 and _jit_cache_is_missing():
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import prebuild_jit
            print("[conftest] aegolius_b200/jit/ is empty: building the program-compiled kernels (nvcc) ...")
            prebuild_jit.main(verbose=True)
    except Exception as e:  # pragma: no cover
        print(f"[conftest] could not prebuild the kernel cache: {e}")


def _jit_cache_is_missing():
    d = os.path.join(ROOT, "aegolius_b200", "jit")
    return not os.path.isdir(d) or sum(f.endswith(".so") for f in os.listdir(d)) < 100


def has_cuda():
    try:
        from aegolius_b200 import cabi
        return cabi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    ref = os.path.isdir("/root/reference/Code/spomso")
    for item in items:
        if "reference" in item.keywords and not ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present (GPU box)"))


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "scenarios.npz")
    return np.load(path, allow_pickle=False)


def golden_names():
    path = os.path.join(ROOT, "tests", "golden", "scenarios.npz")
    with np.load(path, allow_pickle=False) as d:
        return [str(n) for n in d["__names__"]]


def load_case(golden, name):
    from aegolius_b200.program import Program
    keys = {k[len(name) + 1:]: golden[k] for k in golden.files if k.startswith(name + "/")}
    prog = Program.from_arrays(keys, prefix="prog_")
    return dict(prog=prog, expected=keys["expected"], size=tuple(float(s) for s in keys["size"]),
                res=tuple(int(r) for r in keys["res"]), extent=float(keys["extent"]))


# golden scenarios whose field + gradient kernels are prebuilt (tools/prebuild_jit.py) and checked on point lists
# (tests/test_gpu_jit.py): every coordinate op with a gradient pull-back (csrc/ab_adjoint.cuh), the primitives with a
# written-out local gradient, and the combine / value ops on top of them
GRADIENT_CASES = (
    "c1_sphere_box_smooth_union", "c3_deep_tree", "mod_twist", "mod_bend", "prim3_torus", "comb_SMOOTH_INTERSECT2_BOLTZMANN",
    "struct_extruded_combo", "prim3_sphere", "prim3_box", "prim3_cylinder", "prim3_cone", "mod_elongation", "mod_revolution",
    "mod_axis_revolution", "mod_extrusion", "mod_shear_generic", "mod_infinite_repetition", "mod_finite_repetition_rescaled",
    "mod_symmetry", "mod_mirror", "mod_rotational_symmetry", "mod_linear_instancing", "mod_curve_instancing",
    "mod_aligned_curve_instancing", "mod_fully_aligned_curve_instancing", "mod_scale_sdf", "mod_rotate_sdf",
    "mod_chain_twist_elong_round", "mod_onion", "comb_UNION_3", "comb_SMOOTH_SUBTRACT2", "struct_plate_revolved_union",
    "struct_deep_combines", "struct_2d_mirror_rotsym", "ex_pawn_3D", "ex_rod_3D", "ex_chip_3D", "ex_pointcloud_terrain_3D",
)
