"""The C-ABI contract without a GPU: include/aegolius_b200.h, opcodes.py and the built library agree."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "aegolius_b200.h")).read()


def test_opcode_numbers_match_the_header():
    from aegolius_b200 import opcodes as oc
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"AB_OP_([A-Z0-9_]+)\s*=\s*(\d+)", HEADER))
    assert enum.pop("_COUNT") == oc.OP_COUNT
    py = {k: v for k, v in vars(oc).items() if k.isupper() and isinstance(v, int) and k in enum}
    assert py == enum and len(enum) > 70
    for name, val in (("AB_MAX_OPS", oc.MAX_OPS), ("AB_MAX_ARGS", oc.MAX_ARGS), ("AB_MAX_PSLOTS", oc.MAX_PSLOTS),
                      ("AB_MAX_VSLOTS", oc.MAX_VSLOTS), ("AB_MAX_BLOBS", oc.MAX_BLOBS)):
        assert int(re.search(rf"#define {name} (\d+)", HEADER).group(1)) == val


def test_library_loads_and_exports_every_declared_symbol():
    from aegolius_b200 import cabi
    lib = cabi.lib()
    declared = set(re.findall(r"\b(ab_[a-z0-9_]+)\s*\(", HEADER))
    declared -= {"ab_program", "ab_grid", "ab_op", "ab_blob"}
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ab_version() == int(re.search(r"#define AB_VERSION (\d+)", HEADER).group(1))


def test_struct_layouts_match_the_header():
    from aegolius_b200 import cabi
    from aegolius_b200.program import OP_DTYPE
    assert C.sizeof(cabi.ab_op) == 8 == OP_DTYPE.itemsize
    assert C.sizeof(cabi.ab_grid) == 3 * 8 + 3 * 4 + 2 * 4 + 4  # doubles, res, slab, tail padding to 8
    assert C.sizeof(cabi.ab_blob) == 24
    assert cabi.ab_program.ops.offset == 0 and cabi.ab_program.n_ops.offset == 8


def test_without_a_gpu_every_compute_call_fails_loudly():
    import numpy as np
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    if cabi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cabi.NoDeviceError):
        ab.create(ab.Sphere(1.0), ab.GridSpec((2, 2, 2), (4, 4, 4)))
    with pytest.raises(cabi.NoDeviceError):
        ab.from_sdf(np.zeros(27), (3, 3, 3))
    with pytest.raises(cabi.NoDeviceError):
        ab.point_cloud_sdf(ab.GridSpec((2, 2, 2), (4, 4, 4)), np.zeros((3, 5)))


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "aegolius_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "interp_np" not in text, f"{f} references the oracle"
