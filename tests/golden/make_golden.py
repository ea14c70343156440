"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (SPOMSO, NumPy path) in the build container.

    PYTHONPATH=/root/reference/Code/spomso:/root/repo python tests/golden/make_golden.py

For every scenario in tests/scenarios.py: build the object with the real SPOMSO classes, evaluate obj.create(co)
on SPOMSO's own generate_grid coordinates (the expected field), flatten the SAME object through closure
introspection (aegolius_b200.introspect) and store program + expected output. The GPU box has no SPOMSO, so these
fixtures are what pins both the oracle and the CUDA path to the reference there.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/Code/spomso")

import scipy  # noqa: E402
from spomso.cores.helper_functions import generate_grid  # noqa: E402

from aegolius_b200.program import flatten  # noqa: E402
from scenarios import SCENARIOS, make_namespace  # noqa: E402


def main():
    ns = make_namespace("reference")
    out = {}
    meta = []
    for name, sc in SCENARIOS.items():
        co, res = generate_grid(sc["size"], sc["res"])
        obj = sc["build"](ns)
        expected = np.asarray(obj.create(co.copy())).reshape(-1)  # conv_edge_detection returns the N-D array
        prog = flatten(obj)
        out.update(prog.to_arrays(prefix=f"{name}/prog_"))
        out[f"{name}/expected"] = expected
        out[f"{name}/size"] = np.asarray(list(sc["size"]) + ([0.0] if len(sc["size"]) == 2 else []))
        r = (res[0], res[1], res[2] if len(sc["size"]) == 3 else 1)
        out[f"{name}/res"] = np.asarray(r, dtype=np.int64)
        out[f"{name}/extent"] = np.asarray(sc["extent"], dtype=np.float64)
        meta.append(name)
        print(f"{name:45s} N={expected.size:6d} ops={prog.n_ops:3d} args={prog.args.size:4d} "
              f"range=[{np.nanmin(expected):+.3f},{np.nanmax(expected):+.3f}] nan={int(np.isnan(expected).sum())}")
    out["__names__"] = np.asarray(meta)
    out["__versions__"] = np.asarray([f"numpy {np.__version__}", f"scipy {scipy.__version__}", "spomso 1.4.0"])
    np.savez_compressed(os.path.join(HERE, "scenarios.npz"), **out)
    print("wrote", os.path.join(HERE, "scenarios.npz"), os.path.getsize(os.path.join(HERE, "scenarios.npz")), "bytes")


if __name__ == "__main__":
    main()
