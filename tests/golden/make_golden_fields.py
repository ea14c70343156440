"""Generates tests/golden/fields.npz by running the UNMODIFIED reference (SPOMSO, /root/reference/Code/spomso) on
seeded inputs: the grid stencils of post_processing.py and a VectorFieldFromSDF pipeline with every vector modifier.
Run in the build container only (the reference does not travel):  python tests/golden/make_golden_fields.py"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/Code/spomso")
from spomso.cores import post_processing as pp  # noqa: E402
from spomso.cores.geom_vector import VectorFieldFromSDF  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from field_cases import AVG_CASES, EDGE_CASES, RES, vector_inputs, pipelines  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(0)
    for k, (shape, ks, it) in enumerate(AVG_CASES):
        u = rng.normal(size=shape)
        out[f"avg{k}_in"] = u
        out[f"avg{k}_out"] = pp.conv_averaging(u, ks, it)
    for k, shape in enumerate(EDGE_CASES):
        u = rng.normal(size=shape)
        out[f"edge{k}_in"] = u
        out[f"edge{k}_out"] = pp.conv_edge_detection(u)
    inp = vector_inputs()
    for name, ops in pipelines(inp).items():
        vf = VectorFieldFromSDF(RES)
        for op, *a in ops:
            getattr(vf, op)(*a)
        out[f"vec_{name}"] = vf.create(inp["sdf"])
        for comp in ("x", "y", "z", "phi", "theta", "length"):
            with np.errstate(invalid="ignore"):
                out[f"vec_{name}_{comp}"] = getattr(vf, comp)(inp["sdf"])
    out["vec_plain"] = VectorFieldFromSDF(RES).create(inp["sdf"])
    import scipy
    out["versions"] = np.array([np.__version__, scipy.__version__])
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **out)
    print("wrote fields.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
