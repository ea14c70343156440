"""Seeded inputs shared by tests/golden/make_golden_fields.py (runs the reference) and the stencil / vector-field tests."""
import numpy as np

AVG_CASES = [((9, 7, 5), (5, 5, 1), 1), ((9, 7, 5), (3, 3, 3), 2), ((8, 6), (2, 4), 1), ((7, 9, 4), (4, 3, 2), 3),
             ((5, 5), (7, 7), 1), ((6, 5, 3), (2, 2, 1), 1), ((6, 5, 1), (3, 2, 4), 2)]
EDGE_CASES = [(9, 7), (6, 8, 3), (2, 2), (3, 3, 1)]
RES = (11, 9, 7)


def vector_inputs():
    rng = np.random.default_rng(42)
    n = int(np.prod(RES))
    ii = np.indices(RES).reshape(3, -1).astype(np.float64)
    sdf = np.sqrt(((ii - np.array([[5.2], [3.9], [3.1]])) ** 2).sum(axis=0)) - 3.0 + 0.05 * rng.normal(size=n)
    sdf[:RES[1] * RES[2]] = 1.0  # a flat plane: zero gradient -> zero vectors that must stay zero
    sdf[RES[1] * RES[2]:2 * RES[1] * RES[2]] = 1.0
    return dict(sdf=sdf, phis=rng.uniform(-3, 3, n), thetas=rng.uniform(-1, 1, n), alpha=rng.uniform(-3, 3, n),
                axes=rng.normal(size=(3, n)), co=rng.normal(size=(3, n)), other=rng.normal(size=(3, n)),
                scale=rng.uniform(0.5, 2, n))


def pipelines(i):
    """Modifier lists in call order: (method name, operands...). Shared with the tests."""
    return {
        "example": [("rotate_z", i["phis"]), ("rotate_axis", (1, 0, 0), i["thetas"])],  # sdf_vector_field.py:155-162
        "everything": [("add", (0.1, -0.2, 0.3)), ("rotate_phi", i["phis"]), ("rotate_theta", i["thetas"]),
                       ("subtract", i["other"]), ("rescale", i["scale"]), ("rotate_x", 0.4), ("rotate_y", i["alpha"]),
                       ("rotate_z", -1.1), ("rotate_axis", i["axes"], 0.7), ("revolution_x", i["co"]),
                       ("revolution_y", i["co"]), ("revolution_z", i["co"]), ("add", 0.25), ("normalize",),
                       ("rescale", 1.5), ("subtract", (0.0, 0.0, 0.2))],
    }
