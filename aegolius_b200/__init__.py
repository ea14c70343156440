"""aegolius_b200 — B200-native (sm_100a) evaluator for SPOMSO's composed-SDF hot path.

Front end: the same object / functional API as SPOMSO (frontend.py mirrors it; introspect.py accepts unmodified
SPOMSO objects). Back end: a flattened op list interpreted by a hand-written CUDA kernel behind a C ABI
(include/aegolius_b200.h, csrc/). There is no CPU evaluation path in this package.
"""
from .grid import (GridSpec, GridCoords, generate_grid, generate_grid_spec, resolution_conversion,  # noqa: F401
                   smarter_reshape, detect_grid)
from .frontend import *  # noqa: F401,F403
from .frontend import LEAVES, LeafSDF, GenericGeometry, CombineGeometry  # noqa: F401
from .program import Program, flatten, FlattenError  # noqa: F401
from .introspect import to_frontend  # noqa: F401
from . import workloads  # noqa: F401
from . import codegen  # noqa: F401


def __getattr__(name):
    # engine (ctypes + CUDA library) is imported lazily so that flattening works without the built library
    if name in ("create", "create_with_gradient", "patch", "unpatch", "from_sdf", "VectorFieldFromSDF",
                "point_cloud_sdf", "conv_averaging", "conv_edge_detection", "specialize", "set_auto_specialize", "compile_program", "wait_for_compilations", "from_sdf_torch", "cloud_records", "point_cloud_sdf_torch", "grid_min_step", "wait_for_specializations", "engine", "create_torch", "jacfwd", "value_and_grad", "program_tangent", "library_path", "PinnedArray", "slab_ranges"):
        import importlib
        engine = importlib.import_module(__name__ + ".engine")
        return engine if name == "engine" else getattr(engine, name)
    raise AttributeError(name)


__version__ = "0.1.0"
