"""Closure introspection: an UNMODIFIED SPOMSO object -> the explicit tree of frontend.py.

SPOMSO has no tree data structure: every modification wraps the previous SDF callable in a closure
(Code/spomso/spomso/cores/modifications.py:88-98) and every combine closes over its children
(Code/spomso/spomso/cores/combine.py:129-138). The structure is fully recoverable:
  * `fn.__qualname__`  = "ModifyObject.<name>.<locals>.new_geo_object" names the modification,
  * `fn.__code__.co_freevars` + `fn.__closure__` hold its parameters and the wrapped callable (`geo_object`),
  * a combine closure holds `combined_objects` and `self.operation_type` (+ `parameters`),
  * a bound `GenericGeometry.propagate` marks a nested node (Code/examples/scalar/3D/basics_3D.py:111),
  * a module-level `sdf_*` function is a leaf whose parameters sit in `node._geo_parameters` (geom.py:26).
The parameter names captured by the reference closures are exactly the keys frontend.ModifyObject records.
"""
from __future__ import annotations

import inspect

import numpy as np

from . import frontend as fe

_RENAME = {"rep": "rep"}


def _cells(fn):
    if fn.__closure__ is None:
        return {}
    out = {}
    for name, cell in zip(fn.__code__.co_freevars, fn.__closure__):
        try:
            out[name] = cell.cell_contents
        except ValueError:  # empty cell
            pass
    return out


def _blank(kind):
    n = fe.GenericGeometry.__new__(fe.GenericGeometry)
    fe.EuclideanTransform.__init__(n)
    fe.ModifyObject.__init__(n)
    n._geo_parameters = ()
    n.kind = kind
    n.leaf = n.inner = n.combine_op = n.children = n.combine_parameter = None
    return n


def to_frontend(obj, _stack=()):
    """Converts a SPOMSO GenericGeometry (any subclass) into a frontend.GenericGeometry tree."""
    if isinstance(obj, fe.GenericGeometry):
        return obj
    if any(obj is s for s in _stack):
        raise NotImplementedError("geometry tree contains a cycle")
    _stack = _stack + (obj,)
    for attr in ("rotation_matrix", "center", "scale", "geo_object"):
        if not hasattr(obj, attr):
            raise NotImplementedError(f"{type(obj).__name__} is not a SPOMSO geometry object (no .{attr})")

    mods_outer_first = []
    fn = obj.geo_object
    node = None
    while node is None:
        qn = getattr(fn, "__qualname__", "")
        if inspect.ismethod(fn) and fn.__func__.__name__ in ("propagate", "create"):
            node = _blank("nested")
            node.inner = to_frontend(fn.__self__, _stack)
        elif qn.startswith("ModifyObject.") and qn.endswith(".<locals>.new_geo_object"):
            name = qn.split(".")[1]
            cells = _cells(fn)
            if "geo_object" not in cells:
                raise NotImplementedError(f"closure {qn} has no wrapped geo_object")
            params = {k: v for k, v in cells.items() if k not in ("geo_object", "self")}
            if name not in fe.ModifyObject.__dict__ or name in (
                    "custom_modification", "custom_post_process", "displacement", "define_volume",
                    "recover_volume", "signed_old"):
                raise NotImplementedError(
                    f"modification '{name}' (closure {qn}) takes a Python callable or a grid stencil and cannot "
                    f"enter the GPU op list; evaluate this object with SPOMSO itself")
            if name == "elongation":
                params = {"ev": params["ev"]}
            mods_outer_first.append((name, params))
            fn = cells["geo_object"]
        elif qn.startswith("CombineGeometry.combine") and qn.endswith(".<locals>.new_geo_object"):
            cells = _cells(fn)
            node = _blank("combine")
            node.combine_op = cells["self"].operation_type
            node.children = tuple(to_frontend(c, _stack) for c in cells["combined_objects"])
            node.combine_parameter = cells.get("parameters")
        elif qn in ("SegmentedLine.polygon.<locals>.new_geo_object", "SegmentedParametricCurve.polygon.<locals>.new_geo_object"):
            cells = _cells(fn)  # geom_2d.py:530-555, 601-626
            mods_outer_first.append(("polygon", {"points": np.asarray(cells["self"]._points).copy()}))
            fn = cells["geo_object"]
        elif qn == "ParametricCurve.shape.<locals>.new_geo_object":
            cells = _cells(fn)  # geom_2d.py:415-457
            src = cells["self"]
            ts_ = np.zeros(src.steps + 1)
            ts_[:src.steps] = src.ts
            mods_outer_first.append(("shape", {"points": np.asarray(src._curve(ts_, *src._c_params), dtype=np.float64)}))
            fn = cells["geo_object"]
        elif qn in ("SegmentedLine.sdf_closed_curve.<locals>.new_geo_object",
                    "SegmentedLine3D.sdf_closed_curve.<locals>.new_geo_object"):
            node = _blank("leaf")
            node.leaf = "sdf_closed_segmented_line_2d" if qn.startswith("SegmentedLine.") else \
                "sdf_closed_segmented_line_3d"
            node._geo_parameters = tuple(getattr(obj, "_geo_parameters", ()))
        elif qn in ("ParametricCurve.sdf_closed_curve.<locals>.new_geo_object",
                    "ParametricCurve3D.sdf_closed_curve.<locals>.new_geo_object"):
            node = _blank("leaf")
            node.leaf = "sdf_closed_parametric_curve_2d" if qn.startswith("ParametricCurve.") else \
                "sdf_closed_parametric_curve_3d"
            node._geo_parameters = tuple(getattr(obj, "_geo_parameters", ()))
        elif qn in ("SegmentedParametricCurve.sdf_closed_curve.<locals>.new_geo_object",
                    "SegmentedParametricCurve3D.sdf_closed_curve.<locals>.new_geo_object"):
            node = _blank("leaf")
            node.leaf = "sdf_closed_segmented_curve_3d" if qn.startswith("SegmentedParametricCurve3D.") else \
                "sdf_closed_segmented_curve_2d"
            node._geo_parameters = tuple(getattr(obj, "_geo_parameters", ()))
        elif inspect.isfunction(fn) and fn.__name__ in fe.LEAVES and "<locals>" not in qn:
            node = _blank("leaf")
            node.leaf = fn.__name__
            params = getattr(obj, "_geo_parameters", None)
            if params is None:
                params = getattr(obj, "geo_parameters", ())
            node._geo_parameters = tuple(params)
        else:
            raise NotImplementedError(
                f"SDF callable {qn or fn!r} is not a known SPOMSO primitive / modification / combine closure and "
                f"cannot enter the GPU op list; evaluate this object with SPOMSO itself")

    node._rot_matrix = np.array(obj.rotation_matrix, dtype=np.float64)
    node._center = np.array(obj.center, dtype=np.float64)
    node._scale = obj.scale
    node._mods = list(reversed(mods_outer_first))
    node._mod = [m[0] for m in node._mods]
    return node
