"""Grid helpers mirroring Code/spomso/spomso/cores/helper_functions.py:10-148.

`generate_grid` returns a `GridCoords` — an ndarray subclass that is bit-identical to what SPOMSO's generate_grid
returns but also carries (size, resolution), so the engine can regenerate coordinates in-kernel instead of reading
24 B/point. For grids too large to materialise on the host (1025^3 = 25.8 GB of fp64) use `GridSpec`, which
carries the same description without allocating; the engine accepts either.
"""
from __future__ import annotations

import numpy as np


def resolution_conversion(resolution) -> int:
    """helper_functions.py:10-20: resolutions are forced odd."""
    return int(resolution if resolution % 2 == 1 else resolution + 1)


def _convert(size, resolution):
    resolution = np.asarray(resolution)
    size = np.asarray(size, dtype=np.float64).reshape(-1)
    if resolution.size == 1:
        r = resolution_conversion(int(resolution.reshape(-1)[0]))
        res = (r, r, r)
    elif resolution.size == 2:
        r0 = resolution_conversion(int(resolution[0]))
        res = (r0, resolution_conversion(int(resolution[1])), r0)
    elif resolution.size == 3:
        res = tuple(resolution_conversion(int(r)) for r in resolution)
    else:
        raise ValueError("resolution must have 1, 2 or 3 entries")
    if size.size not in (2, 3):
        # the reference's 1-D branch indexes a 0-d array and raises (helper_functions.py:56-61)
        raise IndexError("too many indices for array: size must have 2 or 3 entries")
    return size, res


class GridSpec:
    """Description of a generate_grid grid without the (3,N) array. dims = 2 or 3."""

    def __init__(self, size, resolution):
        size, res = _convert(size, resolution)
        self.dims = int(size.size)
        self.size = tuple(float(s) for s in size) + ((0.0,) if size.size == 2 else ())
        self.co_resolution = res  # what generate_grid returns as its second value
        self.res = (res[0], res[1], res[2] if self.dims == 3 else 1)

    @property
    def n_points(self):
        return self.res[0] * self.res[1] * self.res[2]

    @property
    def extent(self):
        return max(self.size)

    @property
    def shape(self):
        return (3, self.n_points)

    def axes(self):
        return [np.linspace(-self.size[i] / 2, self.size[i] / 2, self.res[i]) for i in range(self.dims)]

    def slab_coords(self, x0, x1):
        """fp64 (3, n) coordinates of the ix planes [x0, x1): identical values to generate_grid's rows."""
        ax = self.axes()
        if self.dims == 3:
            co = np.asarray(np.meshgrid(ax[0][x0:x1], ax[1], ax[2], indexing="ij")).reshape(3, -1)
        else:
            c2 = np.asarray(np.meshgrid(ax[0][x0:x1], ax[1], indexing="ij")).reshape(2, -1)
            co = np.zeros((3, c2.shape[1]))
            co[:2] = c2
        return co

    def materialize(self):
        return self.slab_coords(0, self.res[0])

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        return a if dtype is None else a.astype(dtype)


class GridCoords(np.ndarray):
    """(3,N) float64 coordinates that remember the grid they came from."""

    def __new__(cls, array, spec):
        obj = np.asarray(array).view(cls)
        obj.spec = spec
        return obj

    def __array_finalize__(self, obj):
        # any derived array (slice, arithmetic result) is no longer known to be the full regular grid
        self.spec = None


def generate_grid(size, resolution):
    """helper_functions.py:23-93. Returns (coordinates (3,N) float64, converted resolution tuple)."""
    spec = GridSpec(size, resolution)
    return GridCoords(spec.materialize(), spec), spec.co_resolution


def generate_grid_spec(size, resolution):
    """Same arguments as generate_grid, but returns (GridSpec, converted resolution) without allocating."""
    spec = GridSpec(size, resolution)
    return spec, spec.co_resolution


def detect_grid(co):
    """Recognises a (3,N) array produced by SPOMSO's own generate_grid (plain ndarray): returns a GridSpec or None.
    Arrays that carry their provenance (GridSpec, GridCoords.spec) are taken at their word. For a plain array the grid is
    proposed from O(nx+ny+nz) entries and then EVERY coordinate is compared with the proposed axes (one pass over the
    array, chunked: about what the host-to-device upload it replaces would cost) — in grid mode the kernel regenerates the
    coordinates and never reads the array, so a generate_grid output with a few edited points must not pass."""
    if isinstance(co, GridSpec):
        return co
    spec = getattr(co, "spec", None)
    if isinstance(spec, GridSpec):
        return spec
    co = np.asarray(co)
    if co.ndim != 2 or co.shape[0] != 3 or co.dtype != np.float64 or co.shape[1] < 8:
        return None
    n = co.shape[1]
    z0 = co[2, 0]
    if np.all(co[2, :min(n, 4)] == 0.0) and z0 == 0.0 and not np.any(co[2, ::max(1, n // 1024)]):
        nz = 1  # 2D grid: zero z row (helper_functions.py:72-75)
    else:
        # z is the fastest axis: first index where y changes
        lim = min(n, 1 << 16)
        ch = np.nonzero(co[1, :lim] != co[1, 0])[0]
        if ch.size == 0:
            return None
        nz = int(ch[0])
    ych = co[0, ::nz]
    c2 = np.nonzero(ych[:min(ych.size, 1 << 16)] != ych[0])[0]
    if c2.size == 0:
        return None
    ny = int(c2[0])
    if nz < 1 or ny < 2 or n % (ny * nz):
        return None
    nx = n // (ny * nz)
    if nx < 2:
        return None
    if nz == 1:
        size = (-2 * co[0, 0], -2 * co[1, 0])
        res = (nx, ny)
    else:
        size = (-2 * co[0, 0], -2 * co[1, 0], -2 * co[2, 0])
        res = (nx, ny, nz)
    if any(r % 2 == 0 for r in res) or any(not (s > 0) for s in size):
        return None
    spec = GridSpec(size, res)
    ax = spec.axes()
    if not (co[0, -1] == ax[0][-1] and co[1, -1] == ax[1][-1]):
        return None
    # full verification, x-chunks of about 2^22 points: co[0] is constant over each (ny, nz) plane, co[1] over each z row
    plane = ny * nz
    step = max(1, (1 << 22) // plane)
    for x0 in range(0, nx, step):
        x1 = min(nx, x0 + step)
        sl = slice(x0 * plane, x1 * plane)
        if not np.array_equal(co[0, sl].reshape(x1 - x0, plane), np.broadcast_to(ax[0][x0:x1, None], (x1 - x0, plane))):
            return None
        if not np.array_equal(co[1, sl].reshape(x1 - x0, ny, nz), np.broadcast_to(ax[1][None, :, None], (x1 - x0, ny, nz))):
            return None
        if nz > 1:
            if not np.array_equal(co[2, sl].reshape(x1 - x0, ny, nz), np.broadcast_to(ax[2][None, None, :], (x1 - x0, ny, nz))):
                return None
        elif np.any(co[2, sl]):
            return None
    return spec


def smarter_reshape(pattern, resolution):
    """helper_functions.py:96-148."""
    pattern = np.asarray(pattern)
    n_ele = pattern.shape[0]
    resolution = np.asarray(resolution)
    if resolution.size == 1:
        res = resolution_conversion(int(resolution.reshape(-1)[0]))
        if n_ele // res == 1:
            return pattern
        if n_ele // (res ** 2) == 1:
            return pattern.reshape(res, res)
        if n_ele // (res ** 3) == 1:
            return pattern.reshape(res, res, res)
        raise ValueError(f"Cannot reshape the pattern with shape {pattern.shape}")
    if resolution.size == 2:
        r0, r1 = (resolution_conversion(int(r)) for r in resolution)
        div = n_ele // (r0 * r1)
        if div == 1:
            return pattern.reshape(r0, r1)
        return pattern.reshape(r0, r1, int(div))
    if resolution.size == 3:
        r0, r1, r2 = (resolution_conversion(int(r)) for r in resolution)
        if n_ele // (r0 * r1 * r2) == 1:
            return pattern.reshape(r0, r1, r2)
        raise ValueError(f"Cannot reshape the pattern with shape {pattern.shape}")
    raise ValueError("resolution must have 1, 2 or 3 entries")
