"""The five BASELINE.json configurations made concrete. Each builder takes a namespace exposing
SPOMSO's class names (aegolius_b200.frontend by default; the real spomso.cores in the golden generator), so the
same construction code drives the reference and this package."""
from __future__ import annotations

import numpy as np


def default_namespace():
    from . import frontend
    return frontend


def spiral(t, radius, height, freq):  # Code/examples/scalar/3D/spiral_instancing_3D.py:16-22
    x = radius * np.cos(2 * np.pi * freq * t)
    y = radius * np.sin(2 * np.pi * freq * t)
    z = height * t - height / 2
    return np.asarray((x, y, z))


def build_c1(ns=None):
    """C1: sphere (+) box smooth union. Grid (4,4,4) / 128^3 -> 129^3."""
    ns = ns or default_namespace()
    s = ns.Sphere(1.0)
    s.move((0.5, 0, 0))
    b = ns.Box(1.5, 1.0, 0.8)
    b.rotate(np.pi / 5, (0, 0, 1))
    b.move((-0.4, 0.2, 0.1))
    return ns.CombineGeometry("SMOOTH_UNION2").combine_parametric(s, b, parameters=0.3)


def build_c2(ns=None):
    """C2: 20 transformed 2D shapes folded through all 13 combine ops. Grid (8,8) / 4096^2 -> 4097^2."""
    ns = ns or default_namespace()
    rng = np.random.default_rng(0)

    def _rounded(o, r):
        o.rounding(r)
        return o

    makers = [
        lambda: ns.Circle(0.8), lambda: ns.Rectangle(1.4, 0.9),
        lambda: ns.RoundedRectangle(1.5, 1.0, (0.1, 0.2, 0.3, 0.15)), lambda: ns.NGon(0.8, 5),
        lambda: ns.Triangle((-0.6, -0.4), (0.7, -0.3), (0.1, 0.8)), lambda: ns.Sector(0.9, 0.3, 2.0),
        lambda: _rounded(ns.Arc(0.7, 0.2, 2.4), 0.1),
    ]
    counter = [0]

    def fresh():
        o = makers[counter[0] % len(makers)]()
        counter[0] += 1
        o.rotate(float(rng.uniform(0, 2 * np.pi)), (0, 0, 1))
        o.rescale(float(rng.uniform(0.6, 1.6)))
        t = rng.uniform(-2.5, 2.5, size=2)
        o.move((float(t[0]), float(t[1]), 0.0))
        return o

    nonparam = ["UNION2", "UNION", "SUBTRACT2", "INTERSECT2", "INTERSECT", "SUM", "DIFFERENCE"]
    param = ["SMOOTH_UNION2_2", "SMOOTH_UNION2", "SMOOTH_INTERSECT2", "SMOOTH_INTERSECT2_BOLTZMANN",
             "SMOOTH_SUBTRACT2", "SMOOTH_SUBTRACT2_BOLTZMANN"]
    acc = fresh()
    for op in nonparam:
        if op in ("UNION2", "UNION"):
            acc = ns.CombineGeometry(op).combine(acc, fresh())
        else:
            side = ns.CombineGeometry(op).combine(fresh(), fresh())
            acc = ns.CombineGeometry("UNION2").combine(acc, side)
    for op in param:
        w = float(rng.uniform(0.2, 0.5))
        if op in ("SMOOTH_UNION2_2", "SMOOTH_UNION2"):
            acc = ns.CombineGeometry(op).combine_parametric(acc, fresh(), parameters=w)
        else:
            side = ns.CombineGeometry(op).combine_parametric(fresh(), fresh(), parameters=w)
            acc = ns.CombineGeometry("UNION2").combine(acc, side)
    return acc


def build_c3(ns=None):
    """C3 / C5: deep tree (elongation, twist, bend, rotational symmetry, fully aligned curve instancing x21, smooth
    union with an onion'd sphere, rounding, mirror). Grid (6,6,6) / 512^3 -> 513^3 (C5: 1024^3 -> 1025^3)."""
    ns = ns or default_namespace()
    t = ns.Torus(0.25, 0.2)
    t.elongation((2, 0, 0))
    t.rotate(np.pi / 2, (0, 1, 0))
    g = ns.GenericGeometry(t.propagate, ())
    g.twist(np.pi)
    g.bend(2.0, 1.0)
    g.rotational_symmetry(6, 1.5, 0.1)
    g.fully_aligned_curve_instancing(spiral, (1, 2, 2), (0, 1, 21))
    s = ns.Sphere(0.4)
    s.move((0.2, 0, 0))
    s.onion(0.05)
    u = ns.CombineGeometry("SMOOTH_UNION2").combine_parametric(g, s, parameters=0.3)
    u.rounding(0.01)
    u.mirror((-1, 0, 0), (1, 0, 0))
    return u


def c4_cloud(m=1_000_000, seed=0):
    """C4: terrain-like surface z = 0.2 sin(3x) cos(2y) + 0.01 N(0,1), x,y ~ U(-1,1); returns (3, m) float64."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, m)
    y = rng.uniform(-1, 1, m)
    z = 0.2 * np.sin(3 * x) * np.cos(2 * y) + 0.01 * rng.normal(size=m)
    return np.stack([x, y, z])


CONFIGS = {
    "C1": dict(build=build_c1, size=(4.0, 4.0, 4.0), res=(128, 128, 128)),
    "C2": dict(build=build_c2, size=(8.0, 8.0), res=(4096, 4096)),
    "C3": dict(build=build_c3, size=(6.0, 6.0, 6.0), res=(512, 512, 512)),
    "C4": dict(cloud=c4_cloud, size=(2.5, 2.5, 1.5), res=(256, 256, 256)),
    "C5": dict(build=build_c3, size=(6.0, 6.0, 6.0), res=(1024, 1024, 1024)),
}
