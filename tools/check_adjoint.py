#!/usr/bin/env python
"""Gradient kernels built with pull-backs (csrc/ab_adjoint.cuh) against the interpreter's forward-mode gradient, for every
golden scenario without a grid stage, on a ragged list of random points, fp64 and fp32.

  python tools/check_adjoint.py --prebuild     # here (nvcc, no GPU): builds the kernels into aegolius_b200/jit/
  python tools/check_adjoint.py                # on the GPU: one JSON line per scenario + a summary line
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def scenarios():
    from aegolius_b200.program import Program
    path = os.path.join(ROOT, "tests", "golden", "scenarios.npz")
    with np.load(path, allow_pickle=False) as d:
        for name in [str(n) for n in d["__names__"]]:
            keys = {k[len(name) + 1:]: d[k] for k in d.files if k.startswith(name + "/")}
            prog = Program.from_arrays(keys, prefix="prog_")
            if prog.stages:
                continue
            yield name, prog, np.asarray(keys["size"], dtype=np.float64), float(keys["extent"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prebuild", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    from aegolius_b200 import codegen
    todo = [s for s in scenarios() if not a.only or a.only in s[0]]
    if a.prebuild:
        items = []
        for name, prog, size, ext in todo:
            for dt in ("f32", "f64"):
                items.append((prog, dt, "spatial", False))
        print(codegen.prebuild(items, verbose=True))
        return
    import aegolius_b200 as ab
    from aegolius_b200 import cabi
    from oracle import interp_np
    lib = cabi.lib()
    worst = {"f64": 0.0, "f32": 0.0}
    n_adj = 0
    for name, prog, size, ext in todo:
        rng = np.random.default_rng(11)
        size3 = np.ones(3)
        size3[:len(size)] = size
        co = rng.uniform(-0.45, 0.45, size=(3, 10007)) * size3.reshape(3, 1)
        sig = codegen.signature(prog)
        adjoint = "ab_adjoint.cuh" in codegen.generate(sig, "f64", "spatial") and "pb_" in codegen.generate(sig, "f64", "spatial").split("namespace ab {")[1]
        try:
            _, margin = interp_np.run(prog, co, return_margin=True)
        except Exception as exc:  # programs the point-list oracle does not serve
            print(json.dumps({"case": name, "skipped": str(exc)[:80]}), flush=True)
            continue
        keep = margin > 1e-4 * ext
        line = {"case": name, "ops": len(sig), "pullbacks": adjoint, "kept": float(keep.mean())}
        for dt in ("f64", "f32"):
            h0 = lib.ab_prog_hits()
            fa, ga = ab.create(prog, co, dtype=dt, grad="spatial")
            hits = lib.ab_prog_hits() - h0
            old = lib.ab_prog_enable(0)
            try:
                fb, gb = ab.create(prog, co, dtype=dt, grad="spatial")
            finally:
                lib.ab_prog_enable(old)
            dv = np.abs(fa.astype(np.float64) - fb)[keep]
            dg = np.abs(ga.astype(np.float64) - gb)[:, keep]
            ok = np.isfinite(dg)
            line[dt] = {"compiled": int(hits), "value_max": float(np.nanmax(dv)) if dv.size else 0.0,
                        "grad_max": float(dg[ok].max()) if ok.any() else 0.0,
                        "grad_scale": float(np.nanmax(np.abs(gb[:, keep]))) if keep.any() else 0.0}
            worst[dt] = max(worst[dt], line[dt]["grad_max"])
        n_adj += adjoint
        print(json.dumps(line), flush=True)
    print(json.dumps({"summary": True, "scenarios": len(todo), "with_pullbacks": n_adj, "worst_grad_diff": worst}))


if __name__ == "__main__":
    main()
