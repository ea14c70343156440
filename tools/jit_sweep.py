#!/usr/bin/env python
"""Sweep of the program-compiled kernels (aegolius_b200/codegen.py) against the interpreter: device time per case and
per option set (points per thread, occupancy asked of the register allocator, slots in registers / shared memory), plus
a bit-for-bit comparison of the outputs.

    python tools/jit_sweep.py --prebuild          # here (no GPU): nvcc every variant into aegolius_b200/jit/
    python tools/jit_sweep.py [--cases ...]       # on the GPU box: times them (never runs nvcc: AB_JIT=cache)
"""
import argparse
import itertools
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases():
    import aegolius_b200 as ab
    from aegolius_b200 import workloads
    sph = ab.Sphere(1.0)
    sph.move((0.3, 0.1, -0.2))
    sph0 = ab.Sphere(1.0)
    # a short program with transcendentals (twist + torus) and a long cheap one (union of 12 moved boxes / spheres)
    tw = ab.Torus(1.0, 0.3)
    tw.twist(1.5)
    tw.move((0.1, -0.2, 0.05))
    parts = []
    for i in range(12):
        g = ab.Sphere(0.3 + 0.02 * i) if i % 2 else ab.Box(0.4, 0.3 + 0.02 * i, 0.5)
        g.move((-1.5 + 0.27 * i, 0.3 * ((i * 5) % 7) - 0.9, 0.25 * ((i * 3) % 5) - 0.5))
        parts.append(g)
    un = ab.CombineGeometry("UNION").combine(*parts)
    return {
        "sphereg": (sph, ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", "spatial"),
        "twistg": (tw, ab.GridSpec((4, 4, 4), (768,) * 3), "f32", "spatial"),
        "union12g": (un, ab.GridSpec((4, 4, 4), (768,) * 3), "f32", "spatial"),
        "twist": (tw, ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "union12": (un, ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "sphere0": (sph0, ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "sphere": (sph, ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "c1": (workloads.build_c1(), ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "c1g": (workloads.build_c1(), ab.GridSpec((4, 4, 4), (768,) * 3), "f32", "spatial"),
        "c2": (workloads.build_c2(), ab.GridSpec((8, 8), (4096, 4096)), "f32", None),
        "c3": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (512,) * 3), "f32", None),
        "c3g": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (512,) * 3), "f32", "spatial"),
        "c5g": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (1024,) * 3), "f32", "spatial"),
        "c1_64": (workloads.build_c1(), ab.GridSpec((4, 4, 4), (768,) * 3), "f64", None),
        "c3_64": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (384,) * 3), "f64", None),
        "c3g_64": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (256,) * 3), "f64", "spatial"),
    }


QUICK = {  # narrowed after the first full sweep (profiles/r02_jit_sweep.md)
    "sphere0": [(4, 8), (8, 6), (8, 8)], "sphere": [(4, 6), (4, 8), (8, 6), (8, 8)], "c1": [(4, 8), (8, 6), (8, 8)],
    "c1g": [(4, 3), (4, 4), (2, 4)], "c2": [(4, 5), (4, 6), (8, 5)], "c3": [(4, 5), (8, 5), (8, 6)],
    "c3g": [(2, 5), (2, 6), (2, 7)], "c5g": [(2, 5), (2, 7)], "c1_64": [(2, 4), (2, 5), (2, 6)],
    "c3_64": [(2, 4), (2, 5), (2, 6)], "c3g_64": [(1, 4), (1, 5), (1, 6)],
}


def variants(dtype, grad, lite):
    if dtype == "f64":
        ws = (2,) if grad is None else (1,)
        cs = (3, 4, 5, 6)
    elif grad is None:
        ws = (4, 8)
        cs = (4, 5, 6, 8) if lite else (3, 4, 5, 6, 7)
    else:
        ws = (2, 4) if lite else (2,)
        cs = (2, 3, 4) if lite else (2, 3, 4, 5, 6, 7)
    out = []
    for w, c, sl in itertools.product(ws, cs, ("reg", "smem")):
        out.append(dict(width=w, min_ctas=c, slots=sl))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="sphere0,sphere,c1,c1g,c2,c3,c3g")
    ap.add_argument("--prebuild", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--opts", default=None, help="JSON list of option dicts to use for every case instead of the tables")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    args = ap.parse_args()
    import aegolius_b200 as ab
    from aegolius_b200 import codegen as cg, engine, cabi
    all_cases = cases()
    names = [c for c in args.cases.split(",") if c]
    work = []
    for name in [c for c in names if c != "fill"]:
        obj, spec, dtype, grad = all_cases[name]
        prog = ab.flatten(obj)
        sig = cg.signature(prog)
        lite = all((int(w) & 0xffff) in cg.LITE_OPS for w in sig)
        vs = variants(dtype, grad, lite)
        if args.quick:
            vs = [dict(width=w, min_ctas=c, slots=sl) for (w, c) in QUICK[name] for sl in (("reg", "smem") if name == "c2" else ("reg",))]
        if args.opts:
            vs = json.loads(args.opts)
        for o in vs:
            work.append((name, prog, sig, spec, dtype, grad, o))
    if args.prebuild:
        os.environ["AB_JIT_JOBS"] = str(args.jobs)
        cg._build_slots = __import__("threading").Semaphore(args.jobs)

        def build(item):
            name, prog, sig, spec, dtype, grad, o = item
            src = cg.generate(sig, dtype, cg._GRAD_NAMES[grad], is2d=engine._is_2d(spec), **o)
            try:
                path = cg.build_source(src, keep_ptxas=True)
            except RuntimeError as exc:
                return name, o, None, str(exc)[-400:]
            info = ""
            pt = path[:-3] + ".ptxas.txt"
            if os.path.exists(pt):
                for line in open(pt):
                    if "Used" in line or "spill" in line:
                        info += " " + line.strip().replace("ptxas info    : ", "")
            return name, o, path, info
        with ThreadPoolExecutor(args.jobs) as ex:
            for name, o, path, info in ex.map(build, work):
                print(name, o, os.path.basename(path) if path else "FAILED", info[:150], flush=True)
        return

    import numpy as np
    import torch
    peak = 6452.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    lib = cabi.lib()

    bufs = {}

    def timed(prog, spec, dtype, grad):
        best = 1e30
        n = spec.n_points
        tdt = torch.float32 if dtype == "f32" else torch.float64
        key = (n, dtype, bool(grad))
        if key not in bufs:
            bufs.clear()
            torch.cuda.empty_cache()
            bufs[key] = (torch.empty(n, dtype=tdt, device="cuda"),
                         torch.empty((3, (n + 7) // 8 * 8), dtype=tdt, device="cuda") if grad else None)
        field, gbuf = bufs[key]
        out = None
        for _ in range(args.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = engine.create_torch(prog, spec, dtype=dtype, grad=grad, out=field, out_grad=gbuf)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, out

    if "fill" in names:  # the write-only ceiling: a plain fill of the same 4.3 GB
        names.remove("fill")
        buf = torch.empty(1025 ** 3, dtype=torch.float32, device="cuda")
        best = 1e30
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            buf.fill_(1.5)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"case": "fill 1025^3 f32 (torch fill_)", "ms": round(best, 4), "GBps": round(buf.numel() * 4 / best / 1e6, 1),
                          "hbm_frac": round(buf.numel() * 4 / best / 1e6 / peak, 4)}), flush=True)
        del buf

    os.environ["AB_JIT"] = "off"  # this tool registers the variants itself
    done_interp = {}
    for name, prog, sig, spec, dtype, grad, o in work:
        n = spec.n_points
        bpp = (4 if dtype == "f32" else 8) * (4 if grad else 1)
        if name not in done_interp:
            lib.ab_prog_enable(0)
            ms, ref = timed(prog, spec, dtype, grad)
            ref = [t.clone() for t in (ref if isinstance(ref, tuple) else (ref,))]
            lib.ab_prog_enable(1)
            done_interp[name] = (ms, ref)
            print(json.dumps({"case": name, "kernel": "interpreter", "ms": round(ms, 4), "Gpts_s": round(n / ms / 1e6, 2),
                              "hbm_frac": round(n * bpp / ms / 1e6 / peak, 4)}), flush=True)
        src = cg.generate(sig, dtype, cg._GRAD_NAMES[grad], is2d=engine._is_2d(spec), **o)
        path = cg.binary_path(src)
        if not os.path.exists(path):
            print(json.dumps({"case": name, "opts": o, "error": "not prebuilt"}), flush=True)
            continue
        lib.ab_prog_clear()  # one variant at a time (a compact-tile build would otherwise stay preferred)
        cg._register(path, sig, dtype, cg._GRAD_NAMES[grad], int(engine._is_2d(spec)) | (4 if o.get("compact") else 0))
        h0 = lib.ab_prog_hits()
        try:
            ms, out = timed(prog, spec, dtype, grad)
        except Exception as exc:
            print(json.dumps({"case": name, "opts": o, "error": str(exc)[:200]}), flush=True)
            continue
        assert lib.ab_prog_hits() > h0, "compiled kernel was not used"
        out = out if isinstance(out, tuple) else (out,)
        same = all(torch.equal(a.view(torch.int32 if dtype == "f32" else torch.int64),
                               b.view(torch.int32 if dtype == "f32" else torch.int64)) for a, b in zip(out, done_interp[name][1]))
        maxd = max(float((a - b).abs().max()) for a, b in zip(out, done_interp[name][1]))
        print(json.dumps({"case": name, "opts": o, "ms": round(ms, 4), "Gpts_s": round(n / ms / 1e6, 2),
                          "hbm_frac": round(n * bpp / ms / 1e6 / peak, 4), "vs_interp": round(done_interp[name][0] / ms, 3),
                          "bit_identical": bool(same), "max_abs_diff": maxd}), flush=True)
        del out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
