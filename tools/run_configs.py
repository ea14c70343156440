#!/usr/bin/env python
"""Measures all five BASELINE.json configurations at full size on one B200 (device time, CUDA events) next to the NumPy
port of the reference path on one host core (bounded sample), and writes profiles/<tag>_configs.json."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_time(fn, reps=5):
    import torch
    fn()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    import torch
    import aegolius_b200 as ab
    from aegolius_b200 import engine, workloads, cabi
    from oracle import interp_np
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    args = ap.parse_args()
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda", 0)
    out = {"hbm_peak_gbs": peak, "host_cores": os.cpu_count(), "configs": {}}
    for name in ("C1", "C2", "C3", "C5"):
        cfg = workloads.CONFIGS[name]
        spec = ab.GridSpec(cfg["size"], cfg["res"])
        prog = ab.flatten(cfg["build"]())
        n = spec.n_points
        modes = [("f32", None)] + ([("f32", "spatial")] if name in ("C3", "C5") else []) + [("f64", None)]
        entry = {"grid": spec.res, "points": n, "ops": prog.n_ops}
        for dt, grad in modes:
            if name == "C5" and dt == "f64":
                continue
            tdt = torch.float32 if dt == "f32" else torch.float64
            field = torch.empty(n, dtype=tdt, device=dev)
            gbuf = torch.empty((3, (n + 7) // 8 * 8), dtype=tdt, device=dev) if grad else None
            ms = gpu_time(lambda: engine.create_torch(prog, spec, dtype=dt, grad=grad, out=field, out_grad=gbuf))
            bpp = (4 if dt == "f32" else 8) * (4 if grad else 1)
            entry[f"gpu_{dt}{'_grad' if grad else ''}"] = {
                "ms": round(ms, 4), "Gpts_per_s": round(n / ms / 1e6, 2), "output_GBps": round(bpp * n / ms / 1e6, 1),
                "hbm_frac": round(bpp * n / ms / 1e6 / peak, 4)}
            del field, gbuf
            torch.cuda.empty_cache()
        # CPU: NumPy port on a bounded sample of x planes (one core)
        per_plane = spec.res[1] * spec.res[2]
        planes = max(1, min(spec.res[0], int(2.0e6 // per_plane)))
        x0 = spec.res[0] // 2 - planes // 2
        t0 = time.perf_counter()
        interp_np.run_grid(prog, spec.size, spec.res, x0, x0 + planes)
        dt_cpu = time.perf_counter() - t0
        entry["cpu_port_1core"] = {"Mpts_per_s": round(planes * per_plane / dt_cpu / 1e6, 3),
                                   "sample": f"{planes} x-plane(s) = {planes * per_plane} points, field only"}
        out["configs"][name] = entry
        print(name, json.dumps(entry), flush=True)
    # C4: 1M-point cloud on 257^3
    cfg = workloads.CONFIGS["C4"]
    spec = ab.GridSpec(cfg["size"], cfg["res"])
    pts = workloads.c4_cloud(1_000_000)
    lib = cabi.lib()
    d_cloud = C.c_void_p()
    cabi.check(lib.ab_cloud_upload(pts.ctypes.data, pts.shape[1], 3, pts.shape[1], cabi.AB_F32, 0, C.byref(d_cloud)))
    o = torch.empty(spec.n_points, dtype=torch.float32, device=dev)
    g = cabi.make_grid(spec.size, spec.res)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms = gpu_time(lambda: cabi.check(lib.ab_nn_grid(d_cloud, pts.shape[1], 3, C.byref(g), cabi.AB_F32, o.data_ptr(), 0, st)),
                  reps=2)
    pairs = spec.n_points * pts.shape[1]
    from scipy.spatial import cKDTree
    q = spec.slab_coords(spec.res[0] // 2, spec.res[0] // 2 + 2)
    t0 = time.perf_counter()
    cKDTree(pts.T).query(q.T)
    dt_cpu = time.perf_counter() - t0
    out["configs"]["C4"] = {"grid": spec.res, "points": spec.n_points, "cloud": pts.shape[1],
                            "gpu_f32": {"ms": round(ms, 2), "Mqueries_per_s": round(spec.n_points / ms / 1e3, 2),
                                        "Tpairs_per_s": round(pairs / ms / 1e9, 3)},
                            "cpu_ckdtree_1core": {"Mqueries_per_s": round(q.shape[1] / dt_cpu / 1e6, 4),
                                                  "sample": f"2 x-planes = {q.shape[1]} queries, tree build included"}}
    print("C4", json.dumps(out["configs"]["C4"]), flush=True)
    path = os.path.join(ROOT, "gpurun_out", f"{args.tag}_configs.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
