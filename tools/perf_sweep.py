#!/usr/bin/env python
"""Device-time sweep over the BASELINE configs (and the shallow-tree roofline probes). Prints one JSON line per case:
Gpts/s, achieved GB/s of algorithmic output bytes and the fraction of the measured HBM peak. Used to steer kernel work;
bench.py stays the contract benchmark."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import aegolius_b200 as ab
    from aegolius_b200 import engine, workloads, cabi
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="sphere,c1,c3,c3g,c2,c1_64,c3_64,nn")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nn-res", type=int, default=128)
    ap.add_argument("--spec", action="store_true", help="register a program-specialised kernel (engine.specialize) first")
    args = ap.parse_args()
    peak = 6452.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    dev = torch.device("cuda", 0)

    sph = ab.Sphere(1.0)
    sph.move((0.3, 0.1, -0.2))
    cases = {
        "sphere": (sph, ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "c1": (workloads.build_c1(), ab.GridSpec((4, 4, 4), (1024,) * 3), "f32", None),
        "c1g": (workloads.build_c1(), ab.GridSpec((4, 4, 4), (768,) * 3), "f32", "spatial"),
        "c3": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (512,) * 3), "f32", None),
        "c3g": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (512,) * 3), "f32", "spatial"),
        "c2": (workloads.build_c2(), ab.GridSpec((8, 8), (4096, 4096)), "f32", None),
        "c1_64": (workloads.build_c1(), ab.GridSpec((4, 4, 4), (768,) * 3), "f64", None),
        "c3_64": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (384,) * 3), "f64", None),
        "c3g_64": (workloads.build_c3(), ab.GridSpec((6, 6, 6), (256,) * 3), "f64", "spatial"),
    }
    for name in args.cases.split(","):
        if name == "nn":
            pts = workloads.c4_cloud(1_000_000)
            spec = ab.GridSpec((2.5, 2.5, 1.5), (args.nn_res,) * 3)
            import ctypes as C
            lib = cabi.lib()
            d_cloud = C.c_void_p()
            cabi.check(lib.ab_cloud_upload(pts.ctypes.data, pts.shape[1], 3, pts.shape[1], cabi.AB_F32, 0, C.byref(d_cloud)))
            out = torch.empty(spec.n_points, dtype=torch.float32, device=dev)
            g = cabi.make_grid(spec.size, spec.res)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

            def run():
                cabi.check(lib.ab_nn_grid(d_cloud, pts.shape[1], 3, C.byref(g), cabi.AB_F32, out.data_ptr(), 0, st))
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            pairs = spec.n_points * pts.shape[1]
            print(json.dumps({"case": f"nn {spec.res} x 1M", "ms": ms, "Tpairs_per_s": pairs / ms / 1e9,
                              "Mqueries_per_s": spec.n_points / ms / 1e3}), flush=True)
            continue
        if name == "fields":  # whole-field streaming kernels on a 513^3 fp32 field (device time, algorithmic bytes)
            import ctypes as C
            lib = cabi.lib()
            res = (513, 513, 513)
            n = res[0] * res[1] * res[2]
            stride = (n + 7) // 8 * 8
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            f = torch.randn(n, dtype=torch.float32, device=dev)
            o = torch.empty(n, dtype=torch.float32, device=dev)
            v = torch.randn(3, stride, dtype=torch.float32, device=dev)
            ang = torch.randn(n, dtype=torch.float32, device=dev)
            r3 = (C.c_uint32 * 3)(*res)
            k3 = (C.c_uint32 * 3)(5, 5, 1)
            g = cabi.make_grid((0.0, 0.0, 0.0), res)
            ops = (cabi.ab_vec_op * 2)()
            ops[0].opcode, ops[0].kind0, ops[0].a0 = cabi.AB_VOP_ROT_Z, cabi.AB_VK_ARRAY, ang.data_ptr()
            ops[1].opcode, ops[1].kind0, ops[1].kind1, ops[1].a1 = cabi.AB_VOP_ROT_AXIS, cabi.AB_VK_VEC3, cabi.AB_VK_ARRAY, ang.data_ptr()
            ops[1].c = (C.c_double * 3)(1.0, 0.0, 0.0)
            jobs = {
                "box_filter_5x5x1 (2 passes: 16 B/pt)": (lambda: cabi.check(lib.ab_box_filter(f.data_ptr(), r3, k3, 1, cabi.AB_F32, o.data_ptr(), 0, st)), 16),
                "edge_filter (8 B/pt)": (lambda: cabi.check(lib.ab_edge_filter(f.data_ptr(), r3, cabi.AB_F32, o.data_ptr(), 0, st)), 8),
                "from_sdf (16 B/pt)": (lambda: cabi.check(lib.ab_fd_gradient(f.data_ptr(), 0, C.byref(g), 3, cabi.AB_F32, 1, v.data_ptr(), stride, 0, st)), 16),
                "vec rotate_z+rotate_axis (32 B/pt)": (lambda: cabi.check(lib.ab_vec_apply(v.data_ptr(), stride, n, ops, 2, cabi.AB_F32, 0, st)), 32),
                "vec component phi (12 B/pt)": (lambda: cabi.check(lib.ab_vec_component(v.data_ptr(), stride, n, cabi.AB_VC_PHI, cabi.AB_F32, o.data_ptr(), 0, st)), 12),
            }
            for jn, (fn, bpp) in jobs.items():
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                print(json.dumps({"case": jn, "res": res, "ms": round(ms, 4), "GBps": round(bpp * n / ms / 1e6, 1),
                                  "hbm_frac": round(bpp * n / ms / 1e6 / peak, 4)}), flush=True)
            continue
        obj, spec, dt, grad = cases[name]
        prog = ab.flatten(obj)
        if args.spec:
            cabi.lib().ab_spec_clear()
            engine.specialize(prog, dtype=dt, grad=grad)
        tdt = torch.float32 if dt == "f32" else torch.float64
        n = spec.n_points
        field = torch.empty(n, dtype=tdt, device=dev)
        gbuf = torch.empty((3, (n + 7) // 8 * 8), dtype=tdt, device=dev) if grad else None

        def run():
            engine.create_torch(prog, spec, dtype=dt, grad=grad, out=field, out_grad=gbuf)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        bpp = (4 if dt == "f32" else 8) * (4 if grad else 1)
        gbs = bpp * n / ms / 1e6
        print(json.dumps({"case": name, "res": spec.res, "dtype": dt, "grad": bool(grad), "ops": prog.n_ops,
                          "ms": round(ms, 4), "best_ms": round(min(times), 4), "Gpts_per_s": round(n / ms / 1e6, 2),
                          "GBps": round(gbs, 1), "hbm_frac": round(gbs / peak, 4)}), flush=True)


if __name__ == "__main__":
    main()
