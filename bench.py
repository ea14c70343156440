#!/usr/bin/env python
"""bench.py — throughput of the composed-SDF hot path on B200 (BASELINE.json: "Gpts/s for composed 3D SDF at
512^3/1024^3 at 1/2/4/8 B200; % HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C5|C3|C1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one evaluation of the C5 workload: the C3 deep tree on the 1025^3 grid, scalar field plus
the analytic gradient field, fp32, x-slab sharded over the N ranks (no data-path collective). `value` is device
time with the result resident in HBM; `e2e` is the same job through the public API (aegolius_b200.create) into
pinned HOST buffers, i.e. including flattening, the program upload and the device->host copy of field + gradient.

--impl reference times the CPU restatement of the reference's NumPy path (oracle/interp_np.py, kind "port": the
reference itself is pure Python/NumPy and cannot travel to the GPU box) on a bounded sample of the same workload,
slab-parallel over all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid points/s, composed 3D SDF (field + analytic gradient)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu capture
    (profiles/r01_c5_dual_kernel_ncu_summary.json, taken on the whole 1025^3 grid), scaled to this rank's slab."""
    p = os.path.join(ROOT, "profiles", "r01_c5_dual_kernel_ncu_summary.json")
    try:
        with open(p) as fh:
            d = json.load(fh)

        def gb(key):
            v, unit = d[key].split()[:2]
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        total = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        return total * n_local / 1076890625.0
    except Exception:
        return None


def _workload(name):
    from aegolius_b200 import workloads, GridSpec, flatten
    cfg = workloads.CONFIGS[name]
    obj = cfg["build"]()
    spec = GridSpec(cfg["size"], cfg["res"])
    return obj, flatten(obj), spec


# ---- clocks ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.lines, self.gpu = None, [], gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = the upper half of the samples (the sampler also sees the idle edges)
        if sm:
            srt = sorted(sm)
            load = srt[len(srt) // 2:]
            med = float(np.median(load))
        else:
            med = None
        return {"sm_mhz": med, "sm_max_mhz": (max(mx) if mx else None), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- reference arm / cpu baseline -------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    (ops, a, slots, size, res, x0, x1) = args
    from aegolius_b200.program import Program
    from oracle import interp_np
    prog = Program(ops, a, [], slots[0], slots[1])
    # field on planes [x0, x1), gradient via the reference's from_sdf on the slab. No halo planes: the slab's two edge
    # planes get np.gradient's one-sided stencil, which costs the same per point as the interior one (timing sample)
    lo, hi = x0, x1
    t0 = time.perf_counter()
    f = interp_np.run_grid(prog, size, res, lo, hi)
    g = interp_np.from_sdf(f, (hi - lo, res[1], res[2]))
    dt = time.perf_counter() - t0
    return (x1 - x0) * res[1] * res[2], dt, float(f.sum()) + float(g[:, :8].sum())


def cpu_sample(prog, spec, planes_per_worker, workers):
    """One bounded CPU pass: `workers` processes, each evaluating `planes_per_worker` x-planes (field + from_sdf
    gradient), taken from the middle of the grid where the geometry is."""
    import multiprocessing as mp
    mid = spec.res[0] // 2
    start = max(1, mid - (workers * planes_per_worker) // 2)
    jobs = []
    for w in range(workers):
        x0 = start + w * planes_per_worker
        jobs.append((prog.ops, prog.args, (prog.n_pslots, prog.n_vslots), spec.size, spec.res, x0,
                     x0 + planes_per_worker))
    t0 = time.perf_counter()
    if workers == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    pts = sum(r[0] for r in res)
    return pts, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    obj, prog, spec = _workload(args.workload)
    cores = os.cpu_count() or 1
    planes = 2
    for _ in range(args.warmup):
        cpu_sample(prog, spec, planes, cores)
    pts_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        pts, wall = cpu_sample(prog, spec, planes, cores)
        pts_tot += pts
        t_tot += wall
    value = pts_tot / t_tot
    sample = (f"{cores} processes x {planes} x-plane(s) of {spec.res[1]}x{spec.res[2]} points per step from the middle "
              f"of the {spec.res[0]}^3 grid (field via the NumPy port of the reference path + from_sdf gradient)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_name(args.workload, spec), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _workload_name(name, spec):
    desc = {"C5": "C5: C3 deep tree (elongation, twist, bend, rot. symmetry, aligned curve instancing x21, smooth union, "
                  "mirror)", "C3": "C3 deep tree", "C1": "C1 sphere (+) box smooth union"}[name]
    return f"{desc} on the {spec.res[0]}x{spec.res[1]}x{spec.res[2]} grid, field + gradient, x-slab sharded"


# ---- GPU arm -----------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import aegolius_b200 as ab
    from aegolius_b200 import cabi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    if cabi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: aegolius_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    obj, prog, spec = _workload(args.workload)
    slabs = engine.slab_ranges(spec.res[0], world)
    x0, x1 = slabs[rank]
    n_local = (x1 - x0) * spec.res[1] * spec.res[2]
    n_total = spec.n_points
    stride = (n_local + 3) // 4 * 4
    field = torch.empty(n_local, dtype=torch.float32, device=dev)
    grad = torch.empty((3, stride), dtype=torch.float32, device=dev)

    def step():
        engine.create_torch(prog, spec, dtype="f32", grad="spatial", device=local, slab=(x0, x1), out=field,
                            out_grad=grad)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = cabi.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- e2e: public API, host buffers, D2H inside the timed region ----
    numa_cpus = None if os.environ.get("AB_NO_NUMA_BIND") else engine.bind_to_device_numa(local)
    pf = engine.PinnedArray((n_local,), np.float32)
    pg = engine.PinnedArray((3, n_local), np.float32)
    e2e_steps = max(1, min(args.steps, 3))
    ab.create(obj, spec, dtype="f32", grad="spatial", device=local, slab=(x0, x1), out=pf.array, out_grad=pg.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ab.create(obj, spec, dtype="f32", grad="spatial", device=local, slab=(x0, x1), out=pf.array,
                  out_grad=pg.array)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    checksum = float(pf.array[::4097].astype(np.float64).sum())
    pf.free()
    pg.free()

    # ---- secondary device-time probes (rank 0, N=1 only): the shallow-tree side of the roofline ----
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = {}
        sph = ab.Sphere(1.0)
        sph.move((0.3, 0.1, -0.2))
        probes = {"sphere_1025^3_f32": (ab.flatten(sph), ab.GridSpec((4, 4, 4), (1024,) * 3), None),
                  "C1_tree_1025^3_f32": (ab.flatten(ab.workloads.build_c1()), ab.GridSpec((4, 4, 4), (1024,) * 3), None),
                  "C5_tree_1025^3_f32_value_only": (prog, spec, None)}
        peak_hbm, _ = _peaks()
        for name, (pg, sp, gr) in probes.items():
            buf = torch.empty(sp.n_points, dtype=torch.float32, device=dev)
            for _ in range(3):
                engine.create_torch(pg, sp, dtype="f32", grad=gr, device=local, out=buf)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record()
            for _ in range(5):
                engine.create_torch(pg, sp, dtype="f32", grad=gr, device=local, out=buf)
            b.record()
            torch.cuda.synchronize(dev)
            t = a.elapsed_time(b) / 5
            secondary[name] = {"ms": round(t, 4), "Gpts_per_s": round(sp.n_points / t / 1e6, 1),
                               "output_GBps": round(4 * sp.n_points / t / 1e6, 1),
                               "hbm_frac": round(4 * sp.n_points / t / 1e6 / peak_hbm, 4), "ops": pg.n_ops}
            del buf
        # the same C5 step on a program-specialised build of the interpreter (engine.specialize: only this program's ops
        # are compiled in; identical results). Reported beside the headline, which stays on the ahead-of-time kernels.
        try:
            path = engine.specialize(prog, dtype="f32", grad="spatial")
            if path:
                fbuf = torch.empty(spec.n_points, dtype=torch.float32, device=dev)
                gbuf2 = torch.empty((3, (spec.n_points + 3) // 4 * 4), dtype=torch.float32, device=dev)
                for _ in range(3):
                    engine.create_torch(prog, spec, dtype="f32", grad="spatial", device=local, out=fbuf, out_grad=gbuf2)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                a.record()
                for _ in range(5):
                    engine.create_torch(prog, spec, dtype="f32", grad="spatial", device=local, out=fbuf, out_grad=gbuf2)
                b.record()
                torch.cuda.synchronize(dev)
                t = a.elapsed_time(b) / 5
                secondary["C5_field+gradient_specialised_kernel"] = {
                    "ms": round(t, 4), "Gpts_per_s": round(spec.n_points / t / 1e6, 1),
                    "hbm_frac": round(16 * spec.n_points / t / 1e6 / peak_hbm, 4),
                    "note": "opt-in aegolius_b200.specialize(obj); compiled (about 10 s, cached) outside any timed region"}
                del fbuf, gbuf2
            cabi_mod = __import__("aegolius_b200.cabi", fromlist=["lib"])
            cabi_mod.lib().ab_spec_clear()
        except Exception as exc:  # no nvcc on this host: the probe is skipped, nothing else depends on it
            secondary["C5_field+gradient_specialised_kernel"] = {"skipped": str(exc)[:200]}
        # C4: point cloud -> unsigned distance (exact octree nearest neighbour), device time of ab_nn_grid incl. tree build
        import ctypes as C
        from aegolius_b200 import cabi, workloads
        cloud = workloads.c4_cloud()
        c4 = ab.GridSpec(workloads.CONFIGS["C4"]["size"], workloads.CONFIGS["C4"]["res"])
        lib, d_cloud = cabi.lib(), C.c_void_p()
        cabi.check(lib.ab_cloud_upload(cloud.ctypes.data, cloud.shape[1], 3, cloud.shape[1], cabi.AB_F32, local, C.byref(d_cloud)))
        buf = torch.empty(c4.n_points, dtype=torch.float32, device=dev)
        g = cabi.make_grid(c4.size, c4.res)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        l0 = cabi.launch_count()
        for _ in range(2):
            cabi.check(lib.ab_nn_grid(d_cloud, cloud.shape[1], 3, C.byref(g), cabi.AB_F32, buf.data_ptr(), local, st))
        per_call = (cabi.launch_count() - l0) // 2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(5):
            cabi.check(lib.ab_nn_grid(d_cloud, cloud.shape[1], 3, C.byref(g), cabi.AB_F32, buf.data_ptr(), local, st))
        b.record()
        torch.cuda.synchronize(dev)
        t = a.elapsed_time(b) / 5
        secondary["C4_cloud_1M_points_257^3_f32"] = {"ms": round(t, 4), "Mqueries_per_s": round(c4.n_points / t / 1e3, 1),
                                                    "launches_per_call": per_call,
                                                    "equivalent_Tpairs_per_s": round(c4.n_points * cloud.shape[1] / t / 1e9, 1)}
        lib.ab_device_free(d_cloud, local)
        del buf
        torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = 16.0 * n_local  # 4 B field + 12 B gradient written per point, 0 B read (grid mode)
        kern_ms = ms_per_step  # one interpreter launch per step per rank
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": _workload_name(args.workload, spec), "points": n_total,
                       "slab_planes_rank0": x1 - x0, "program_ops": prog.n_ops, "program_bytes": prog.nbytes(),
                       "l2": "outputs (16 B/point, >= 2 GB per rank) far exceed the 126 MB L2; no inputs are read"},
            "clocks": clocks,
            "e2e": {"value": n_total / e2e_s, "unit": "points/s", "h2d_bytes_per_step": prog.nbytes(),
                    "d2h_bytes_per_step": 16 * n_local, "ms_per_step": e2e_s * 1e3,
                    "note": "aegolius_b200.create(obj, grid, grad='spatial') into pinned host arrays (per rank)",
                    "numa_bound_cpus": len(numa_cpus) if numa_cpus else 0},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(n_local), "peak_source": peak_src,
                         "kernel": "ab_interp_kernel<Dual<Pack<float,2>,3>,float>", "algorithmic_bytes_per_point": 16,
                         "note": "deep tree: FP32-issue-bound, not HBM-bound (SURVEY §8d); 'secondary' holds the shallow-tree "
                                 "probes that are write-bound; traffic: ncu capture in profiles/ scaled to this slab"},
            "checksum": checksum,
        }
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu and world == 1:
            cores = 1
            pts, wall = cpu_sample(prog, spec, 2, 1)
            line["cpu_baseline"] = {"value": pts / wall, "unit": "points/s", "cores": cores, "kind": "port",
                                    "sample": f"2 x-planes of {spec.res[1]}x{spec.res[2]} points from the middle of "
                                              f"the grid, NumPy port of the reference path (field + from_sdf gradient)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=["C5", "C3", "C1"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the shallow-tree roofline probes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
