#!/usr/bin/env python
"""python -m torch.distributed.run --nproc-per-node N tools/check_multi_gpu.py [--big] : the multi-GPU paths on real
hardware, each checked bit-for-bit against a single-GPU evaluation of the whole grid on rank 0:

  * x-slab sharding of a deep tree, field + analytic gradient (aegolius_b200.distributed.evaluate_sharded),
  * in-place assembly of the field and of the gradient rows (gather_field / gather_rows: no staging buffers),
  * sharded point cloud -> distance (cloud replicated by one broadcast) and sharded from_sdf (recomputed halo planes),
  * assembly by the evaluation kernel itself through the NVLS multicast address (evaluate_multicast), when the node has it.

--big adds the timings on the BASELINE grid (1025^3 fp32 field): compute per rank, in-place all-gather, multicast
evaluation, as device times (max over ranks). One JSON line per check on rank 0; exit code 1 on any mismatch."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--res", type=int, default=192)
    args = ap.parse_args()
    import numpy as np
    import aegolius_b200 as ab
    from aegolius_b200 import distributed as abd, engine
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    ok_all = True

    def report(name, ok, **kw):
        nonlocal ok_all
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag[0])
        ok_all = ok_all and ok
        if rank == 0:
            print(json.dumps(dict(check=name, world=world, ok=ok, **kw)), flush=True)

    def timed(fn, reps=3):
        """device time of fn, max over ranks (ms), after one warm-up"""
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t = torch.tensor([best], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- 1. sharded field + gradient, assembled in place, against the whole grid on one GPU --------------------------------
    spec = ab.GridSpec((6, 6, 6), (args.res,) * 3)
    prog = ab.flatten(ab.workloads.build_c3())
    engine.compile_program(prog, dtype="f32", grad="spatial")
    engine.compile_program(prog, dtype="f32")
    full_f, full_g = abd.evaluate_sharded(prog, spec, dtype="f32", grad="spatial", gather=True)
    whole_f, whole_g = engine.create_torch(prog, spec, dtype="f32", grad="spatial", device=local)
    report("sharded field + gradient, gathered in place == single-GPU whole grid",
           torch.equal(full_f, whole_f) and torch.equal(full_g, whole_g.contiguous()), grid=spec.res)

    # ---- 2. sharded point cloud ------------------------------------------------------------------------------------------------
    pts = ab.workloads.c4_cloud(200_000, seed=3) if rank == 0 else None
    cspec = ab.GridSpec((2.5, 2.5, 1.5), (96, 96, 96))
    d_full = abd.point_cloud_sharded(cspec, pts, dtype="f32", gather=True)
    ok = True
    if rank == 0:
        d_whole = engine.point_cloud_sdf_torch(cspec, engine.cloud_records(pts, 3, "f32", local), device=local)
        ok = torch.equal(d_full, d_whole)
    report("sharded point cloud (cloud broadcast, query slabs) == single-GPU", ok, grid=cspec.res, cloud=200_000)

    # ---- 3. sharded from_sdf with recomputed halo planes ---------------------------------------------------------------------
    f_slab, v_slab = abd.from_sdf_sharded(prog, spec, dtype="f32")
    v_full = abd.gather_rows(v_slab.contiguous(), spec)
    whole_v = engine.from_sdf_torch(engine.create_torch(prog, spec, dtype="f32", device=local), spec.res, device=local)
    report("sharded from_sdf (halo planes recomputed) == whole-grid from_sdf", torch.equal(v_full, whole_v.contiguous()))

    # ---- 4. multicast assembly ---------------------------------------------------------------------------------------------------
    mc = None
    try:
        mc = abd.MulticastField(spec, "f32", grad=True)
    except Exception as exc:  # no NVLS on this node
        if rank == 0:
            print(json.dumps({"check": "multicast assembly", "skipped": str(exc)[:300]}), flush=True)
    if mc is not None:
        mc.buf.zero_()
        mc.barrier()
        abd.evaluate_multicast(prog, mc, grad="spatial")
        mc.barrier()
        report("multimem.st assembly: every rank holds the whole field + gradient, == single-GPU",
               torch.equal(mc.field, whole_f) and torch.equal(mc.grad, whole_g))
        del mc

    # ---- 5. timings on the BASELINE grid --------------------------------------------------------------------------------------
    if args.big:
        del full_f, full_g, whole_f, whole_g, v_full, whole_v
        torch.cuda.empty_cache()
        big = ab.GridSpec((4, 4, 4), (1024,) * 3)
        sph = ab.Sphere(1.0)
        sph.move((0.3, 0.1, -0.2))
        for name, obj in (("sphere (write-bound)", sph), ("C5 deep tree", ab.workloads.build_c3())):
            p = ab.flatten(obj)
            engine.compile_program(p, dtype="f32")
            x0, x1 = abd.rank_slab(big.res[0], rank, world)
            n_local = (x1 - x0) * big.res[1] * big.res[2]
            loc = torch.empty(n_local, dtype=torch.float32, device=dev)
            out = torch.empty(big.n_points, dtype=torch.float32, device=dev)
            t_comp = timed(lambda: engine.create_torch(p, big, dtype="f32", device=local, slab=(x0, x1), out=loc))
            t_gather = timed(lambda: abd.gather_field(loc, big, out=out))
            res = {"case": name, "grid": big.res, "world": world, "field_GB": big.n_points * 4 / 1e9,
                   "compute_ms": round(t_comp, 3), "gather_in_place_ms": round(t_gather, 3),
                   "gather_recv_GBps_per_gpu": round((big.n_points - n_local) * 4 / t_gather / 1e6, 1)}
            try:
                mcb = abd.MulticastField(big, "f32")
                t_mc = timed(lambda: abd.evaluate_multicast(p, mcb))
                mcb.barrier()
                res["multicast_eval_ms"] = round(t_mc, 3)
                res["multicast_recv_GBps_per_gpu"] = round(big.n_points * 4 / t_mc / 1e6, 1)
                res["multicast_bit_identical_to_gather"] = bool(torch.equal(mcb.field, out))
                del mcb
            except Exception as exc:
                res["multicast"] = "unavailable: " + str(exc)[:200]
            if rank == 0:
                print(json.dumps(res), flush=True)
            del loc, out
            torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()
    if not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
