#!/usr/bin/env python
"""The device -> host ceiling the end-to-end number runs against (VERDICT r01 "Next" #4): every rank copies a device
buffer into page-locked host memory with plain cudaMemcpyAsync calls, one per chunk, all ranks at the same time, and the
rate is reported per rank and in aggregate. Variants: NUMA binding before the allocation, portable / write-combined pages,
chunk size, one or two copy streams.

    python tools/d2h_ceiling.py                                           # 1 GPU
    python -m torch.distributed.run --nproc-per-node N tools/d2h_ceiling.py   # N GPUs copying concurrently

One JSON line per variant on rank 0 (aggregate = sum over ranks of bytes / max over ranks of the wall time)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0, help="bytes copied per rank and variant")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from aegolius_b200 import cabi, engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.lib()
    nbytes = int(args.gb * (1 << 30))
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    src.fill_(7)
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def run(host_ptr, chunk, n_streams):
        best = 1e30
        for _ in range(args.reps + 1):
            barrier()
            t0 = time.perf_counter()
            off, i = 0, 0
            while off < nbytes:
                n = min(chunk, nbytes - off)
                st = streams[i % n_streams]
                cabi.check(lib.ab_memcpy_d2h(C.c_void_p(host_ptr + off), C.c_void_p(src.data_ptr() + off), n, local,
                                             C.c_void_p(st.cuda_stream)))
                off += n
                i += 1
            for st in streams[:n_streams]:
                st.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        return best

    variants = []
    for bind in (False, True):
        for flags, fname in ((0, "default"), (1, "portable"), (2, "write-combined"), (3, "portable+write-combined")):
            variants.append((bind, flags, fname, 256 << 20, 1))
    variants += [(True, 0, "default", 64 << 20, 1), (True, 0, "default", 1 << 30, 1), (True, 0, "default", 256 << 20, 2),
                 (True, 2, "write-combined", 256 << 20, 2)]
    bound = False
    cpus0 = os.sched_getaffinity(0)
    for bind, flags, fname, chunk, ns in variants:
        if bind and not bound:
            engine.bind_to_device_numa(local)
            bound = True
        if not bind and bound:
            os.sched_setaffinity(0, cpus0)
            bound = False
        p = C.c_void_p()
        cabi.check(lib.ab_host_alloc_pinned_flags(nbytes, flags, C.byref(p)))
        dt = run(p.value, chunk, ns)
        lib.ab_host_free_pinned(p)
        if rank == 0:
            print(json.dumps({"world": world, "numa_bind_before_alloc": bind, "pages": fname, "chunk_MB": chunk >> 20,
                              "copy_streams": ns, "GB_per_rank": args.gb, "seconds": round(dt, 4),
                              "per_rank_GBps": round(nbytes / dt / 1e9, 2),
                              "aggregate_GBps": round(world * nbytes / dt / 1e9, 2)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
