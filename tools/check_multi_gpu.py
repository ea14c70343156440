#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multi_gpu.py : every rank evaluates its x-slab on its own GPU, the field is
assembled with the optional NCCL all_gather, and rank 0 checks it bit-for-bit against a single-GPU evaluation of the
whole grid. Also times the gather separately from the compute (gather >> compute)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import aegolius_b200 as ab
    from aegolius_b200 import distributed as abd, engine
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    spec = ab.GridSpec((6, 6, 6), (512, 512, 512))
    prog = ab.flatten(ab.workloads.build_c3())
    x0, x1 = abd.rank_slab(spec.res[0], rank, world)
    local_field = engine.create_torch(prog, spec, dtype="f32", device=local, slab=(x0, x1))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    local_field = engine.create_torch(prog, spec, dtype="f32", device=local, slab=(x0, x1))
    e1.record()
    full = abd.gather_field(local_field, spec)
    e2.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = None
    if rank == 0:
        whole = engine.create_torch(prog, spec, dtype="f32", device=local)
        ok = bool(torch.equal(whole, full))
        print(json.dumps({"world": world, "grid": spec.res, "bit_identical_to_single_gpu": ok,
                          "compute_ms_max": float(t[0]), "gather_ms_max": float(t[1]),
                          "field_GB": spec.n_points * 4 / 1e9}))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
