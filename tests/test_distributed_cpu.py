"""The N>1 path on CPU: world_size-2 gloo processes shard the grid into x-slabs (no data-path collective), evaluate
their slab (here with the oracle, since there is no GPU) and assemble the field with the optional all_gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aegolius_b200 as ab
    from aegolius_b200 import distributed as abd
    from oracle import interp_np
    spec = ab.GridSpec((4, 4, 4), (10, 6, 8))  # 11 planes over 2 ranks: 6 + 5 (unequal slabs -> padded gather)
    prog = ab.flatten(ab.workloads.build_c1())
    x0, x1 = abd.rank_slab(spec.res[0], rank, world)
    local = torch.from_numpy(interp_np.run_grid(prog, spec.size, spec.res, x0, x1))
    full = abd.gather_field(local, spec)
    whole = interp_np.run_grid(prog, spec.size, spec.res)
    ok = bool(np.array_equal(full.numpy(), whole))
    q.put((rank, (x0, x1), ok, int(abd.padded_slab_points(spec, world))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_sharding_and_gather_on_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 6), (6, 11)]
    assert all(r[2] for r in res)
    assert res[0][3] == 6 * 7 * 9


def test_slab_ranges_match_array_split():
    from aegolius_b200.engine import slab_ranges
    for n, parts in ((1025, 8), (513, 4), (11, 2), (7, 7), (129, 3)):
        ref = [(int(a[0]), int(a[-1]) + 1) for a in np.array_split(np.arange(n), parts)]
        assert slab_ranges(n, parts) == ref
