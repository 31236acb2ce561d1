"""world_size-2 gloo test of the N > 1 host logic (runs on CPU): chains shard across ranks by
chain_offset with no data-path collective; the only exchange is the end-of-run reduction of the
ESS sufficient statistics, which must equal the single-process result."""
import os
import socket
import sys

import numpy as np
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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from walnuts_b200 import diagnostics as dg
    rng = np.random.default_rng(42)
    draws = rng.standard_normal((64, 50)).cumsum(axis=1) * 0.1 + rng.standard_normal((64, 50))
    per = 64 // world
    mine = draws[rank * per:(rank + 1) * per]            # shard = contiguous chain ids (chain_offset = rank*per)
    st = dg.chain_stats(mine, 20)
    vec = torch.tensor([st["m"], st["sum_mean"], st["sum_mean2"], st["sum_var"], *st["acov_sum"]], dtype=torch.float64)
    dist.all_reduce(vec)
    v = vec.numpy()
    ess = dg.ess_from_stats(dict(m=v[0], n=50, sum_mean=v[1], sum_mean2=v[2], sum_var=v[3], acov_sum=v[4:]))[0]
    if rank == 0:
        whole = dg.ess_from_stats(dg.chain_stats(draws, 20))[0]
        q.put((ess, whole))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ess_reduction_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ess, whole = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert np.isclose(ess, whole, rtol=1e-10)
