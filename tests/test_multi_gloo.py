"""World-size-2 CPU test (gloo) of the N>1 host logic: contiguous point shards from
ordm_shard_range, each rank regenerating its slice with the counter-based generator, partial
scalars combined by one all-reduce(SUM) -- the same structure bench.py runs over NCCL, with the
CPU oracle standing in for the kernels (tests may use the oracle; the product never does)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["n_tot", "n_a", "n_b", "trn_tot", "trn_a", "trn_b", "ex", "ec"]
CASE = dict(seed=3, npts=3000 + 17, ncore=2, nact=4, nel=2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import openrdm_b200 as ordm
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    c = CASE
    n = c["ncore"] + c["nact"]
    A1a, A1b, D2 = ordm.synth_rdms(c["seed"], c["nact"], c["nel"], c["nel"], 4)
    p0, cnt = ordm.shard_range(c["npts"], world, rank)
    w, phi, g = ordm.synth_fill_host(c["seed"], c["npts"], p0, cnt, n, True)
    r = O.Port().energy_coreactive(w, phi, c["ncore"], A1a, A1b, D2, "PBE", g, eclass=0.0, want_vecs=False)
    t = torch.tensor([r.scalars[k] for k in KEYS], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    cover = torch.tensor([cnt], dtype=torch.int64)
    dist.all_reduce(cover)
    if rank == 0:
        out.put((t.tolist(), int(cover.item())))
    dist.destroy_process_group()


def test_two_rank_sharded_energy_equals_single_rank():
    sys.path.insert(0, ROOT)
    import openrdm_b200 as ordm
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    vals, cover = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = CASE
    n = c["ncore"] + c["nact"]
    assert cover == c["npts"]
    A1a, A1b, D2 = ordm.synth_rdms(c["seed"], c["nact"], c["nel"], c["nel"], 4)
    w, phi, g = ordm.synth_fill_host(c["seed"], c["npts"], 0, c["npts"], n, True)
    whole = O.Port().energy_coreactive(w, phi, c["ncore"], A1a, A1b, D2, "PBE", g, eclass=0.0, want_vecs=False)
    for k, v in zip(KEYS, vals):
        assert abs(v - whole.scalars[k]) <= 1e-12 * max(1.0, abs(whole.scalars[k])), k
