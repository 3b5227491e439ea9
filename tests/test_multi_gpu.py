"""Two-GPU test (skipped with fewer devices) of the real N>1 path: one process per GPU, grid sharded
by ordm_shard_range, NCCL communicator attached, and the eight result scalars combined (i) by this
library's NVLink peer-store kernel and (ii) by ncclAllReduce -- both against the single-GPU result
and the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["e_tot", "ex", "ec", "n_tot", "n_a", "n_b", "trn_tot", "trn_a", "trn_b"]
CASE = dict(seed=3, npts=50_000 + 17, ncore=2, nact=6, nel=3)


def _worker(rank, world, port, no_peer, out):
    sys.path.insert(0, ROOT)
    if no_peer:
        os.environ["ORDM_NO_PEER_ALLREDUCE"] = "1"
    import torch.distributed as dist
    import openrdm_b200 as ordm
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    c = CASE
    n = c["ncore"] + c["nact"]
    A1a, A1b, D2 = ordm.synth_rdms(c["seed"], c["nact"], c["nel"], c["nel"], 4)
    eng = ordm.Engine(rank)
    eng.set_active_space(c["ncore"], A1a, A1b, D2)
    p0, cnt = ordm.shard_range(c["npts"], world, rank)
    eng.synth_fill(c["seed"], c["npts"], p0, cnt, n, True)
    uid = [ordm.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_attach(world, rank, uid[0])
    info = eng.comm_info()
    res = [eng.energy(ordm.FUNC_PBE, -1.0) for _ in range(3)][-1]  # repeated calls exercise the epoch / double buffer
    out.put((rank, info["allreduce"], {k: res[k] for k in KEYS}))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("no_peer", [False, True])
def test_two_gpu_sharded_energy(ordm, port, no_peer):
    if ordm.load().ordm_device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    tcp = 29700 + os.getpid() % 200 + (50 if no_peer else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, tcp, no_peer, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = CASE
    n = c["ncore"] + c["nact"]
    A1a, A1b, D2 = ordm.synth_rdms(c["seed"], c["nact"], c["nel"], c["nel"], 4)
    w, phi, g = ordm.synth_fill_host(c["seed"], c["npts"], 0, c["npts"], n, True)
    want = port.energy_coreactive(w, phi, c["ncore"], A1a, A1b, D2, "PBE", g, eclass=-1.0, want_vecs=False).scalars
    for rank, how, vals in got:
        assert ("ncclAllReduce" in how) == no_peer, how
        for k in KEYS:
            t = 1e-7 if k in ("trn_a", "trn_b") else 1e-10
            assert abs(vals[k] - want[k]) <= t * max(1.0, abs(want[k])), (rank, k, vals[k], want[k])
    # every rank holds the same sums (bitwise with the rank-ordered peer kernel)
    if not no_peer:
        assert got[0][2] == got[1][2]
