"""world_size-2 gloo test of the multi-GPU host logic (partition + gather) on CPU.  The aligner injected into
run_sharded is the oracle (allowed: this is a test); on a GPU box bench.py injects the CUDA path instead."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import checkers as ck
from bsalign_b200 import shard, synth


def test_balanced_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    work = rng.integers(1000, 2000000, size=5000)
    for g in (2, 4, 8):
        parts = shard.balanced_partition(work, g)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(work)))
        tot = np.array([work[p].sum() for p in parts], dtype=np.float64)
        assert tot.max() / tot.mean() < 1.01


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = synth.make_pairs(60, 150, seed=11)
    mtx = synth.score_matrix(2, -6)

    def align(sub):
        r, c, _ = ck.oracle_batch("epi8", sub, 1, 64, mtx, (-3, -2, 0, 0))
        return r, np.zeros(sub.n, np.int32), c
    out = shard.run_sharded(batch, "epi8", 64, align, dist)
    if rank == 0:
        exp_r, exp_c, _ = ck.oracle_batch("epi8", batch, 1, 64, mtx, (-3, -2, 0, 0))
        ok = np.array_equal(out[0], exp_r) and all(np.array_equal(a, b) for a, b in zip(out[2], exp_c))
        q.put(bool(ok))
    dist.destroy_process_group()


def test_sharded_run_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _poa_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import poa_jobs as pj
    from bsalign_b200 import poa, synth_poa
    jobs = [synth_poa.make_sweep_job(40 + i, tlen=300 + 40 * (i % 5)) for i in range(11)]
    mine, idx = shard.shard_jobs(jobs, rank, world)
    r = pj.oracle_sweep_batch(poa.SweepBatch(mine), nthreads=1, want_rows=False)   # the oracle stands in for the GPU sweep (a test)
    out = shard.gather_jobs_to_rank0(r["best"], None, idx, len(jobs), dist)
    if rank == 0:
        exp = pj.oracle_sweep_batch(poa.SweepBatch(jobs), nthreads=1, want_rows=False)
        q.put(bool(np.array_equal(out[0], exp["best"])))
    dist.destroy_process_group()


def test_sharded_poa_jobs_equal_single_process():
    work = shard.job_work
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_poa_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
