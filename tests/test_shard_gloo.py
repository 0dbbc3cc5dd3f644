"""world_size-2 gloo test of the multi-GPU path (plan, compact arenas, scatter, gather, pair-ordered merge) on CPU.  The aligner
injected into run_sharded_device is the oracle (allowed: this is a test); on GPUs bench.py and tests/test_gpu_parity.py inject
shard.cuda_aligner, the CUDA path."""
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


def _oracle_aligner(kind, mode, bw, mtx, gaps):
    """Stands in for shard.cuda_aligner on a box without GPUs: same contract (arena tensor + ShardView in, records + dense cigars out)."""
    import torch

    def align(arena, view):
        sub = synth.PairBatch(arena.numpy(), view.qoff, view.qlen, view.toff, view.tlen)
        r, c, _ = ck.oracle_batch(kind, sub, mode, bw, mtx, gaps)
        rec = np.zeros((sub.n, 12), dtype=np.int32)
        rec[:, :10] = r
        rec[:, 11] = [len(x) for x in c]
        dense = np.concatenate(c).astype(np.uint32) if sub.n and rec[:, 11].sum() else np.zeros(0, np.uint32)
        return torch.from_numpy(rec), torch.from_numpy(dense.view(np.int32).copy()), {}
    return align


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mtx = synth.score_matrix(2, -6)
    # ragged lengths, one empty pair: rank 0 owns the batch, the others only learn the lengths
    rng = np.random.default_rng(3)
    pairs = [(rng.integers(0, 4, int(rng.integers(1, 300))).astype(np.uint8), rng.integers(0, 4, int(rng.integers(1, 300))).astype(np.uint8)) for _ in range(59)]
    pairs.insert(7, (np.zeros(0, np.uint8), np.array([1, 2], np.uint8)))
    batch = synth.PairBatch.from_lists(pairs) if rank == 0 else None
    timers = {}
    out = shard.run_sharded_device(batch, "epi8", 64, _oracle_aligner("epi8", 1, 64, mtx, (-3, -2, 0, 0)), dist, device="cpu", nthreads=2, timers=timers)
    if rank == 0:
        res, st, ncg, dense, goff = out
        valid = np.array([len(a) > 0 and len(b) > 0 for a, b in pairs])
        vb = synth.PairBatch.from_lists([p for p, v in zip(pairs, valid) if v])
        exp_r, exp_c, _ = ck.oracle_batch("epi8", vb, 1, 64, mtx, (-3, -2, 0, 0))
        ok = st[7] == 16 and not res[7].any() and int(ncg[7]) == 0
        k = 0
        for i in range(len(pairs)):
            if not valid[i]:
                continue
            ok = ok and np.array_equal(res[i], exp_r[k]) and np.array_equal(dense[int(goff[i]):int(goff[i + 1])], exp_c[k]) and st[i] == 0
            k += 1
        ok = ok and timers["scatter_bytes"] > 0 and timers["gather_bytes"] > 0
        q.put(bool(ok))
    else:
        assert out is None
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
