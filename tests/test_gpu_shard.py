"""The multi-GPU path on real devices (-m gpu): rank 0 owns the batch, compact shard arenas travel to the ranks' GPUs, every rank aligns
its shard in place through bsb200_batch_upload_dev / bsb200_batch_fetch_dense_dev, records + dense cigars travel back and are merged into
pair order.  NCCL with one rank per GPU when the box has more than one, else a single rank (same C-ABI path, no transfers)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import checkers as ck
from bsalign_b200 import api, shard, synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world, **({"device_id": dev} if world > 1 else {}))
    mtx = synth.score_matrix(2, -6)
    ctx = api.Context(rank)
    ok = True
    for kind, mode, bw, n, qlen in (("epi8", 0, 0, 301, 400), ("epi8", 1, 128, 200, 700), ("edit", 0, 64, 500, 300)):
        batch = synth.make_pairs(n, qlen, seed=n) if rank == 0 else None
        timers = {}
        # (the edit case packs the shards on the GPU: bsb200_pack_pairs_dev; the others by host threads)
        out = shard.run_sharded_device(batch, kind, bw, shard.cuda_aligner(ctx, kind, mode, bw, mtx, (-3, -2, 0, 0)), dist, device=dev, nthreads=4, timers=timers,
                                       packer=shard.cuda_packer(ctx) if (kind == "edit" or mode == 1) else None)
        if rank == 0:
            res, st, ncg, dense, goff = out
            exp, ecg, _ = ck.oracle_batch(kind, batch, mode, bw, mtx, (-3, -2, 0, 0), nthreads=8)
            ok = ok and np.array_equal(res, exp) and not st.any()
            ok = ok and all(np.array_equal(dense[int(goff[i]):int(goff[i + 1])], ecg[i]) for i in range(batch.n))
            if world > 1:
                ok = ok and timers["scatter_bytes"] > 0 and timers["gather_bytes"] > 0
    ctx.close()
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


def test_sharded_batch_on_gpus_equals_oracle():
    world = min(2, torch.cuda.device_count())
    assert world >= 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    c = mp.get_context("spawn")
    q = c.Queue()
    procs = [c.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
