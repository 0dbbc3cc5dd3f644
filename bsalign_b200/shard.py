"""Multi-GPU sharding of a batch of independent pairs (BASELINE.json north_star: "sharded across the
8 GPUs of one box with NCCL over NVLink only to scatter the input batch and gather results").

Pairs never interact, so there is no collective on the DP path.  The batch lives on rank 0: the pair lengths are broadcast,
every rank derives the same balanced partition from them, rank 0 packs each shard's sequences into a compact arena and sends it
to its rank's GPU, every rank aligns its shard in place (bsb200_batch_upload_dev) and sends fixed-size records plus its dense,
pair-ordered cigar words back (bsb200_batch_fetch_dense_dev); rank 0 merges them into pair order.  NCCL when the tensors live
on GPUs, gloo in the CPU tests (which inject the oracle as the aligner).
"""
import numpy as np


def pair_work(batch, kind, bandwidth):
    """Nominal DP cells per pair (the same bw_eff * tlen the GCUPS metric counts)."""
    q = batch.qlen.astype(np.int64)
    t = batch.tlen.astype(np.int64)
    if kind == "epi8":
        bw = np.where(bandwidth == 0, q, bandwidth)
        bw = (bw + 15) // 16 * 16
    else:
        bw = np.where((bandwidth == 0) | (bandwidth > q), (q + 63) // 64 * 64, (bandwidth + 63) // 64 * 64)
    return bw * t


def balanced_partition(work, nparts):
    """Deal pairs, heaviest first, in snake order (0..g-1, g-1..0, ...): per-part totals agree within a fraction
    of one heavy pair.  Returns a list of index arrays (each sorted ascending, so shards keep input order)."""
    work = np.asarray(work, dtype=np.int64)
    n = len(work)
    if n == 0:
        return [np.zeros(0, np.int64) for _ in range(nparts)]
    # the order only has to be roughly heaviest-first: keys quantised to 16 bits sort with numpy's radix sort (O(n)); ties keep input order
    wmax = int(work.max())
    sh = max(0, wmax.bit_length() - 16)
    key = (np.uint16(0xFFFF) - (work >> sh).astype(np.uint16)) if wmax else np.zeros(n, np.uint16)
    order = np.argsort(key, kind="stable")
    pos = np.arange(n)
    rnd, k = pos // nparts, pos % nparts
    part_of_pos = np.where(rnd % 2 == 0, k, nparts - 1 - k)
    part = np.empty(n, dtype=np.int64)
    part[order] = part_of_pos                      # part of every pair, in pair order
    return [np.flatnonzero(part == p) for p in range(nparts)]


# ---- the product path: one batch on rank 0, shards over NVLink, results back to rank 0 ----------------------------------------
def plan_shards(qlen, tlen, kind, bandwidth, world):
    """Every rank computes the same plan from the broadcast pair lengths: per rank the global pair ids of its shard (ascending) and
    the offsets of its pairs inside the shard's compact arena (query then target of each pair, in shard order)."""
    class _L:  # pair_work only reads lengths
        pass
    b = _L(); b.qlen, b.tlen = qlen, tlen
    parts = balanced_partition(pair_work(b, kind, bandwidth), world)
    plans = []
    for idx in parts:
        ql, tl = qlen[idx].astype(np.uint64), tlen[idx].astype(np.uint64)
        pos = np.zeros(len(idx) + 1, dtype=np.uint64)
        np.cumsum(ql + tl, out=pos[1:])
        plans.append(dict(idx=idx, qoff=pos[:-1].copy(), toff=pos[:-1] + ql, qlen=np.ascontiguousarray(qlen[idx], dtype=np.uint32),
                          tlen=np.ascontiguousarray(tlen[idx], dtype=np.uint32), nbytes=int(pos[-1])))
    return plans


class ShardView:
    """The offset / length tables of one shard (host arrays) over its compact arena; quacks like synth.PairBatch for the C ABI."""

    def __init__(self, plan, seqs=None):
        self.qoff, self.qlen, self.toff, self.tlen, self.seqs = plan["qoff"], plan["qlen"], plan["toff"], plan["tlen"], seqs

    @property
    def n(self):
        return len(self.qlen)


def cuda_aligner(ctx, kind, mode, bandwidth, mtx=None, gaps=(0, 0, 0, 0)):
    """align_fn for run_sharded_device on a GPU rank: the arena that arrived over NVLink is aligned in place (bsb200_batch_upload_dev),
    results and the dense pair-ordered cigars stay in device tensors for the way back (bsb200_batch_fetch_dense_dev)."""
    import torch

    def align(arena, view):
        n = view.n
        dev = arena.device
        rec = torch.zeros((n, 12), dtype=torch.int32, device=dev)
        if n == 0:
            return rec, torch.zeros(0, dtype=torch.int32, device=dev), {}
        rb = ctx.upload_dev(kind, arena.data_ptr(), view, mode, bandwidth, mtx, gaps, want_cigar=True)
        try:
            rb.run()
            tm = ctx.timing()
            res = torch.empty((n, 10), dtype=torch.int32, device=dev)
            st = torch.empty(n, dtype=torch.int32, device=dev)
            ncg = torch.empty(n, dtype=torch.int32, device=dev)
            total, _ = rb.fetch_dense_dev(res.data_ptr(), 0, 0, ncg.data_ptr(), st.data_ptr())
            cig = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            rb.fetch_dense_dev(0, cig.data_ptr(), cig.numel(), 0, 0)
        finally:
            rb.free()
        rec[:, :10] = res
        rec[:, 10] = st
        rec[:, 11] = ncg
        return rec, cig[:total], tm
    return align


def cuda_packer(ctx):
    """packer for run_sharded_device on GPUs: rank 0's whole arena goes to its GPU in ONE copy (issued before the plan, so it travels while
    the lengths are broadcast and the shards planned) and a kernel gathers every shard's compact arena there (bsb200_pack_pairs_dev) -
    instead of a gather by host threads followed by the same copy."""
    from . import api

    def pack(d_src_ptr, batch, idx_all, d_dst_ptr):
        api.pack_pairs_dev(ctx, d_src_ptr, batch, idx_all, d_dst_ptr)
    return pack


def run_sharded_device(batch, kind, bandwidth, align_fn, dist, device="cpu", pinned=None, nthreads=8, timers=None, packer=None):
    """One batch on rank 0 (`batch` is None elsewhere) -> shards of equal DP cells -> every rank aligns its shard on its own device ->
    rank 0 gets (results[n,10] int32, status[n] int32, ncigar[n] uint32, dense pair-ordered cigar words uint32).  Other ranks return None.

    Collectives (NCCL over NVLink when device is a cuda device, gloo in the CPU tests): one broadcast of the pair lengths, one send per
    rank of its compact sequence arena (rank 0 packs it in host memory and moves it to its own GPU first - or, with `packer`, moves the
    whole arena and packs on the GPU), one all_gather of the cigar
    word counts, and per rank one send of its fixed-size records and one of its cigar words.  Nothing is exchanged during the DP.
    align_fn(arena tensor on `device`, ShardView) -> (records int32 [n_r, 12] = 10 result ints, status, ncigar; cigar words as int32; timing dict)
    timers: optional dict that receives wall-clock milliseconds of the stages on this rank."""
    import time
    import torch
    from . import api
    world, rank = dist.get_world_size(), dist.get_rank()
    t0 = time.perf_counter()
    lap = {}

    def mark(name):
        if str(device).startswith("cuda"):
            torch.cuda.synchronize()
        lap[name] = (time.perf_counter() - t0) * 1e3
    src_d = None
    if rank == 0 and packer is not None and batch.n:
        src_d = torch.from_numpy(batch.seqs).to(device, non_blocking=True)   # the whole arena starts crossing PCIe now
    # ---- pair lengths to every rank -----------------------------------------------------------------------------------------
    hdr = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        hdr[0] = batch.n
    if world > 1:
        dist.broadcast(hdr, 0)
    n = int(hdr.item())
    lens = torch.zeros((2, n), dtype=torch.int32, device=device)
    if rank == 0:
        lens[0] = torch.from_numpy(batch.qlen.astype(np.int32)).to(device)
        lens[1] = torch.from_numpy(batch.tlen.astype(np.int32)).to(device)
    if world > 1:
        dist.broadcast(lens, 0)
    lens_h = lens.cpu().numpy()
    qlen, tlen = lens_h[0].astype(np.uint32), lens_h[1].astype(np.uint32)
    plans = plan_shards(qlen, tlen, kind, bandwidth, world)
    mine = plans[rank]
    mark("plan")
    # ---- scatter: rank 0 packs every shard's compact arena, moves it to its GPU and sends it on --------------------------------
    sends, keep = [], []
    arena = None
    scatter_bytes = 0
    if rank == 0:
        # one pass over the batch packs every shard's compact arena, shard after shard in one (pinned) host arena; one copy moves it to
        # rank 0's GPU, and the shards of the other ranks are sent on from there
        idx_all = np.concatenate([p["idx"] for p in plans]).astype(np.uint64)
        total = int(sum(p["nbytes"] for p in plans))
        if src_d is not None:
            whole = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
            torch.cuda.current_stream().synchronize()
            packer(src_d.data_ptr(), batch, idx_all, whole.data_ptr())
            keep.append((src_d, whole))
        else:
            host = pinned(max(total, 1)) if pinned is not None else None
            pb, nb = api.pack_pairs(batch, idx_all, out_seqs=host, nthreads=nthreads)
            whole = torch.from_numpy(pb.seqs[:max(nb, 1)]).to(device, non_blocking=True)
            keep.append((pb, whole))
        off = 0
        for r in range(world):
            nbr = plans[r]["nbytes"]
            t = whole[off:off + max(nbr, 1)] if nbr else torch.zeros(1, dtype=torch.uint8, device=device)
            off += nbr
            if r == 0:
                arena = t
            else:
                sends.append(dist.isend(t, r))
                scatter_bytes += nbr
    else:
        arena = torch.empty(max(mine["nbytes"], 1), dtype=torch.uint8, device=device)
        dist.recv(arena, 0)
    mark("scatter_issued")
    # ---- this rank's shard ------------------------------------------------------------------------------------------------------
    if str(device).startswith("cuda"):
        torch.cuda.current_stream().synchronize()
    rec, cig, tm = align_fn(arena, ShardView(mine, None))
    for w in sends:
        w.wait()
    mark("aligned")
    # ---- gather ---------------------------------------------------------------------------------------------------------------------
    cnt = torch.tensor([cig.numel()], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    if world > 1:
        dist.all_gather(counts, cnt)
    else:
        counts = [cnt]
    counts = [int(c.item()) for c in counts]
    if rank != 0:
        dist.send(rec.contiguous(), 0)
        if counts[rank]:
            dist.send(cig.contiguous(), 0)
        mark("gathered")
        if timers is not None:
            timers.update(lap); timers["kernel"] = tm
        return None
    recs, cigs = [rec], [cig]
    gather_bytes = 0
    for r in range(1, world):
        rr = torch.empty((len(plans[r]["idx"]), 12), dtype=torch.int32, device=device)
        dist.recv(rr, r)
        cc = torch.empty(counts[r], dtype=torch.int32, device=device)
        if counts[r]:
            dist.recv(cc, r)
        recs.append(rr); cigs.append(cc)
        gather_bytes += rr.numel() * 4 + cc.numel() * 4
    mark("gathered")
    # ---- merge into pair order on rank 0's device (index arithmetic on tensors), then one copy per array to the host -------------
    idx_all = torch.from_numpy(np.concatenate([p["idx"] for p in plans]).astype(np.int64)).to(device)
    rec_all = torch.cat(recs, 0)
    rec_g = torch.empty_like(rec_all)
    rec_g[idx_all] = rec_all                                  # record of pair i at row i
    ncg_g = rec_g[:, 11].to(torch.int64)
    goff_d = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(ncg_g, 0, out=goff_d[1:])
    total = int(goff_d[-1].item())
    dense_d = torch.empty(max(total, 1), dtype=torch.int32, device=device)
    if total:
        cig_all = torch.cat(cigs, 0)
        ln = rec_all[:, 11].to(torch.int64)                   # words per pair, in shard order (= order of cig_all)
        soff = torch.cumsum(ln, 0) - ln
        shift = goff_d[:-1][idx_all] - soff                   # where a pair's words move: destination - source offset
        dst = torch.repeat_interleave(shift, ln) + torch.arange(total, dtype=torch.int64, device=device)
        dense_d[dst] = cig_all
    if pinned is not None and str(device).startswith("cuda"):
        rec_h = torch.from_numpy(pinned(rec_g.numel() * 4)[:rec_g.numel() * 4].view(np.int32)).view(n, 12)
        den_h = torch.from_numpy(pinned(max(total, 1) * 4)[:max(total, 1) * 4].view(np.int32))
        rec_h.copy_(rec_g, non_blocking=True); den_h.copy_(dense_d, non_blocking=True)
        torch.cuda.synchronize()
        rec_np, dense = rec_h.numpy(), den_h.numpy().view(np.uint32)
    else:
        rec_np, dense = rec_g.cpu().numpy(), dense_d.cpu().numpy().view(np.uint32)
    results = np.ascontiguousarray(rec_np[:, :10])
    status = np.ascontiguousarray(rec_np[:, 10])
    ncigar = np.ascontiguousarray(rec_np[:, 11]).astype(np.uint32)
    status[(qlen == 0) | (tlen == 0)] |= 16   # BSB200_ST_EMPTY (a host-side flag of the C ABI)
    goff = goff_d.cpu().numpy().astype(np.uint64)
    mark("assembled")
    if timers is not None:
        timers.update(lap); timers["kernel"] = tm
        timers["scatter_bytes"] = scatter_bytes; timers["gather_bytes"] = gather_bytes
    return results, status, ncigar, dense[:int(goff[-1])], goff


# ---- POA sweep jobs (MSA jobs never interact either) -------------------------------------------------
def job_work(jobs):
    """Nominal work of a sweep job: out-edges (row updates) x band width."""
    return np.array([len(j.edst) * int(j.par[0]) for j in jobs], dtype=np.int64)


def shard_jobs(jobs, rank, world):
    """This rank's balanced share of a list of bsalign_b200.poa.SweepJob (and their global indices)."""
    idx = balanced_partition(job_work(jobs), world)[rank]
    return [jobs[i] for i in idx], idx


def gather_jobs_to_rank0(best, trace, idx, n_total, dist, device="cpu"):
    """best (n_local,3) int32, trace (n_local,8) int32 or None, idx: global job ids.  Rank 0 returns (best[n_total,3], trace[n_total,8])."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    tr = trace if trace is not None else np.zeros((len(idx), 8), dtype=np.int32)
    rec = np.concatenate([idx.astype(np.int64)[:, None], best.astype(np.int64), tr.astype(np.int64)], axis=1)
    size = torch.tensor([rec.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, size)
    mx = int(max(s.item() for s in sizes))
    rec_t = torch.zeros((max(mx, 1), 12), dtype=torch.int64, device=device)
    rec_t[:rec.shape[0]] = torch.from_numpy(rec).to(device)
    rec_all = [torch.zeros_like(rec_t) for _ in range(world)] if rank == 0 else None
    dist.gather(rec_t, rec_all, dst=0)
    if rank != 0:
        return None
    ob = np.zeros((n_total, 3), dtype=np.int32)
    ot = np.zeros((n_total, 8), dtype=np.int32)
    for r in range(world):
        a = rec_all[r][:int(sizes[r].item())].cpu().numpy()
        ob[a[:, 0]] = a[:, 1:4]
        ot[a[:, 0]] = a[:, 4:12]
    return ob, ot
