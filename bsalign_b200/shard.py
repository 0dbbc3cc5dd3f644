"""Multi-GPU sharding of a batch of independent pairs.

Pairs never interact, so there is no collective on the DP path: every rank aligns its own shard on its own
GPU.  torch.distributed is used only as plumbing around it: the balanced partition is computed identically on
every rank from the pair lengths, and results travel back to rank 0 with one gather of fixed-size records
plus one gather of the dense cigar words (NCCL over NVLink when the tensors live on GPUs, gloo in the CPU tests).
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
    order = np.argsort(-np.asarray(work, dtype=np.int64), kind="stable")
    pos = np.arange(len(order))
    rnd, k = pos // nparts, pos % nparts
    part = np.where(rnd % 2 == 0, k, nparts - 1 - k)
    return [np.sort(order[part == p]) for p in range(nparts)]


def shard(batch, kind, bandwidth, rank, world):
    idx = balanced_partition(pair_work(batch, kind, bandwidth), world)[rank]
    return batch.subset(idx), idx


def gather_to_rank0(results, status, cigars, idx, n_total, dist, device="cpu"):
    """results (n_local,10) int32, status (n_local,) int32, cigars: list of uint32 arrays, idx: global pair ids.
    Rank 0 returns (results[n_total,10], status[n_total], list of n_total cigar arrays); other ranks return None."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    ncig = np.array([len(c) for c in cigars], dtype=np.int64)
    rec = np.concatenate([idx.astype(np.int64)[:, None], results.astype(np.int64), status.astype(np.int64)[:, None], ncig[:, None]], axis=1)
    dense = np.concatenate(cigars).astype(np.int64) if len(cigars) and ncig.sum() else np.zeros(0, np.int64)
    sizes = torch.tensor([rec.shape[0], dense.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    max_rec = int(max(s[0].item() for s in all_sizes))
    max_den = int(max(s[1].item() for s in all_sizes))
    rec_t = torch.zeros((max_rec, 13), dtype=torch.int64, device=device)
    rec_t[:rec.shape[0]] = torch.from_numpy(rec).to(device)
    den_t = torch.zeros(max(max_den, 1), dtype=torch.int64, device=device)
    den_t[:dense.shape[0]] = torch.from_numpy(dense).to(device)
    rec_all = [torch.zeros_like(rec_t) for _ in range(world)] if rank == 0 else None
    den_all = [torch.zeros_like(den_t) for _ in range(world)] if rank == 0 else None
    dist.gather(rec_t, rec_all, dst=0)
    dist.gather(den_t, den_all, dst=0)
    if rank != 0:
        return None
    out_res = np.zeros((n_total, 10), dtype=np.int32)
    out_st = np.zeros(n_total, dtype=np.int32)
    out_cg = [None] * n_total
    for r in range(world):
        nr = int(all_sizes[r][0].item())
        rr = rec_all[r][:nr].cpu().numpy()
        dd = den_all[r].cpu().numpy()
        off = 0
        for row in rr:
            g = int(row[0])
            out_res[g] = row[1:11]
            out_st[g] = row[11]
            k = int(row[12])
            out_cg[g] = dd[off:off + k].astype(np.uint32)
            off += k
    return out_res, out_st, out_cg


def run_sharded(batch, kind, bandwidth, align_fn, dist, device="cpu"):
    """align_fn(sub_batch) -> (results, status, cigars).  In production align_fn is Context.epi8_batch / edit_batch
    on this rank's GPU; the CPU tests inject the oracle."""
    sub, idx = shard(batch, kind, bandwidth, dist.get_rank(), dist.get_world_size())
    res, st, cg = align_fn(sub)
    return gather_to_rank0(res, st, cg, idx, batch.n, dist, device)


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
