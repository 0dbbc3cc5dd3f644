"""Synthetic POA sweep jobs without running any reference code.

A BSPOA job aligns read r against the graph of the previous nrec+1 reads (bspoa.h:2636-2642).  All reads are
independent mutations of one random template, so the graph the reference would have built from their (correct)
alignments is known in closed form: one node per distinct (template column, base) -- substituted bases of a column are
alternative nodes, inserted bases get their own columns behind the template position -- and one edge per distinct pair
of consecutive nodes on a read, plus head/tail links.  Band offsets follow the reference's rule
rpos = clamp(rmap[column] - bw/2, 0, slen - bw) (bspoa.h:2168-2174) with rmap taken from the query's true alignment.
The result has the shape of the graphs the reference sweeps (a backbone with ~1.8 nodes per column and bubbles
at every error); it is an INPUT generator only -- parity is always checked against the oracle on the same jobs.
"""
import numpy as np

from . import poa

KMAX = 4  # inserted bases kept per template position (the mutation model inserts at most one)


def _mutate(rng, tmpl, p_sub, p_ins, p_del):
    """One read: returns (bases, column key per base, template position per base)."""
    L = len(tmpl)
    r = rng.random(L, dtype=np.float32)
    sub = r < p_sub
    ins = (r >= p_sub) & (r < p_sub + p_ins)
    dele = (r >= p_sub + p_ins) & (r < p_sub + p_ins + p_del)
    base = tmpl.copy()
    base[sub] = (base[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    cnt = np.ones(L, dtype=np.int64)
    cnt[ins] = 2
    cnt[dele] = 0
    pos = np.repeat(np.arange(L, dtype=np.int64), cnt)
    seq = np.repeat(base, cnt)
    ends = np.cumsum(cnt)
    ins_at = ends[ins] - 1
    k = np.zeros(len(seq), dtype=np.int64)
    k[ins_at] = 1
    seq[ins_at] = rng.integers(0, 4, size=len(ins_at), dtype=np.uint8)
    key = (pos * KMAX + k) * 4 + seq.astype(np.int64)
    return seq.astype(np.uint8), key, pos


def make_sweep_job(seed, tlen=15000, ngraph=21, par=None, p_sub=0.03, p_ins=0.03, p_del=0.04):
    """One steady-state sweep job: a new read against the graph of `ngraph` earlier reads of the same template."""
    rng = np.random.default_rng(seed)
    d = dict(poa.DEFAULT_BSPOA_PAR)
    if par:
        d.update(par)
    tmpl = rng.integers(0, 4, size=tlen, dtype=np.uint8)
    keys = []
    for _ in range(ngraph):
        _s, key, _p = _mutate(rng, tmpl, p_sub, p_ins, p_del)
        keys.append(key)
    query, _qkey, qpos_t = _mutate(rng, tmpl, p_sub, p_ins, p_del)
    slen = len(query)
    bw = min(int(d["bandwidth"]), slen) if d["bandwidth"] else slen
    bw = (bw + 15) // 16 * 16
    allk = np.concatenate(keys)
    uniq, first = np.unique(allk, return_index=True)
    order = np.argsort(first, kind="stable")               # node ids in first-occurrence order (the walk of sel_nodes_bspoa)
    nid_of_sorted = np.empty(len(uniq), dtype=np.int64)
    nid_of_sorted[order] = np.arange(len(uniq)) + 2         # 0 = head, 1 = tail
    nn = len(uniq) + 2
    ukey = uniq[order]
    # edges
    src, dst = [], []
    for key in keys:
        ids = nid_of_sorted[np.searchsorted(uniq, key)]
        src.append(np.concatenate([[0], ids])); dst.append(np.concatenate([ids, [1]]))
    src = np.concatenate(src); dst = np.concatenate(dst)
    e, ecov = np.unique(src * nn + dst, return_counts=True)    # coverage = reads that use the edge (bspoaedge_t.cov)
    src, dst = e // nn, e % nn
    eoff = np.zeros(nn + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=nn), out=eoff[1:])
    nct = np.bincount(dst, minlength=nn)
    base = np.zeros(nn, dtype=np.uint8); base[2:] = (ukey & 3).astype(np.uint8); base[:2] = 4
    col_t = np.zeros(nn, dtype=np.int64); col_t[2:] = ukey // (4 * KMAX); col_t[1] = tlen
    is_ins = np.zeros(nn, dtype=bool); is_ins[2:] = ((ukey // 4) % KMAX) != 0
    bonus = np.zeros(nn, dtype=np.uint8)
    bonus[2:] = ((~is_ins[2:]) & (base[2:] == tmpl[np.minimum(col_t[2:], tlen - 1)])).astype(np.uint8)   # nodes on the consensus
    # rmap: query position of every template position (first query base derived from it, else the next one)
    rmap = np.searchsorted(qpos_t, np.arange(tlen + 1), side="left")
    rpos = rmap[col_t] - bw // 2
    rpos = np.where(rpos < 0, 0, rpos)
    if bw >= slen:
        rpos[:] = 0
    else:
        rpos = np.minimum(rpos, slen - bw)
    rpos[0] = 0
    par_arr = np.array([bw] + [int(d[k]) for k in poa.PAR_FIELDS[1:]], dtype=np.int32)
    # reverse edges (the erev lists the traceback walks): grouped by destination, sources in ascending id order
    ro = np.argsort(dst * nn + src, kind="stable")
    reoff = np.zeros(nn + 1, dtype=np.int64)
    np.cumsum(nct, out=reoff[1:])
    return poa.SweepJob(par_arr, query, base, bonus, rpos.astype(np.int32), nct.astype(np.int32), eoff.astype(np.int32), dst.astype(np.int32), 0, 1,
                        reoff=reoff.astype(np.int32), resrc=src[ro].astype(np.int32), recov=ecov[ro].astype(np.int32))
