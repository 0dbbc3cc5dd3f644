"""Synthetic inputs of the read re-alignment step (remsa_pedit_rd_bspoacore, bspoa.h:3916) in the reference's own layout, for timing
and for parity checks against the oracle at shapes no committed record covers (bench.py, tests).  What the reference builds from an
MSA (bspoa.h:4230-4450) is imitated: a consensus with gap columns, per-column base counts of `nreads` reads, one read in MSA
coordinates with substitutions / deletions / insertions into gap columns, and its homopolymer run counters."""
import numpy as np


class RemsaJob:
    pass


def make_job(mlen, bw=32, nreads=40, seed=0, p_err=0.08, p_gapcol=0.22):
    rng = np.random.default_rng(seed)
    hw = bw // 2
    sz1 = (mlen + bw + 15) // 16 * 16
    j = RemsaJob()
    j.mlen, j.bw, j.sz1 = mlen, bw, sz1
    j.szm = ((2 * mlen + 1) * (bw + 2) + 15) // 16 * 16
    cns = rng.integers(0, 4, mlen).astype(np.uint8)
    cns[rng.random(mlen) < p_gapcol] = 4
    j.seqs0 = np.full(sz1, 4, np.uint8)      # memset 4 (bspoa.h:4349)
    j.seqs1 = np.zeros(sz1, np.uint8)
    j.mats = np.zeros((2, 4, sz1), np.uint8)
    j.seqs1[hw:hw + mlen] = cns[::-1]
    # column profile: reads carry the consensus base with probability 1 - p_err, something else otherwise; gap columns are sparsely filled
    cnt = np.zeros((4, mlen), np.int64)
    for b in range(4):
        is_c = cns == b
        cnt[b, is_c] = rng.binomial(nreads, 1 - p_err, int(is_c.sum()))
        other = ~is_c
        cnt[b, other] = rng.binomial(nreads, np.where(cns[other] == 4, 0.01, p_err / 3))
    cnt = np.minimum(cnt, 255).astype(np.uint8)
    for b in range(4):
        j.mats[1, b, hw:hw + mlen] = cnt[b, ::-1]
    # the read in MSA coordinates
    lo, hi = int(rng.integers(0, 4)), mlen - int(rng.integers(0, 4))
    r = rng.random(mlen)
    base = cns.copy()
    sub = (cns < 4) & (r < p_err / 2)
    base[sub] = (cns[sub] + rng.integers(1, 4, int(sub.sum()))) & 3
    base[(cns < 4) & (r >= p_err / 2) & (r < p_err)] = 4                       # deletion
    ins = (cns == 4) & (r < 0.03)
    base[ins] = rng.integers(0, 4, int(ins.sum()))
    base[:lo] = 4; base[hi:] = 4
    pos = np.nonzero(base < 4)[0]
    j.seqs0[hw + pos] = base[pos]
    j.rdlen = len(pos)
    j.mbeg, j.mend = int(pos[0]), int(pos[-1]) + 1
    # homopolymer run counters, walking the read backwards (bspoa.h:4432-4446)
    lc, cc = 4, 0
    for p in pos[::-1]:
        b = int(base[p])
        if b == lc:
            cc = min(cc + 1, 255)
            j.mats[0, b, hw + p] = cc
        else:
            lc, cc = b, 0
    return j
