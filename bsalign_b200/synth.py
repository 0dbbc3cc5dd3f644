"""Deterministic synthetic 4-base read batches.

Bases are uint8 codes 0..3 (the reference's one-base-per-byte `u1i` input, bsalign.h:399).  A target is
the query mutated per base with probabilities (p_sub, p_ins, p_del); an insertion keeps the base and
appends one random base after it.  Everything is vectorised numpy so the 100k x 1kb batch of
BASELINE.json config 2 builds in seconds; the same arrays feed the CUDA path, the oracle and the
reference arm, so all sides read identical bytes.
"""
import numpy as np


class PairBatch:
    """Packed batch: one uint8 arena + per-pair offsets/lengths (the layout the C-ABI takes)."""

    def __init__(self, seqs, qoff, qlen, toff, tlen):
        self.seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self.qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        self.qlen = np.ascontiguousarray(qlen, dtype=np.uint32)
        self.toff = np.ascontiguousarray(toff, dtype=np.uint64)
        self.tlen = np.ascontiguousarray(tlen, dtype=np.uint32)

    @property
    def n(self):
        return len(self.qlen)

    def query(self, i):
        return self.seqs[int(self.qoff[i]):int(self.qoff[i]) + int(self.qlen[i])]

    def target(self, i):
        return self.seqs[int(self.toff[i]):int(self.toff[i]) + int(self.tlen[i])]

    def subset(self, idx):
        idx = np.asarray(idx)
        return PairBatch(self.seqs, self.qoff[idx], self.qlen[idx], self.toff[idx], self.tlen[idx])

    @staticmethod
    def from_lists(pairs):
        """pairs: iterable of (query uint8 array, target uint8 array)."""
        chunks, qoff, qlen, toff, tlen, pos = [], [], [], [], [], 0
        for q, t in pairs:
            q = np.asarray(q, dtype=np.uint8)
            t = np.asarray(t, dtype=np.uint8)
            qoff.append(pos); qlen.append(len(q)); pos += len(q); chunks.append(q)
            toff.append(pos); tlen.append(len(t)); pos += len(t); chunks.append(t)
        seqs = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
        return PairBatch(seqs, qoff, qlen, toff, tlen)


def mutate_batch(rng, queries, p_sub, p_ins, p_del):
    """queries: (n, qlen) uint8.  Returns (flat targets, per-pair target lengths)."""
    n, qlen = queries.shape
    r = rng.random((n, qlen), dtype=np.float32)
    sub = r < p_sub
    ins = (r >= p_sub) & (r < p_sub + p_ins)
    dele = (r >= p_sub + p_ins) & (r < p_sub + p_ins + p_del)
    base = queries.copy()
    shift = rng.integers(1, 4, size=(n, qlen), dtype=np.uint8)
    base[sub] = (base[sub] + shift[sub]) & 3
    cnt = np.ones((n, qlen), dtype=np.int64)
    cnt[ins] = 2
    cnt[dele] = 0
    tlen = cnt.sum(axis=1)
    flat_cnt = cnt.ravel()
    out = np.repeat(base.ravel(), flat_cnt)
    # the second copy of every inserted base becomes a random base
    ends = np.cumsum(flat_cnt)
    ins_pos = ends[ins.ravel()] - 1
    out[ins_pos] = rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)
    return out, tlen.astype(np.uint32)


def make_pairs(n, qlen, seed, p_sub=0.03, p_ins=0.03, p_del=0.04):
    """n pairs of (random query of length qlen, mutated target)."""
    rng = np.random.default_rng(seed)
    queries = rng.integers(0, 4, size=(n, qlen), dtype=np.uint8)
    tflat, tlen = mutate_batch(rng, queries, p_sub, p_ins, p_del)
    # arena = all queries, then all targets
    seqs = np.concatenate([queries.ravel(), tflat])
    qoff = np.arange(n, dtype=np.uint64) * np.uint64(qlen)
    toff = np.uint64(n * qlen) + np.concatenate([[0], np.cumsum(tlen[:-1], dtype=np.uint64)]).astype(np.uint64)
    return PairBatch(seqs, qoff, np.full(n, qlen, np.uint32), toff, tlen)


def ont_like(total_error):
    """Error split 23:31:46 (sub:ins:del), example/ScriptsForPaper.txt:9 of the reference."""
    return (0.23 * total_error, 0.31 * total_error, 0.46 * total_error)


def score_matrix(match, mismatch):
    """Same layout as banded_striped_epi8_seqalign_set_score_matrix (bsalign.h:323): mtx[q*4+t]."""
    return np.array([mismatch if ((i ^ (i >> 2)) & 3) else match for i in range(16)], dtype=np.int8)
