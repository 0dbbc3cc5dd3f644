"""Golden vectors of kmer_striped_seqedit_pairwise (bsalign.h:1209) from the UNMODIFIED reference (oracle/_ref/libbsref.so:
bsref_kmer_edit_batch).  Run in the build container only:   python tests/golden/make_kmer_golden.py  ->  kmer_golden.npz
Pairs: related pairs of 20..3000 bases at several error rates, clipped ends, unrelated pairs (no anchors: plain global edit), an empty one."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck
from bsalign_b200 import synth

KS = (13, 9, 15, 4)


def pairs(seed=21):
    rng = np.random.default_rng(seed)
    out = []
    for n, qlen, ps, pi, pd in ((24, 300, .03, .03, .04), (10, 1000, .05, .05, .05), (3, 3000, .02, .02, .02), (12, 100, .1, .1, .1), (12, 60, .02, .02, .02),
                                (8, 500, .2, .1, .1), (10, 20, .05, 0, 0), (6, 200, 0, 0, 0)):
        q = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
        t, tl = synth.mutate_batch(rng, q, ps, pi, pd)
        o = 0
        for i in range(n):
            out.append((q[i].copy(), t[o:o + tl[i]].copy()))
            o += tl[i]
    for i in range(0, len(out), 3):
        a, b = out[i]
        if len(b) > 30:
            out[i] = (a, b[int(rng.integers(0, 15)):len(b) - int(rng.integers(0, 15))])
    out += [(rng.integers(0, 4, int(rng.integers(1, 80))).astype(np.uint8), rng.integers(0, 4, int(rng.integers(1, 80))).astype(np.uint8)) for _ in range(8)]
    out.insert(5, (np.zeros(0, np.uint8), np.array([1, 2, 3], np.uint8)))
    core = rng.integers(0, 4, 1500).astype(np.uint8)
    gap = np.concatenate([core[:200], rng.integers(0, 4, 900).astype(np.uint8), core[1200:]])
    out += [(core, gap), (gap, core)]
    return out


def main():
    assert ck.have_ref()
    batch = synth.PairBatch.from_lists(pairs())
    z = {"seqs": batch.seqs, "qoff": batch.qoff, "qlen": batch.qlen, "toff": batch.toff, "tlen": batch.tlen, "ks": np.array(KS, np.int32)}
    for ci, k in enumerate(KS):
        res, cigs, _ = ck.kmer_batch("ref", batch, k)
        z["res%d" % ci] = res
        z["ncig%d" % ci] = np.array([len(c) for c in cigs], np.uint32)
        z["cig%d" % ci] = np.concatenate(cigs)
        print("k", k, "pairs", batch.n, "cigar words", len(z["cig%d" % ci]))
    np.savez_compressed(os.path.join(HERE, "kmer_golden.npz"), **z)


if __name__ == "__main__":
    main()
