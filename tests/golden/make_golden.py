"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref/libbsref.so,
built from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed):
  epi8_golden.npz   pairs x configs for banded_striped_epi8_seqalign_pairwise (bsalign.h:3854)
  edit_golden.npz   pairs x configs for striped_seqedit_pairwise (bsalign.h:1046)
  readme_pair.npz   the pair 29.1/29.2 of example/real.ont.b10M.txt whose result is printed in README.md:36-42
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck
from bsalign_b200 import synth

assert ck.have_ref(), "oracle/_ref/libbsref.so missing: run `make -C oracle` where /root/reference exists"


def pack(batch, configs, kind):
    out = {"seqs": batch.seqs, "qoff": batch.qoff, "qlen": batch.qlen, "toff": batch.toff, "tlen": batch.tlen,
           "configs": np.array(configs, dtype=np.int32)}
    for ci, cfg in enumerate(configs):
        # pairs on which the reference itself reads out of bounds / never terminates (the oracle flags them, see
        # DESIGN.md "parity domain") are masked out: valid%d marks the pairs that carry a golden answer
        errs = np.zeros(batch.n, dtype=np.int32)
        if kind == "epi8":
            mode, bw, M, X, go1, ge1, go2, ge2 = cfg
            ck.oracle_batch("epi8", batch, mode, bw, synth.score_matrix(M, X), (go1, ge1, go2, ge2), errs=errs)
            keep = np.nonzero(errs == 0)[0]
            got = ck.forked(ck.ref_batch, "epi8", batch.subset(keep), mode, bw, synth.score_matrix(M, X), (go1, ge1, go2, ge2))
        else:
            mode, bw = cfg
            ck.oracle_batch("edit", batch, mode, bw, errs=errs)
            keep = np.nonzero(errs == 0)[0]
            got = ck.forked(ck.ref_batch, "edit", batch.subset(keep), mode, bw)
        assert got is not None, "reference crashed on config %r" % (cfg,)
        sub_res, sub_cigs, _ = got
        res = np.zeros((batch.n, 10), dtype=np.int32)
        cigs = [np.zeros(0, np.uint32)] * batch.n
        for k, i in enumerate(keep):
            res[i] = sub_res[k]; cigs[i] = sub_cigs[k]
        out["valid%d" % ci] = (errs == 0)
        out["res%d" % ci] = res
        out["ncig%d" % ci] = np.array([len(c) for c in cigs], dtype=np.uint32)
        out["cig%d" % ci] = np.concatenate(cigs) if cigs else np.zeros(0, np.uint32)
        print(kind, cfg, "valid pairs", len(keep), "of", batch.n, flush=True)
    return out


def mixed_pairs(seed):
    rng = np.random.default_rng(seed)
    pairs = []
    for qlen, err in [(1, .1), (3, .1), (16, .1), (17, .2), (33, .1), (64, .05), (75, .1), (100, .1), (150, .2), (200, .05), (255, .1), (256, .1), (300, .1), (400, .15)]:
        for _ in range(3):
            q = rng.integers(0, 4, qlen).astype(np.uint8)
            t, _ = synth.mutate_batch(rng, q[None, :], err * .3, err * .3, err * .4)
            if len(t) == 0:
                t = q.copy()
            pairs.append((q, t))
    # homopolymer runs and a long deletion
    q = np.repeat(rng.integers(0, 4, 60).astype(np.uint8), 5)
    t, _ = synth.mutate_batch(rng, q[None, :], .02, .04, .04)
    pairs.append((q, t))
    q = rng.integers(0, 4, 350).astype(np.uint8)
    pairs.append((q, np.concatenate([q[:100], q[140:]])))
    pairs.append((np.concatenate([q[:200], q[230:]]), q))
    return synth.PairBatch.from_lists(pairs)


def main():
    batch = mixed_pairs(2024)
    epi8_cfgs = []
    for mode in (0, 1, 2):
        for bw in (0, 32, 128):
            epi8_cfgs.append((mode, bw, 2, -6, -3, -2, 0, 0))       # affine (CLI defaults, main.c:264)
        epi8_cfgs.append((mode, 64, 2, -2, -4, -2, 0, 0))           # example/run.sh scores
        epi8_cfgs.append((mode, 0, 2, -6, 0, -2, 0, 0))             # linear gaps
        epi8_cfgs.append((mode, 64, 2, -6, -3, -2, -8, -1))         # two-piece (DEFAULT_BSPOA_PAR, bspoa.h:75-77)
    epi8_cfgs.append((0, 48, 30, -40, -40, -20, 0, 0))              # int8 saturation binds
    np.savez_compressed(os.path.join(HERE, "epi8_golden.npz"), **pack(batch, epi8_cfgs, "epi8"))
    edit_cfgs = [(mode, bw) for mode in (0, 1, 2) for bw in (0, 64, 128)]
    np.savez_compressed(os.path.join(HERE, "edit_golden.npz"), **pack(batch, edit_cfgs, "edit"))
    # README example: pair 29.1 / 29.2 of the reference's example file
    code = {c: i for i, c in enumerate("ACGT")}
    seqs, name = {}, None
    with open("/root/reference/example/real.ont.b10M.txt") as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                name = line[1:]
            elif name in ("29.1", "29.2"):
                seqs[name] = np.array([code[c] & 3 for c in line.upper()], dtype=np.uint8)
    b = synth.PairBatch.from_lists([(seqs["29.1"], seqs["29.2"])])
    W = (len(seqs["29.1"]) + 15) // 16 * 16     # main.c:315: -W 0 means roundup16(first sequence)
    res, cigs, _ = ck.ref_batch("epi8", b, 1, W, synth.score_matrix(2, -2), (-4, -2, 0, 0))  # run.sh: -M 2 -X 2 -O 4 -E 2, default mode overlap
    np.savez_compressed(os.path.join(HERE, "readme_pair.npz"), q=seqs["29.1"], t=seqs["29.2"], bandwidth=W, res=res[0], cig=cigs[0])
    print("README pair:", res[0], "(README.md:36 says score 128, 71 matches, 4 mismatches, 0 deletions, 1 insertion)")


if __name__ == "__main__":
    main()
