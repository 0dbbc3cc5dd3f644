"""GPU parity sweep of the wavefront forward kernel (full-band affine batches) against the oracle (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import checkers as ck
from bsalign_b200 import api, synth

ctx = api.Context(0)
tot = bad = 0

def cmp(batch, mode, bw, mtx, gaps, tag, nthreads=8):
    global tot, bad
    errs = np.zeros(batch.n, dtype=np.int32)
    exp, ecg, _ = ck.oracle_batch("epi8", batch, mode, bw, mtx, gaps, errs=errs, nthreads=nthreads)
    got = ctx.epi8_batch(batch, mode, bw, mtx, *gaps)
    gcg = got.cigars()
    nb = 0; first = None
    for i in range(batch.n):
        if errs[i]:
            continue
        ok = np.array_equal(got.results[i], exp[i]) and np.array_equal(gcg[i], ecg[i]) and got.status[i] == 0
        if not ok:
            nb += 1
            if first is None: first = i
    tot += batch.n; bad += nb
    msg = "%-46s n=%d bad=%d flagged=%d fwd launches %d" % (tag, batch.n, nb, int((errs != 0).sum()), ctx.timing()["forward_launches"])
    if first is not None:
        i = first
        msg += "\n   first bad pair %d qlen=%d tlen=%d status=%d\n   gpu %s\n   exp %s\n   ncig gpu %d exp %d" % (
            i, batch.qlen[i], batch.tlen[i], got.status[i], got.results[i], exp[i], len(gcg[i]), len(ecg[i]))
    print(msg, flush=True)

m = synth.score_matrix(2, -6)
G = (-3, -2, 0, 0)
rng = np.random.default_rng(11)
for mode in (0, 1, 2):
    for qlen, n in ((5, 16), (17, 32), (40, 64), (130, 64), (300, 64), (1000, 64), (1001, 32), (2048, 12), (3000, 8)):
        cmp(synth.make_pairs(n, qlen, seed=mode * 1000 + qlen), mode, 0, m, G, "mode%d full qlen%d" % (mode, qlen))
    # explicit bandwidth no shorter than the longest query is a full band too
    cmp(synth.make_pairs(32, 200, seed=77 + mode), mode, 256, m, G, "mode%d bw256 qlen200" % mode)
# unrelated / asymmetric pairs
for it in range(6):
    pairs = []
    for k in range(24):
        ql = int(rng.integers(1, 700)); tl = int(rng.integers(1, 700))
        pairs.append((rng.integers(0, 4, ql).astype(np.uint8), rng.integers(0, 4, tl).astype(np.uint8)))
    cmp(synth.PairBatch.from_lists(pairs), it % 3, 0, m, G, "random unrelated set %d mode %d" % (it, it % 3))
# other score sets (saturation regimes must come back through the redo path or stay exact)
for (Mv, Xv), gaps in (((30, -40), (-40, -20, 0, 0)), ((5, -4), (-10, -1, 0, 0)), ((1, -1), (-1, 0, 0, 0)), ((60, -60), (-60, -3, 0, 0)), ((2, -2), (-4, -2, 0, 0))):
    for mode in (0, 1, 2):
        cmp(synth.make_pairs(32, 400, seed=Mv + mode), mode, 0, synth.score_matrix(Mv, Xv), gaps, "M%d X%d gaps%s mode%d" % (Mv, Xv, gaps[:2], mode))
# homopolymers / long indels
pairs = []
for k in range(16):
    q = np.full(500, k % 4, np.uint8); t = np.concatenate([q[:200], rng.integers(0, 4, 150).astype(np.uint8), q[200:]])
    pairs.append((q, t)); pairs.append((t, q))
cmp(synth.PairBatch.from_lists(pairs), 0, 0, m, G, "homopolymer + long indel global")
cmp(synth.PairBatch.from_lists(pairs), 1, 0, m, G, "homopolymer + long indel overlap")
cmp(synth.make_pairs(4, 10000, seed=5), 0, 0, m, G, "10kb global full (anchors)")
cmp(synth.make_pairs(4, 6000, seed=6), 1, 0, m, G, "6kb overlap full (anchors)")
print("TOTAL", tot, "BAD", bad)
# forced splits (sub-blocks per lane): long pairs, every mode
for split in (2, 4):
    os.environ["BSB200_WAVE_SPLIT"] = str(split)
    for mode in (0, 1, 2):
        for qlen, n in ((1200 if split == 2 else 6500, 6), (10000, 3)):
            cmp(synth.make_pairs(n, qlen, seed=split * 100 + mode * 10 + qlen % 7), mode, 0, m, G, "split%d mode%d full qlen%d" % (split, mode, qlen))
    pairs = []
    for k in range(4):
        ql = int(rng.integers(6500, 9000)); tl = int(rng.integers(3, 9000))
        pairs.append((rng.integers(0, 4, ql).astype(np.uint8), rng.integers(0, 4, tl).astype(np.uint8)))
    pairs.append((rng.integers(0, 4, 7000).astype(np.uint8), rng.integers(0, 4, 2).astype(np.uint8)))
    cmp(synth.PairBatch.from_lists(pairs), 0, 0, m, G, "split%d unrelated long, short targets" % split)
    cmp(synth.PairBatch.from_lists(pairs), 1, 0, m, G, "split%d unrelated long overlap" % split)
os.environ.pop("BSB200_WAVE_SPLIT")
print("TOTAL", tot, "BAD", bad)
