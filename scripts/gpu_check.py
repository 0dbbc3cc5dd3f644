"""Ad-hoc GPU parity sweep against the oracle (development aid; the real tests live in tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import checkers as ck
from bsalign_b200 import api, synth

ctx = api.Context(0)
tot = bad = 0

def cmp(kind, batch, mode, bw, mtx=None, gaps=(0, 0, 0, 0), tag=""):
    global tot, bad
    errs = np.zeros(batch.n, dtype=np.int32)
    exp, ecg, _ = ck.oracle_batch(kind, batch, mode, bw, mtx, gaps, errs=errs)
    t0 = time.time()
    if kind == "epi8":
        got = ctx.epi8_batch(batch, mode, bw, mtx, *gaps)
    else:
        got = ctx.edit_batch(batch, mode, bw)
    dt = time.time() - t0
    gcg = got.cigars()
    nb = 0; first = None; stbad = 0
    for i in range(batch.n):
        if errs[i]:
            if got.status[i] == 0: stbad += 1
            continue
        ok = np.array_equal(got.results[i], exp[i]) and np.array_equal(gcg[i], ecg[i]) and got.status[i] == 0
        if not ok:
            nb += 1
            if first is None: first = i
    tot += batch.n; bad += nb
    msg = "%-40s n=%d bad=%d oracle-flagged=%d (gpu unflagged %d) %.3fs" % (tag, batch.n, nb, int((errs != 0).sum()), stbad, dt)
    if first is not None:
        i = first
        msg += "\n   first bad pair %d qlen=%d tlen=%d status=%d\n   gpu %s\n   exp %s\n   ncig gpu %d exp %d" % (
            i, batch.qlen[i], batch.tlen[i], got.status[i], got.results[i], exp[i], len(gcg[i]), len(ecg[i]))
    print(msg, flush=True)

m = synth.score_matrix(2, -6)
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for gaps, name in [((-3, -2, 0, 0), "pw1"), ((0, -2, 0, 0), "pw0"), ((-3, -2, -8, -1), "pw2")]:
    for mode in (0, 1, 2):
        for bw in (0, 16, 64, 128):
            b = synth.make_pairs(64, 300, seed=mode * 100 + bw)
            cmp("epi8", b, mode, bw, m, gaps, "epi8 %s mode%d bw%d 300bp" % (name, mode, bw))
cmp("epi8", synth.make_pairs(1, 1000, seed=42), 0, 128, m, (-3, -2, 0, 0), "C1 single 1kb global bw128")
cmp("epi8", synth.make_pairs(300, 1000, seed=1000), 0, 0, m, (-3, -2, 0, 0), "C2-like 1kb global full")
cmp("epi8", synth.make_pairs(20, 10000, seed=2000, **dict(zip(("p_sub", "p_ins", "p_del"), synth.ont_like(0.12)))), 1, 512, m, (-3, -2, 0, 0), "C3-like 10kb overlap bw512")
for mode in (0, 1, 2):
    for bw in (0, 64, 128, 256):
        b = synth.make_pairs(200, 300, seed=7 + mode * 10 + bw, p_sub=0.02, p_ins=0.02, p_del=0.02)
        cmp("edit", b, mode, bw, tag="edit mode%d bw%d 300bp" % (mode, bw))
cmp("edit", synth.make_pairs(2000, 300, seed=3000, p_sub=0.02, p_ins=0.02, p_del=0.02), 0, 64, tag="C4-like edit 300bp bw64")
cmp("edit", synth.make_pairs(50, 1500, seed=3001), 0, 128, tag="edit 1.5kb bw128")
cmp("edit", synth.make_pairs(20, 1500, seed=3002), 1, 0, tag="edit 1.5kb overlap (W=24)")
print("TOTAL", tot, "BAD", bad)
print(ctx.timing())
