"""Repeat a few epi8 cases many times and count runs that differ from the oracle (flushes out races)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import checkers as ck
from bsalign_b200 import api, synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
small = len(sys.argv) > 2
ctx = api.Context(0)
M26 = synth.score_matrix(2, -6)
cases = [(2000, 6, 0, 0), (1000, 40, 1, 128), (3000, 8, 1, 512), (700, 64, 0, 0)] if not small else [(300, 12, 0, 0), (400, 12, 1, 64)]
for qlen, n, mode, bw in cases:
    b = synth.make_pairs(n, qlen, seed=qlen + mode)
    exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, M26, (-3, -2, 0, 0), nthreads=8)
    for lat in ("0", "1"):
        os.environ["BSB200_LAT"] = lat
        bad = 0
        for r in range(reps):
            got = ctx.epi8_batch(b, mode, bw, M26, -3, -2, 0, 0)
            gc = got.cigars()
            ok = all(got.status[i] == 0 and np.array_equal(got.results[i], exp[i]) and np.array_equal(gc[i], ecg[i]) for i in range(n))
            bad += not ok
        print("qlen %d n %d mode %d bw %d LAT=%s: %d/%d runs differ" % (qlen, n, mode, bw, lat, bad, reps), flush=True)
