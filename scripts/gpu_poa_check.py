"""GPU check of the POA sweep kernel against the reference's own sweep dumps (and the oracle)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poa_jobs as pj
from bsalign_b200 import api, poa

ctx = api.Context(0)
cases = [
    (8, 600, None, ()),
    (12, 1500, [128, 2, -6, -3, -2, 0, 0, 20, 1, 1], ()),
    (12, 1500, [128, 2, -6, -3, -2, -8, -1, 20, 1, 0], ()),
    (12, 1500, [64, 2, -6, -3, -2, -8, -1, 20, 0, 2], ()),
    (12, 1500, [256, 2, -6, 0, -2, 0, 0, 0, 1, 1], ()),
    (12, 1500, [80, 3, -4, -5, -3, -12, -1, 5, 2, 1], (0.05, 0.05, 0.05)),
    (25, 2000, None, (0.08, 0.08, 0.1)),
    (6, 100, None, ()),
    (30, 3000, None, ()),
]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
bad = 0
for nr, tl, po, err in cases:
    reads = pj.make_reads(nr, tl, 11, *err)
    jobs = pj.ref_dump(reads, po)
    batch = poa.SweepBatch([j.to_api() for j in jobs])
    t0 = time.time()
    res = poa.poa_rows_batch(ctx, batch)
    dt = time.time() - t0
    nb = 0
    for k, j in enumerate(jobs):
        rows, ub = res.linear(k)
        # visited set: every node the reference marked done must match; compare best too
        r = pj.compare_rows(j, rows, ub, j.done, res.best[k])
        if r or res.status[k]:
            nb += 1
            if nb <= 3:
                print("  job", k, "status", res.status[k], r, "bw", j.bw, "nnode", j.nnode)
    print("case", nr, tl, po, err, "jobs", len(jobs), "bad", nb, "ops", res.ops.sum(axis=0), "%.1f ms" % (dt * 1e3), ctx.timing()["forward_ms"])
    bad += nb
print("TOTAL BAD", bad)
sys.exit(1 if bad else 0)
