"""Generate tests/golden/poa_golden.npz from the UNMODIFIED reference (oracle/_ref/libbsref.so, bsref_poa_dump in
oracle/ref_harness.c): whole BSPOA jobs (beg/push/end with realn = 0) on small synthetic read sets, every
align_rd_bspoacore call dumped with its inputs (read, selected sub-graph, band offsets, parameters) and outputs (all
node rows, maxscr/maxidx/maxoff).  Run in the build container only:

    python tests/golden/make_poa_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck
import poa_jobs as pj

assert ck.have_ref(), "oracle/_ref/libbsref.so missing: run `make -C oracle` where /root/reference exists"

# (reads, template length, seed, par override {bandwidth, M, X, O, E, Q, P, T, refbonus, alnmode} or None = DEFAULT_BSPOA_PAR, error rates)
CASES = [
    (6, 260, 101, None, (0.03, 0.03, 0.04)),                                   # DEFAULT_BSPOA_PAR: two-piece, overlap, band 128
    (5, 200, 102, [64, 2, -6, -3, -2, 0, 0, 20, 1, 1], (0.05, 0.05, 0.05)),    # affine (one piece), band 64
    (5, 200, 103, [128, 2, -6, -3, -2, -8, -1, 20, 1, 0], (0.03, 0.03, 0.04)),  # global
    (5, 200, 104, [48, 3, -4, -5, -3, -12, -1, 5, 2, 2], (0.06, 0.06, 0.06)),   # extend, band 48, other scores
    (4, 150, 105, [256, 2, -6, 0, -2, 0, 0, 0, 0, 1], (0.03, 0.03, 0.04)),      # linear gaps, band wider than the reads
    (5, 90, 106, None, (0.08, 0.08, 0.08)),                                     # reads shorter than the band
]

out = {"ncase": np.array(len(CASES))}
total = 0
for ci, (nr, tl, seed, po, err) in enumerate(CASES):
    reads = pj.make_reads(nr, tl, seed, *err)
    blob = pj.ref_dump_blob(reads, po)
    jobs = pj.parse_dump(blob)
    total += len(jobs)
    out["blob%d" % ci] = blob
    print("case", ci, "jobs", len(jobs), "bytes", len(blob), "bw", sorted(set(j.bw for j in jobs)), "pw", sorted(set(j.pw for j in jobs)))
np.savez_compressed(os.path.join(HERE, "poa_golden.npz"), **out)
print("sweep jobs:", total, "file bytes:", os.path.getsize(os.path.join(HERE, "poa_golden.npz")))
