"""Golden records of remsa_pedit_rd_bspoacore (bspoa.h:3916-4045) from the UNMODIFIED reference: every call the reference's own end_bspoa
makes on a few small BSPOA jobs, captured by oracle/_ref/libbsref_remsa.so (oracle/ref_remsa_harness.c: the reference compiled with
enter / exit hooks).  Run in the build container only:   python tests/golden/make_remsa_golden.py  ->  remsa_golden.npz
Per record: the inputs in the reference's own byte layout, the rows of the two DP matrices this call wrote, the MSA column every read
position ended up merged into."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import remsa_jobs as rj
from bsalign_b200 import synth


def main():
    assert rj.have_ref(), "oracle/_ref/libbsref_remsa.so missing: make -C oracle refremsa"
    keep = []
    # (reads, template length, realn rounds, error rate, editbw: the band is editbw / 2 cells)
    for seed, (nr, tl, realn, p, ebw) in enumerate([(6, 300, 2, .03, 0), (5, 700, 1, .06, 0), (4, 150, 1, .1, 0), (6, 400, 1, .03, 128), (5, 350, 1, .04, 32), (3, 2500, 1, .03, 0)]):
        rng = np.random.default_rng(100 + seed)
        tmpl = rng.integers(0, 4, (1, tl)).astype(np.uint8)
        reads = [synth.mutate_batch(rng, tmpl, p, p, p)[0] for _ in range(nr)]
        jobs = rj.reference_dump(reads, realn=realn, editbw=ebw)
        print("job", seed, "calls", len(jobs), "mlen", jobs[0].mlen, "bw", jobs[0].bw)
        keep += jobs
    rj.save_golden(keep)
    print("records", len(keep), os.path.getsize(rj.GOLD), "bytes")


if __name__ == "__main__":
    main()
