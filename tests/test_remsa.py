"""Re-alignment of reads against the MSA profile (remsa_pedit_rd_bspoacore, bspoa.h:3916-4045): the oracle's
restatement against records captured from the unmodified reference, and the CUDA path against both."""
import numpy as np
import pytest

import remsa_jobs as rj
from bsalign_b200 import synth


def test_oracle_remsa_core_matches_golden():
    """tests/golden/remsa_golden.npz (made by tests/golden/make_remsa_golden.py): 35 calls of the reference, bands of 16 / 32 / 64 cells;
    the rows of both DP matrices byte for byte and the MSA column every read position was merged into."""
    jobs = rj.load_golden()
    assert len(jobs) >= 30 and {j.bw for j in jobs} == {16, 32, 64}
    for k, j in enumerate(jobs):
        M0, M1, match, scr, err = rj.oracle_core(j)
        assert err == 0 and rj.compare(j, M0, M1, match) is None, (k, rj.compare(j, M0, M1, match))
        assert int((match >= 0).sum()) == j.nev


@pytest.mark.skipif(not rj.have_ref(), reason="oracle/_ref/libbsref_remsa.so not built (no /root/reference on this box)")
def test_oracle_remsa_core_vs_instrumented_reference():
    n = 0
    for seed, (nr, tl, realn, p, ebw) in enumerate([(8, 400, 2, .03, 0), (12, 900, 3, .05, 0), (5, 150, 1, .1, 0), (7, 500, 1, .04, 96), (4, 1800, 1, .04, 0)]):
        rng = np.random.default_rng(seed)
        tmpl = rng.integers(0, 4, (1, tl)).astype(np.uint8)
        reads = [synth.mutate_batch(rng, tmpl, p, p, p)[0] for _ in range(nr)]
        for j in rj.reference_dump(reads, realn=realn, editbw=ebw):
            M0, M1, match, scr, err = rj.oracle_core(j)
            assert err == 0 and rj.compare(j, M0, M1, match) is None, (seed, j.rid)
            n += 1
    assert n > 60


@pytest.mark.gpu
def test_gpu_remsa_matches_reference_records_and_oracle():
    """bsb200_remsa_batch (remsa_kernel: one warp per job) on the golden records as ONE batch: DP matrices, matched columns and the
    score of the walk equal the reference's records / the oracle."""
    from bsalign_b200 import api
    ctx = api.Context(0)
    try:
        jobs = rj.load_golden()
        ms, out, mm = api.remsa_batch(ctx, jobs, want_matrices=True)
        for k, j in enumerate(jobs):
            assert out[k, 1] == 0, (k, out[k])
            assert rj.compare(j, mm[k][0], mm[k][1], ms[k]) is None, (k, rj.compare(j, mm[k][0], mm[k][1], ms[k]))
            _, _, omatch, oscr, _ = rj.oracle_core(j)
            assert int(out[k, 0]) == oscr and int(out[k, 2]) == int((omatch >= 0).sum())
        # the same jobs many times over (a lock-step round of many objects), without the matrices
        big = jobs * 12
        ms2, out2, _ = api.remsa_batch(ctx, big)
        for k in range(len(big)):
            assert np.array_equal(ms2[k], ms[k % len(jobs)]) and np.array_equal(out2[k], out[k % len(jobs)])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_remsa_on_full_shape_synthetic_jobs():
    """MSAs of 20k columns (15 kb reads, BASELINE configs[4]'s shape) and other bands: synthetic inputs in the reference's layout
    (bsalign_b200/synth_remsa.py), CUDA against the oracle: matrices, matched columns, score."""
    from bsalign_b200 import api, synth_remsa
    ctx = api.Context(0)
    try:
        jobs = [synth_remsa.make_job(20000, bw=32, seed=1), synth_remsa.make_job(5000, bw=64, seed=2), synth_remsa.make_job(3000, bw=16, seed=3),
                synth_remsa.make_job(777, bw=48, seed=4), synth_remsa.make_job(9000, bw=32, seed=5, p_err=0.2)]
        ms, out, mm = api.remsa_batch(ctx, jobs, want_matrices=True)
        for k, j in enumerate(jobs):
            M0, M1, match, scr, err = rj.oracle_core(j)
            rl = j.bw + 2
            lo, hi = rl * 2 * j.mbeg, rl * 2 * j.mend
            assert err == 0 and out[k, 1] == 0 and int(out[k, 0]) == scr, (k, out[k], scr)
            assert np.array_equal(mm[k][0][lo:hi], M0[lo:hi]) and np.array_equal(mm[k][1][lo:hi], M1[lo:hi]), k
            assert np.array_equal(ms[k], match), k
    finally:
        ctx.close()
