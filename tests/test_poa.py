"""CPU tests of the POA sweep checker and host logic:
  (1) oracle/bsalign_oracle.c:bso_poa_sweep against the committed sweep dumps of the unmodified reference
      (tests/golden/poa_golden.npz, made by tests/golden/make_poa_golden.py): every node row, the anchors, best end;
  (2) the same against live dumps of the compiled reference (oracle/_ref), when present;
  (3) packing of sweep jobs into the C-ABI arenas (bsalign_b200/poa.py), block size rule g->mmblk (bspoa.h:2217).
"""
import numpy as np
import pytest

import checkers as ck
import poa_jobs as pj
from bsalign_b200 import api, poa


def check_job(j):
    rows, ub, done, best, ops, rc = pj.oracle_sweep(j)
    assert rc == 0
    assert pj.compare_rows(j, rows, ub, done, best) is None, (j.bw, j.pw, j.nnode, pj.compare_rows(j, rows, ub, done, best))
    return ops


def test_oracle_sweep_matches_golden_dumps():
    jobs = pj.load_golden()
    assert len(jobs) >= 30
    assert {j.pw for j in jobs} == {0, 1, 2} and {j.alnmode for j in jobs} == {0, 1, 2}
    nupd = 0
    for j in jobs:
        nupd += int(check_job(j)[0])
    assert nupd > 3000


@pytest.mark.ref
@pytest.mark.skipif(not ck.have_ref(), reason="oracle/_ref/libbsref.so not built")
@pytest.mark.parametrize("nreads,tlen,par,err", [
    (10, 800, None, (0.03, 0.03, 0.04)),
    (8, 500, [128, 2, -6, -3, -2, 0, 0, 20, 1, 1], (0.05, 0.05, 0.05)),
    (8, 500, [96, 2, -6, -3, -2, -8, -1, 20, 1, 0], (0.04, 0.04, 0.04)),
    (8, 500, [64, 2, -6, -3, -2, -8, -1, 0, 0, 2], (0.08, 0.08, 0.10)),
    (16, 1200, None, (0.08, 0.08, 0.10)),
])
def test_oracle_sweep_matches_live_reference(nreads, tlen, par, err):
    jobs = pj.ref_dump(pj.make_reads(nreads, tlen, 31, *err), par)
    assert len(jobs) == nreads
    for j in jobs:
        check_job(j)


def test_block_bytes_rule():
    # g->mmblk = roundup16(bw * (piecewise + 1) + 17 * 4), bspoa.h:2217; DEFAULT_BSPOA_PAR at band 128 is two-piece -> 464
    d = poa.DEFAULT_BSPOA_PAR
    par = [d[k] for k in poa.PAR_FIELDS]
    assert poa.block_bytes(par) == 464
    assert poa.block_bytes([128, 1, 2, -6, -3, -2, 0, 0, 20, 1]) == 336      # affine
    assert poa.block_bytes([64, 1, 2, -6, 0, -2, 0, 0, 20, 1]) == 144        # linear
    L = api.lib()
    poa._bind(L)
    for p in (par, [128, 1, 2, -6, -3, -2, 0, 0, 20, 1], [64, 1, 2, -6, 0, -2, 0, 0, 20, 1], [48, 2, 3, -4, -5, -3, -12, -1, 5, 2]):
        a = np.array(p, dtype=np.int32)
        assert L.bsb200_poa_block_bytes(api._ptr(a)) == poa.block_bytes(p)


def test_sweep_batch_packing():
    jobs = pj.load_golden()[:7]
    b = poa.SweepBatch([j.to_api() for j in jobs])
    assert b.n == 7 and b.par.shape == (7, 10)
    for i, j in enumerate(jobs):
        n0, n1 = int(b.node_off[i]), int(b.node_off[i + 1])
        assert n1 - n0 == j.nnode
        assert np.array_equal(b.rpos[n0:n1], j.rpos) and np.array_equal(b.base[n0:n1], j.base)
        e = b.eoff[n0 + i:n1 + i + 1]
        assert e[0] == 0 and e[-1] == j.nedge
        assert np.array_equal(b.edst[int(b.edge_off[i]):int(b.edge_off[i + 1])], j.edst)
        assert np.array_equal(b.queries[int(b.qoff[i]):int(b.qoff[i]) + int(b.slen[i])], j.query)
        assert int(b.row_off[i + 1] - b.row_off[i]) == j.nnode * poa.block_bytes(j.params())


def test_poa_entry_points_fail_without_gpu():
    L = api.lib()
    if L.bsb200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        poa.poa_rows_batch(api.Context(0), poa.SweepBatch([j.to_api() for j in pj.load_golden()[:1]]))


def test_oracle_walk_matches_golden_dumps():
    """bso_poa_backtrace (the walk of alignment2graph_bspoa) on the reference's own rows: mat/mis/ins/del, the end of the walk and
    the node every read position was merged into equal what the unmodified reference did."""
    n = 0
    for j in pj.load_golden():
        match, out = pj.oracle_backtrace(j, j.rows, j.ub, j.maxidx, j.maxoff)
        assert pj.compare_backtrace(j, match, out) is None, (j.bw, j.pw, j.nnode, pj.compare_backtrace(j, match, out))
        n += int((match >= 0).sum())
    assert n > 3000


@pytest.mark.ref
@pytest.mark.skipif(not ck.have_ref(), reason="oracle/_ref/libbsref.so not built")
def test_oracle_walk_matches_live_reference():
    for par, err in ((None, (0.05, 0.05, 0.06)), ([128, 2, -6, -3, -2, 0, 0, 20, 1, 0], (0.03, 0.03, 0.04)), ([64, 2, -6, -3, -2, -8, -1, 20, 0, 2], (0.04, 0.04, 0.04))):
        for j in pj.ref_dump(pj.make_reads(12, 1200, 17, *err), par):
            match, out = pj.oracle_backtrace(j, j.rows, j.ub, j.maxidx, j.maxoff)
            assert pj.compare_backtrace(j, match, out) is None


def test_realign_driver_rendezvous_with_cpu_checker():
    """The re-alignment driver of include/bsalign_b200_poa_compat.h (b200_poa_realign_run: one host thread per in-flight object, the
    rendezvous, the hook of bsalign_b200_poa_remsa.h, the replay of merge_nodes_bspoa from the matched columns) with the batch entry point
    replaced by the oracle's scalar core (oracle/poa_remsa_dropin_test.c, -DREMSA_CPU_CHECK: no GPU needed) against the reference's own
    end_bspoa: byte-identical consensus + MSA, also when objects hold different numbers of reads and leave the rendezvous early.  The GPU
    arm of the same program is tests/test_gpu_poa.py::test_poa_dropin_realign_rounds_on_gpu_identical_msa."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "poa_remsa_dropin_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/poa_remsa_dropin_cpu was not built (no reference tree at build time)")
    for args in (["4", "8", "1200", "3"], ["5", "6", "2000", "5", "2", "5"]):
        out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "identical=%s/%s" % (args[0], args[0]) in out.stdout and "remsa_jobs=0" not in out.stdout, out.stdout
    # below b200_poa_remsa_min_objects the rounds keep the reference's core on the worker threads: same result, no batch
    out = subprocess.run([exe, "3", "6", "1000", "9", "2", "0", "100"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "identical=3/3" in out.stdout and "remsa_jobs=0 " in out.stdout, out.stdout + out.stderr
