"""GPU parity tests of the POA sweep (bsb200_poa_* in include/bsalign_b200.h) through the C ABI:
every node row block (reference g->memp layout), the anchors and (maxscr, maxidx, maxoff) against
  - the committed sweep dumps of the unmodified reference (tests/golden/poa_golden.npz),
  - the oracle on larger seeded jobs (graphs from the compiled reference when oracle/_ref travelled to this box,
    else synthetic graphs from bsalign_b200.synth_poa)."""
import numpy as np
import pytest

import checkers as ck
import poa_jobs as pj
from bsalign_b200 import api, poa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def compare_with_oracle(job, rows, ub, best):
    orows, oub, odone, obest, _ops, rc = pj.oracle_sweep(job)
    assert rc == 0
    assert list(best) == list(obest), (list(best), list(obest))
    m = odone.astype(bool)
    assert np.array_equal(ub[m], oub[m])
    assert np.array_equal(rows[m], orows[m])


def test_poa_golden_dumps(ctx):
    jobs = pj.load_golden()
    res = poa.poa_rows_batch(ctx, poa.SweepBatch([j.to_api() for j in jobs]))
    for k, j in enumerate(jobs):
        rows, ub = res.linear(k)
        assert res.status[k] == 0
        assert pj.compare_rows(j, rows, ub, j.done, res.best[k]) is None, (k, j.bw, j.pw, pj.compare_rows(j, rows, ub, j.done, res.best[k]))


def test_poa_blocks_are_reference_memp_bytes(ctx):
    # the block arena must be byte-identical to g->memp from block 2 on: striped u/e/q then 17 ubegs ints
    jobs = [j for j in pj.load_golden() if j.nnode > 2][:4]
    res = poa.poa_rows_batch(ctx, poa.SweepBatch([j.to_api() for j in jobs]))
    for k, j in enumerate(jobs):
        blk = res.blocks(k)
        W = j.bw // 16
        p = np.arange(j.bw)
        idx = (p % W) * 16 + p // W
        m = j.done.astype(bool)
        for a in range(j.pw + 1):
            striped = np.zeros((j.nnode, j.bw), dtype=np.int8)
            striped[:, idx] = j.rows[:, a, :]
            assert np.array_equal(blk[m, a * j.bw:(a + 1) * j.bw].view(np.int8), striped[m])
        off = j.bw * (j.pw + 1)
        assert np.array_equal(np.ascontiguousarray(blk[m, off:off + 68]).view(np.int32).reshape(-1, 17), j.ub[m])


@pytest.mark.parametrize("nreads,tlen,par,err", [
    (24, 2500, None, (0.03, 0.03, 0.04)),
    (12, 1500, [128, 2, -6, -3, -2, 0, 0, 20, 1, 1], (0.05, 0.05, 0.05)),
    (12, 1500, [128, 2, -6, -3, -2, -8, -1, 20, 1, 0], (0.03, 0.03, 0.04)),
    (12, 1500, [64, 2, -6, -3, -2, -8, -1, 20, 0, 2], (0.04, 0.04, 0.04)),
    (12, 1200, [256, 2, -6, 0, -2, 0, 0, 0, 1, 1], (0.03, 0.03, 0.04)),
    (12, 1500, [80, 3, -4, -5, -3, -12, -1, 5, 2, 1], (0.05, 0.05, 0.05)),
    (20, 2000, None, (0.08, 0.08, 0.10)),
])
def test_poa_reference_graphs(ctx, nreads, tlen, par, err):
    if not ck.have_ref():
        pytest.skip("oracle/_ref/libbsref.so did not travel to this box")
    jobs = pj.ref_dump(pj.make_reads(nreads, tlen, 77, *err), par)
    res = poa.poa_rows_batch(ctx, poa.SweepBatch([j.to_api() for j in jobs]))
    for k, j in enumerate(jobs):
        rows, ub = res.linear(k)
        assert res.status[k] == 0
        assert pj.compare_rows(j, rows, ub, j.done, res.best[k]) is None, (k, pj.compare_rows(j, rows, ub, j.done, res.best[k]))
        compare_with_oracle(j, rows, ub, res.best[k])


def test_poa_staged_equals_one_shot_and_counts_ops(ctx):
    jobs = pj.load_golden()
    batch = poa.SweepBatch([j.to_api() for j in jobs])
    one = poa.poa_rows_batch(ctx, batch)
    rs = poa.ResidentSweeps(ctx, batch)
    try:
        rs.run(); rs.run()      # re-running resident jobs must be idempotent
        two = rs.fetch()
    finally:
        rs.free()
    assert np.array_equal(one.best, two.best) and np.array_equal(one.ops, two.ops)
    for k, j in enumerate(jobs):
        a, ua = one.linear(k); b, ub = two.linear(k)
        m = j.done.astype(bool)
        assert np.array_equal(a[m], b[m]) and np.array_equal(ua[m], ub[m])
        _r, _u, _d, _b, ops, _rc = pj.oracle_sweep(j)
        assert list(one.ops[k]) == list(ops)


def test_poa_empty_batch_and_trivial_graph(ctx):
    res = poa.poa_rows_batch(ctx, poa.SweepBatch([]))
    assert res.best.shape == (0, 3)
    # head -> tail only (first read of a BSPOA): the end candidate comes from the head's init row
    jobs = [j for j in pj.load_golden() if j.nnode == 2]
    assert jobs
    res = poa.poa_rows_batch(ctx, poa.SweepBatch([j.to_api() for j in jobs]))
    for k, j in enumerate(jobs):
        assert (int(res.best[k][0]), int(res.best[k][1]), int(res.best[k][2])) == (j.maxscr, j.maxidx, j.maxoff)


def test_poa_dropin_end_bspoa_identical_msa():
    """include/bsalign_b200_poa_compat.h against the reference's own end_bspoa: whole BSPOA jobs (graph surgery, consensus and re-alignment
    rounds are the reference's code; every read-vs-graph sweep runs on the GPU, all objects in lock-step) must give byte-identical
    consensus + MSA.  The binary is built from the reference headers where they exist (oracle/Makefile, target dropin) and travels."""
    import os
    import subprocess
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "poa_dropin")):
        pytest.skip("oracle/_ref/poa_dropin was not built (no reference tree at build time)")
    # poa_dropin: sweep + walk on the device, graph surgery replayed on the host; poa_dropin_hosttb: row blocks back, the reference's own traceback
    for name in ("poa_dropin", "poa_dropin_hosttb"):
        # (the last one is ONE job at BASELINE config 5's full shape: 64 reads x 15 kb, DEFAULT_BSPOA_PAR)
        for args in (["6", "12", "1500", "3"], ["3", "24", "4000", "5", "0"]) + ((["1", "64", "15000", "7"],) if name == "poa_dropin" else ()):
            out = subprocess.run([os.path.join(ref_dir, name)] + args, capture_output=True, text=True, timeout=600)
            assert out.returncode == 0, out.stdout + out.stderr
            assert "identical=%s/%s" % (args[0], args[0]) in out.stdout, out.stdout
            # the band placement of every round (kmer_striped_seqedit_pairwise inside prepare_rd_align_bspoa, bspoa.h:2089) went through
            # bsb200_kmer_edit_batch: include/bsalign_b200_poa_kmer.h
            assert "kmer_pairs=" in out.stdout and "kmer_pairs=0" not in out.stdout, out.stdout


def test_poa_dropin_realign_rounds_on_gpu_identical_msa():
    """The re-alignment rounds (remsa_pedits_bspoa, bspoa.h:4178-4457) with the DP + walk of every read (remsa_pedit_rd_bspoacore,
    bspoa.h:3916-4045) on the GPU: include/bsalign_b200_poa_remsa.h + b200_poa_realign_run, one bsb200_remsa_batch per rendezvous of the
    in-flight objects.  Whole jobs through b200_end_bspoa_batch against the reference's end_bspoa: byte-identical consensus + MSA,
    including objects with different read counts (they leave the rendezvous early) and one full-shape job (64 reads x 15 kb)."""
    import os
    import re
    import subprocess
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    exe = os.path.join(ref_dir, "poa_remsa_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/poa_remsa_dropin was not built (no reference tree at build time)")
    for args, reads_min in ((["6", "12", "1500", "3"], 12), (["5", "10", "3000", "5", "2", "6"], 10), (["1", "64", "15000", "7"], 64)):
        out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "identical=%s/%s" % (args[0], args[0]) in out.stdout, out.stdout
        realn = int(args[4]) if len(args) > 4 else 3
        m = re.search(r"remsa_batches=(\d+) remsa_jobs=(\d+)", out.stdout)
        # every re-aligned read of every object and round went through bsb200_remsa_batch (beyond par.seqcore = 40 reads only the last
        # round takes all reads, bspoa.h:4188 / :4346)
        core = min(reads_min, 40)
        assert m and int(m.group(2)) >= int(args[0]) * core * realn and int(m.group(1)) >= core * realn, out.stdout


def _as_dump_like(job):
    """A bsalign_b200.poa.SweepJob viewed through the attribute names the checker helpers use."""
    w = pj.SweepJob()
    w.bw = int(job.par[0]); w.alnmode, w.M, w.X, w.O, w.E, w.Q, w.P, w.T, w.refbonus = [int(x) for x in job.par[1:]]
    w.slen = len(job.query); w.nnode = job.nnode; w.head = job.head; w.tail = job.tail; w.nedge = len(job.edst)
    w.query = job.query; w.base = job.base; w.bonus = job.bonus; w.rpos = job.rpos; w.nct = job.nct; w.eoff = job.eoff; w.edst = job.edst
    w.reoff = job.reoff; w.resrc = job.resrc; w.recov = job.recov
    return w


def test_poa_device_walk_matches_reference_and_oracle(ctx):
    """alignment2graph_bspoa's walk on the device: counts, end state and the node every read position was merged into equal what the
    unmodified reference did (golden dumps), and the per-position matches equal the oracle's restatement of the walk."""
    jobs = pj.load_golden()
    res = poa.poa_align_batch(ctx, poa.SweepBatch([j.to_api() for j in jobs]), want_rows=True)
    for k, j in enumerate(jobs):
        assert pj.compare_backtrace(j, res.match(k), res.trace[k]) is None, (k, pj.compare_backtrace(j, res.match(k), res.trace[k]))
        rows, ub = res.linear(k)
        om, oo = pj.oracle_backtrace(j, rows, ub, int(res.best[k][1]), int(res.best[k][2]))
        assert np.array_equal(om, res.match(k)) and np.array_equal(oo, res.trace[k]), (k, oo.tolist(), res.trace[k].tolist())


def test_poa_device_walk_on_larger_graphs(ctx):
    from bsalign_b200 import synth_poa
    if ck.have_ref():
        jobs = pj.ref_dump(pj.make_reads(20, 2500, 91, 0.05, 0.05, 0.06), None)
        api_jobs = [j.to_api() for j in jobs]
    else:
        jobs = None
        api_jobs = [synth_poa.make_sweep_job(300 + i, tlen=2500) for i in range(8)]
    res = poa.poa_align_batch(ctx, poa.SweepBatch(api_jobs), want_rows=True)
    for k, aj in enumerate(api_jobs):
        w = jobs[k] if jobs is not None else _as_dump_like(aj)
        if jobs is not None:
            assert pj.compare_backtrace(w, res.match(k), res.trace[k]) is None, (k, pj.compare_backtrace(w, res.match(k), res.trace[k]))
        rows, ub = res.linear(k)
        om, oo = pj.oracle_backtrace(w, rows, ub, int(res.best[k][1]), int(res.best[k][2]))
        assert np.array_equal(om, res.match(k)) and np.array_equal(oo, res.trace[k]), (k, oo.tolist(), res.trace[k].tolist())
    # synthetic graphs (what bench.py --workload c5 runs) through the same check
    sj = [synth_poa.make_sweep_job(500 + i, tlen=3000) for i in range(4)]
    res = poa.poa_align_batch(ctx, poa.SweepBatch(sj), want_rows=True)
    for k, aj in enumerate(sj):
        rows, ub = res.linear(k)
        om, oo = pj.oracle_backtrace(_as_dump_like(aj), rows, ub, int(res.best[k][1]), int(res.best[k][2]))
        assert int(oo[7]) == 0 and np.array_equal(om, res.match(k)) and np.array_equal(oo, res.trace[k])
        assert int((om >= 0).sum()) > 0.8 * len(aj.query)
