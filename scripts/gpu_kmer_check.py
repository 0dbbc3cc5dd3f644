"""GPU check of the k-mer guided edit (bsb200_kmer_edit_batch) against the oracle: mixed lengths and k, unrelated pairs (fallback to
the plain global edit), clipped ends, long pairs, a starved gap-trace pool (retry rounds).  Development aid; the tests hold a subset."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
from bsalign_b200 import api, synth  # noqa: E402


def mk(rng, n, qlen, ps, pi, pd):
    q = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
    t, tl = synth.mutate_batch(rng, q, ps, pi, pd)
    pairs, o = [], 0
    for i in range(n):
        pairs.append((q[i].copy(), t[o:o + tl[i]].copy()))
        o += tl[i]
    return pairs


def check(ctx, batch, ksz, tag, dense=False):
    got = ctx.kmer_edit_batch(batch, ksz, dense=dense)
    errs = np.zeros(batch.n, np.int32)
    exp, ecg, _ = ck.kmer_batch("oracle", batch, ksz, nthreads=8, errs=errs)
    gc = got.cigars()
    bad = 0
    for i in range(batch.n):
        ok = np.array_equal(got.results[i], exp[i]) and np.array_equal(gc[i], ecg[i]) and (got.status[i] & 7) == (errs[i] & 7)
        if not ok:
            bad += 1
            if bad < 4:
                print("  BAD", tag, i, batch.qlen[i], batch.tlen[i], got.results[i], exp[i], len(gc[i]), len(ecg[i]), got.status[i], errs[i])
    print("%-40s ksz=%2d n=%6d bad=%d fallback=%d" % (tag, ksz, batch.n, bad, ctx.timing()["waves"]))
    return bad


def main():
    ctx = api.Context(0)
    rng = np.random.default_rng(11)
    bad = 0
    for ksz in (13, 15, 8, 5, 3, 1):
        for (n, qlen, ps, pi, pd) in ((300, 300, .03, .03, .04), (100, 1000, .05, .05, .05), (20, 3000, .02, .02, .02), (100, 100, .1, .1, .1),
                                      (100, 60, .02, .02, .02), (50, 500, .2, .1, .1), (60, 20, .05, 0, 0), (40, 200, 0, 0, 0)):
            pairs = mk(rng, n, qlen, ps, pi, pd)
            pairs += [(rng.integers(0, 4, int(rng.integers(1, 80))).astype(np.uint8), rng.integers(0, 4, int(rng.integers(1, 80))).astype(np.uint8)) for _ in range(9)]
            for i in range(0, len(pairs), 3):
                a, b = pairs[i]
                if len(b) > 30:
                    pairs[i] = (a, b[int(rng.integers(0, 15)):len(b) - int(rng.integers(0, 15))])
            pairs.insert(5, (np.zeros(0, np.uint8), np.array([1, 2, 3], np.uint8)))
            batch = synth.PairBatch.from_lists(pairs)
            bad += check(ctx, batch, ksz, "len %d sub %.2f" % (qlen, ps), dense=(ksz == 15))
    # long reads (the POA use: a 15 kb read against a consensus) and big gaps (anchors only at the ends)
    pairs = mk(rng, 12, 15000, .04, .03, .03)
    core = rng.integers(0, 4, 1500).astype(np.uint8)
    a = np.concatenate([core[:200], rng.integers(0, 4, 900).astype(np.uint8), core[1200:]])
    pairs += [(core, a), (a, core)]
    batch = synth.PairBatch.from_lists(pairs)
    bad += check(ctx, batch, 13, "15 kb reads + 1 kb gaps")
    os.environ["BSB200_KMER_POOL"] = "600000"
    bad += check(ctx, batch, 13, "same, starved pool (retry rounds)")
    del os.environ["BSB200_KMER_POOL"]
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        for (n, qlen) in ((200000, 300), (1000000, 300), (100000, 1000), (2000, 15000)):
            batch = synth.make_pairs(n, qlen, seed=5, p_sub=0.02, p_ins=0.02, p_del=0.02)
            for _ in range(3):
                t0 = time.perf_counter()
                got = ctx.kmer_edit_batch(batch, 13, dense=True)
                dt = time.perf_counter() - t0
            tm = ctx.timing()
            print("n=%d qlen=%d: call %.1f ms, kernel %.2f ms (%d launches), fallback %.2f ms for %d pairs, h2d %.2f d2h %.2f ms" % (
                n, qlen, dt * 1e3, tm["forward_ms"], tm["forward_launches"], tm["traceback_ms"], tm["waves"], tm["h2d_ms"], tm["d2h_ms"]))
            m = min(n, 20000)
            sub = synth.PairBatch(batch.seqs, batch.qoff[:m], batch.qlen[:m], batch.toff[:m], batch.tlen[:m])
            t0 = time.perf_counter()
            r, c, _ = ck.kmer_batch("ref", sub, 13, nthreads=os.cpu_count())
            dtc = time.perf_counter() - t0
            gc = got.cigars()
            eq = np.array_equal(r, got.results[:m]) and all(np.array_equal(c[i], gc[i]) for i in range(m))
            print("   reference on %d threads: %.1f us per pair -> %.1f ms for the batch; GPU call is %.1fx; first %d pairs equal: %s" % (
                os.cpu_count(), dtc / m * 1e6, dtc / m * n * 1e3, (dtc / m * n) / dt, m, eq))
    print("TOTAL BAD", bad)
    ctx.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
