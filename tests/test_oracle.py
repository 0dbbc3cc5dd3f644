"""CPU tests of the oracle (our restatement of the reference algorithm) against
  (1) the committed golden vectors generated from the unmodified reference (tests/golden/make_golden.py),
  (2) the README example of the reference (README.md:36-42),
  (3) the reference itself compiled here (oracle/_ref), when present, on seeded random + adversarial pairs.
"""
import ctypes
import os

import numpy as np
import pytest

import checkers as ck
from bsalign_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name))
    batch = synth.PairBatch(z["seqs"], z["qoff"], z["qlen"], z["toff"], z["tlen"])
    return z, batch


def golden_cases(z, kind):
    for ci, cfg in enumerate(z["configs"]):
        ncig = z["ncig%d" % ci]
        off = np.concatenate([[0], np.cumsum(ncig.astype(np.int64))]).astype(np.int64)
        arr = z["cig%d" % ci]
        cigs = [arr[int(off[i]):int(off[i + 1])] for i in range(len(ncig))]
        yield ci, [int(x) for x in cfg], z["valid%d" % ci], z["res%d" % ci], cigs


def check_against(run, z, batch, kind):
    total = 0
    for ci, cfg, valid, res, cigs in golden_cases(z, kind):
        got_res, got_cigs = run(cfg)
        for i in np.nonzero(valid)[0]:
            assert np.array_equal(got_res[i], res[i]), (kind, cfg, int(i), got_res[i], res[i])
            assert np.array_equal(got_cigs[i], cigs[i]), (kind, cfg, int(i))
            total += 1
    return total


def test_oracle_epi8_matches_golden():
    z, batch = load_golden("epi8_golden.npz")

    def run(cfg):
        mode, bw, M, X, go1, ge1, go2, ge2 = cfg
        r, c, _ = ck.oracle_batch("epi8", batch, mode, bw, synth.score_matrix(M, X), (go1, ge1, go2, ge2))
        return r, c
    assert check_against(run, z, batch, "epi8") > 800


def test_oracle_edit_matches_golden():
    z, batch = load_golden("edit_golden.npz")

    def run(cfg):
        r, c, _ = ck.oracle_batch("edit", batch, cfg[0], cfg[1])
        return r, c
    assert check_against(run, z, batch, "edit") > 400


def kmer_golden():
    z, batch = load_golden("kmer_golden.npz")
    cases = []
    for ci, k in enumerate(z["ks"]):
        ncig = z["ncig%d" % ci]
        off = np.concatenate([[0], np.cumsum(ncig.astype(np.int64))]).astype(np.int64)
        cases.append((int(k), z["res%d" % ci], [z["cig%d" % ci][int(off[i]):int(off[i + 1])] for i in range(len(ncig))]))
    return batch, cases


def test_oracle_kmer_edit_matches_golden():
    """kmer_striped_seqedit_pairwise (bsalign.h:1209): the oracle's restatement against the answers of the unmodified reference
    (tests/golden/make_kmer_golden.py): 96 pairs x 4 k-mer sizes, anchored pairs, pairs without anchors, clipped ends, big gaps."""
    batch, cases = kmer_golden()
    for k, res, cigs in cases:
        r, c, _ = ck.kmer_batch("oracle", batch, k, nthreads=4)
        assert np.array_equal(r, res), k
        assert all(np.array_equal(a, b) for a, b in zip(c, cigs)), k


@pytest.mark.skipif(not ck.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_kmer_edit_vs_compiled_reference():
    rng = np.random.default_rng(77)
    for k in (13, 6, 1, 15):
        pairs = []
        for n, qlen, ps in ((30, 250, .04), (10, 900, .06), (20, 70, .1), (10, 400, .25)):
            q = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
            t, tl = synth.mutate_batch(rng, q, ps, ps, ps)
            o = 0
            for i in range(n):
                pairs.append((q[i].copy(), t[o + (i % 3) * 4:o + tl[i]].copy()))
                o += tl[i]
        pairs += [(rng.integers(0, 4, 50).astype(np.uint8), rng.integers(0, 4, 61).astype(np.uint8)) for _ in range(4)]
        batch = synth.PairBatch.from_lists(pairs)
        r1, c1, _ = ck.kmer_batch("ref", batch, k, nthreads=4)
        r2, c2, _ = ck.kmer_batch("oracle", batch, k, nthreads=4)
        assert np.array_equal(r1, r2), k
        assert all(np.array_equal(a, b) for a, b in zip(c1, c2)), k


@pytest.mark.skipif(not ck.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_kmer_edit_on_repetitive_sequences():
    """homopolymers, tandem repeats, duplications, reverse complements: the multiplicity rules of the anchoring (a k-mer counts only when
    it occurs once in each sequence on the same strand) and the sentinel quirk of the reference's scan."""
    for seed, k in ((1, 13), (2, 3), (3, 15), (4, 2), (5, 7), (6, 1)):
        batch = ck.repetitive_pairs(seed)
        r1, c1, _ = ck.kmer_batch("ref", batch, k)
        r2, c2, _ = ck.kmer_batch("oracle", batch, k)
        assert np.array_equal(r1, r2), (seed, k)
        assert all(np.array_equal(a, b) for a, b in zip(c1, c2)), (seed, k)


def test_readme_example():
    """README.md:36-42 of the reference: pair 29.1/29.2, score 128, 71 matches, 4 mismatches, one indel."""
    z = np.load(os.path.join(GOLD, "readme_pair.npz"))
    b = synth.PairBatch.from_lists([(z["q"], z["t"])])
    r, c, err = ck.oracle_batch("epi8", b, 1, int(z["bandwidth"]), synth.score_matrix(2, -2), (-4, -2, 0, 0))
    assert err == 0
    assert r[0][0] == 128 and r[0][5] == 71 and r[0][6] == 4 and r[0][7] + r[0][8] == 1
    assert np.array_equal(r[0], z["res"]) and np.array_equal(c[0], z["cig"])


def test_empty_inputs_give_zero_result():
    b = synth.PairBatch.from_lists([(np.zeros(0, np.uint8), np.array([1, 2], np.uint8)), (np.array([1], np.uint8), np.zeros(0, np.uint8))])
    for kind in ("epi8", "edit"):
        r, c, _ = ck.oracle_batch(kind, b, 0, 0, synth.score_matrix(2, -6), (-3, -2, 0, 0))
        assert not r.any() and all(len(x) == 0 for x in c)


def _random_pairs(rng, n, lo, hi, err):
    pairs = []
    for _ in range(n):
        ql = int(rng.integers(lo, hi + 1))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        kind = rng.integers(0, 3)
        if kind == 0:
            t, _ = synth.mutate_batch(rng, q[None, :], err / 3, err / 3, err / 3)
        elif kind == 1:
            cut, ln = int(rng.integers(0, ql)), int(rng.integers(0, max(1, ql // 4)))
            t = np.concatenate([q[:cut], q[cut + ln:]])
        else:
            q = np.repeat(rng.integers(0, 4, (ql + 3) // 4).astype(np.uint8), 4)[:ql]
            t, _ = synth.mutate_batch(rng, q[None, :], err / 3, err / 3, err / 3)
        if len(t) == 0:
            t = np.array([0], np.uint8)
        pairs.append((q, t))
    return synth.PairBatch.from_lists(pairs)


@pytest.mark.ref
@pytest.mark.skipif(not ck.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_vs_compiled_reference(seed):
    """Differential check on the parity domain: pairs the oracle flags (reference UB) are skipped."""
    rng = np.random.default_rng(seed)
    params = [((2, -6), (-3, -2, 0, 0)), ((2, -6), (0, -2, 0, 0)), ((2, -6), (-3, -2, -8, -1)), ((2, -2), (-4, -2, 0, 0)), ((30, -40), (-40, -20, 0, 0))]
    checked = 0
    for it in range(6):
        b = _random_pairs(rng, 24, *[(1, 17), (20, 120), (100, 500)][it % 3], 0.15)
        (M, X), gaps = params[(it + seed) % len(params)]
        mtx = synth.score_matrix(M, X)
        for mode in (0, 1, 2):
            bw = int(rng.choice([0, 16, 48, 128]))
            errs = np.zeros(b.n, np.int32)
            r2, c2, _ = ck.oracle_batch("epi8", b, mode, bw, mtx, gaps, errs=errs)
            keep = np.nonzero(errs == 0)[0]
            out = ck.forked(ck.ref_batch, "epi8", b.subset(keep), mode, bw, mtx, gaps, timeout_s=60)
            assert out is not None, "reference crashed on pairs the oracle considers in-domain"
            for k, i in enumerate(keep):
                assert np.array_equal(out[0][k], r2[i]) and np.array_equal(out[1][k], c2[i])
                checked += 1
            bwe = int(rng.choice([0, 64, 128]))
            errs[:] = 0
            r2, c2, _ = ck.oracle_batch("edit", b, mode, bwe, errs=errs)
            keep = np.nonzero(errs == 0)[0]
            out = ck.forked(ck.ref_batch, "edit", b.subset(keep), mode, bwe, timeout_s=60)
            assert out is not None
            for k, i in enumerate(keep):
                assert np.array_equal(out[0][k], r2[i]) and np.array_equal(out[1][k], c2[i])
                checked += 1
    assert checked > 500


@pytest.mark.ref
@pytest.mark.skipif(not ck.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_rows_equal_reference_rows():
    """Forward pass pinned row by row: band offsets, the 17 anchors and the u/e/q bytes of every target row."""
    m = synth.score_matrix(2, -6)
    for seed, mode, bw, gaps in [(1, 0, 32, (-3, -2, 0, 0)), (2, 1, 64, (-3, -2, -8, -1)), (3, 2, 0, (0, -2, 0, 0)), (4, 0, 0, (-3, -2, 0, 0))]:
        b = synth.make_pairs(1, 240, seed)
        o = ck.rows_dump(ck.oracle(), "bso_epi8_pairwise_ex", b.query(0), b.target(0), mode, bw, m, gaps, extra_args=(None, ctypes.c_uint32(0), None))
        r = ck.rows_dump(ck.ref(), "bsref_epi8_rows", b.query(0), b.target(0), mode, bw, m, gaps)
        pw = 2 if gaps[2] else (1 if gaps[0] else 0)
        for k in range(4):
            assert np.array_equal(o[k], r[k])
        if pw >= 1:
            assert np.array_equal(o[4], r[4])
        if pw == 2:
            assert np.array_equal(o[5], r[5])


def test_oracle_matches_reference_on_real_ont_example():
    """Every 6th of the 12,477 ONT pairs of the reference's example/real.ont.b10M.txt under the three configurations of example/run.sh
    (answers of the unmodified reference: tests/golden/make_real_golden.py).  The GPU suite runs all of them."""
    import real_ont
    batch, cfgs, mtx, gaps = real_ont.load()
    assert batch.n == 12477
    idx = np.arange(0, batch.n, 6)
    sub = batch.subset(idx)
    for cfg in cfgs:
        r, c, _ = ck.oracle_batch(cfg["kind"], sub, cfg["mode"], cfg["bandwidth"], mtx, gaps, nthreads=8)
        ncg = np.array([len(x) for x in c], np.uint32)
        words = np.concatenate(c) if ncg.sum() else np.zeros(0, np.uint32)
        assert real_ont.compare(cfg, idx, r, ncg, words) == [], cfg["kind"]
