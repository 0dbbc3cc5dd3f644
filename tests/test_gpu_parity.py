"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (bsalign_b200.api -> libbsalign_b200.so),
against the oracle on the same seeded inputs, against the committed golden vectors, and - at BASELINE.json's
batch shapes - through size-independent properties.  Bar: bit-exact (score, coordinates, counts, CIGAR words)."""
import os

import numpy as np
import pytest

import checkers as ck
from bsalign_b200 import api, synth
from test_oracle import load_golden, golden_cases, _random_pairs

pytestmark = pytest.mark.gpu

M26 = synth.score_matrix(2, -6)


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def assert_same(got, exp_res, exp_cigs, errs=None, tag=""):
    gc = got.cigars()
    for i in range(len(exp_res)):
        if errs is not None and errs[i]:
            assert got.status[i] != 0, (tag, i, "oracle flags reference-UB but the GPU status is clean")
            continue
        assert got.status[i] == 0, (tag, i, int(got.status[i]))
        assert np.array_equal(got.results[i], exp_res[i]), (tag, i, got.results[i], exp_res[i])
        assert np.array_equal(gc[i], exp_cigs[i]), (tag, i)


@pytest.mark.parametrize("gaps", [(-3, -2, 0, 0), (0, -2, 0, 0), (-3, -2, -8, -1)], ids=["affine", "linear", "twopiece"])
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["global", "overlap", "extend"])
def test_epi8_matches_oracle(ctx, mode, gaps):
    for bw in (0, 16, 64, 128):
        b = synth.make_pairs(48, 300, seed=mode * 100 + bw)
        exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, M26, gaps)
        assert_same(ctx.epi8_batch(b, mode, bw, M26, *gaps), exp, ecg, tag=(mode, bw, gaps))


@pytest.mark.parametrize("mode", [0, 1, 2], ids=["global", "overlap", "extend"])
def test_edit_matches_oracle(ctx, mode):
    for bw in (0, 64, 128, 256):
        b = synth.make_pairs(128, 300, seed=7 + mode * 10 + bw, p_sub=0.02, p_ins=0.02, p_del=0.02)
        exp, ecg, _ = ck.oracle_batch("edit", b, mode, bw)
        assert_same(ctx.edit_batch(b, mode, bw), exp, ecg, tag=(mode, bw))


def test_golden_vectors(ctx):
    z, batch = load_golden("epi8_golden.npz")
    n = 0
    for ci, cfg, valid, res, cigs in golden_cases(z, "epi8"):
        mode, bw, M, X, go1, ge1, go2, ge2 = cfg
        got = ctx.epi8_batch(batch, mode, bw, synth.score_matrix(M, X), go1, ge1, go2, ge2)
        gc = got.cigars()
        for i in np.nonzero(valid)[0]:
            assert np.array_equal(got.results[i], res[i]) and np.array_equal(gc[i], cigs[i]), (cfg, int(i))
            n += 1
    z, batch = load_golden("edit_golden.npz")
    for ci, cfg, valid, res, cigs in golden_cases(z, "edit"):
        got = ctx.edit_batch(batch, cfg[0], cfg[1])
        gc = got.cigars()
        for i in np.nonzero(valid)[0]:
            assert np.array_equal(got.results[i], res[i]) and np.array_equal(gc[i], cigs[i]), (cfg, int(i))
            n += 1
    assert n > 1200


def test_readme_example(ctx):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "readme_pair.npz"))
    rs, cg, st = api.banded_striped_epi8_seqalign_pairwise(z["q"], z["t"], api.SEQALIGN_MODE_OVERLAP, int(z["bandwidth"]),
                                                           api.banded_striped_epi8_seqalign_set_score_matrix(2, -2), -4, -2, 0, 0, ctx=ctx)
    assert st == 0 and rs["score"] == 128 and rs["mat"] == 71 and rs["mis"] == 4
    assert np.array_equal(cg, z["cig"]) and [rs[k] for k in api.RESULT_FIELDS] == list(z["res"])


def test_baseline_config_shapes(ctx):
    """configs[0] (single 1 kb pair, global, band 128), config 2/3/4 shapes at oracle-sized counts."""
    b = synth.make_pairs(1, 1000, seed=42)
    exp, ecg, _ = ck.oracle_batch("epi8", b, 0, 128, M26, (-3, -2, 0, 0))
    assert_same(ctx.epi8_batch(b, 0, 128, M26, -3, -2, 0, 0), exp, ecg, tag="c1")
    b = synth.make_pairs(200, 1000, seed=1000)
    exp, ecg, _ = ck.oracle_batch("epi8", b, 0, 0, M26, (-3, -2, 0, 0), nthreads=8)
    assert_same(ctx.epi8_batch(b, 0, 0, M26, -3, -2, 0, 0), exp, ecg, tag="c2")
    b = synth.make_pairs(12, 10000, 2000, *synth.ont_like(0.12))
    exp, ecg, _ = ck.oracle_batch("epi8", b, 1, 512, M26, (-3, -2, 0, 0), nthreads=8)
    assert_same(ctx.epi8_batch(b, 1, 512, M26, -3, -2, 0, 0), exp, ecg, tag="c3")
    b = synth.make_pairs(3000, 300, 3000, 0.02, 0.02, 0.02)
    exp, ecg, _ = ck.oracle_batch("edit", b, 0, 64, nthreads=8)
    assert_same(ctx.edit_batch(b, 0, 64), exp, ecg, tag="c4")


def test_adversarial_and_reference_ub_flags(ctx):
    """Tiny, ragged, unrelated, long-indel and homopolymer pairs with odd bands and saturating scores.  Pairs on
    which the reference reads out of bounds / never terminates must come back flagged, all others bit-exact."""
    rng = np.random.default_rng(5)
    params = [((2, -6), (-3, -2, 0, 0)), ((2, -6), (-3, -2, -8, -1)), ((30, -40), (-40, -20, 0, 0)), ((5, -4), (-10, -1, 0, 0)), ((1, -1), (0, -1, 0, 0))]
    for it in range(10):
        b = _random_pairs(rng, 32, *[(1, 17), (20, 120), (100, 500), (300, 900)][it % 4], 0.2)
        (Mv, Xv), gaps = params[it % len(params)]
        mtx = synth.score_matrix(Mv, Xv)
        for mode in (0, 1, 2):
            bw = int(rng.choice([0, 16, 48, 128, 1024]))
            errs = np.zeros(b.n, np.int32)
            exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, mtx, gaps, errs=errs)
            assert_same(ctx.epi8_batch(b, mode, bw, mtx, *gaps), exp, ecg, errs=errs, tag=("epi8", it, mode, bw))
            bwe = int(rng.choice([0, 1, 64, 192]))
            errs = np.zeros(b.n, np.int32)
            exp, ecg, _ = ck.oracle_batch("edit", b, mode, bwe, errs=errs)
            assert_same(ctx.edit_batch(b, mode, bwe), exp, ecg, errs=errs, tag=("edit", it, mode, bwe))


def test_empty_pairs_and_mixed_lengths(ctx):
    pairs = [(np.zeros(0, np.uint8), np.array([1, 2, 3], np.uint8)), (np.array([0, 1, 2, 3] * 20, np.uint8), np.array([0, 1, 2, 3] * 19, np.uint8)),
             (np.array([2], np.uint8), np.zeros(0, np.uint8)), (np.array([1], np.uint8), np.array([1], np.uint8))]
    b = synth.PairBatch.from_lists(pairs)
    for kind in ("epi8", "edit"):
        exp, ecg, _ = ck.oracle_batch(kind, b, 0, 0, M26, (-3, -2, 0, 0))
        got = ctx.epi8_batch(b, 0, 0, M26, -3, -2, 0, 0) if kind == "epi8" else ctx.edit_batch(b, 0, 0)
        assert got.status[0] == api.ST_EMPTY and got.status[2] == api.ST_EMPTY and not got.results[0].any()
        for i in (1, 3):
            assert np.array_equal(got.results[i], exp[i]) and np.array_equal(got.cigar(i), ecg[i])


def test_waves_and_single_pair_api_agree(ctx):
    """A trace budget that forces several waves, and the single-pair entry point, give the one-wave answer."""
    b = synth.make_pairs(300, 400, seed=77)
    one = ctx.epi8_batch(b, 0, 0, M26, -3, -2, 0, 0)
    small = api.Context(0, trace_budget_bytes=40 << 20)
    many = small.epi8_batch(b, 0, 0, M26, -3, -2, 0, 0)
    assert small.timing()["waves"] > 1
    assert np.array_equal(one.results, many.results) and np.array_equal(one.ncigar, many.ncigar)
    assert all(np.array_equal(x, y) for x, y in zip(one.cigars(), many.cigars()))
    small.close()
    rs, cg, st = api.banded_striped_epi8_seqalign_pairwise(b.query(5), b.target(5), 0, 0, M26, -3, -2, 0, 0, ctx=ctx)
    assert [rs[k] for k in api.RESULT_FIELDS] == list(one.results[5]) and np.array_equal(cg, one.cigar(5))


def test_wide_bands(ctx):
    """Lanes longer than 64 steps carry sub-lane anchors in the trace; very wide bands run fewer groups per warp."""
    for qlen, n, mode, bw, gaps in [(2000, 6, 0, 0, (-3, -2, 0, 0)), (2500, 4, 1, 1600, (-3, -2, -8, -1)), (6000, 3, 0, 0, (-3, -2, 0, 0)), (10000, 3, 0, 0, (-3, -2, 0, 0))]:
        b = synth.make_pairs(n, qlen, seed=qlen + mode)
        exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, M26, gaps, nthreads=8)
        assert_same(ctx.epi8_batch(b, mode, bw, M26, *gaps), exp, ecg, tag=("wide", qlen, mode, bw))


def test_latency_and_throughput_variants(ctx):
    """Affine-gap batches that leave an SM with a few warps take the LAT instantiation of the forward kernel (short F
    chain, prefetched chunks); BSB200_LAT forces one or the other: both must give the oracle's answer."""
    try:
        for mode, bw, qlen, n in [(0, 0, 300, 40), (1, 64, 500, 40), (2, 128, 700, 24), (0, 0, 1000, 16), (1, 512, 3000, 6), (0, 0, 2100, 4)]:
            b = synth.make_pairs(n, qlen, seed=31 * qlen + mode)
            exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, M26, (-3, -2, 0, 0), nthreads=8)
            for lat in ("0", "1"):
                os.environ["BSB200_LAT"] = lat
                assert_same(ctx.epi8_batch(b, mode, bw, M26, -3, -2, 0, 0), exp, ecg, tag=("lat", lat, mode, bw, qlen))
                if bw == 0:   # full bands normally take the instantiation without the band-shift code: also run the general one
                    os.environ["BSB200_NOFULL"] = "1"
                    assert_same(ctx.epi8_batch(b, mode, bw, M26, -3, -2, 0, 0), exp, ecg, tag=("lat-nofull", lat, mode, bw, qlen))
                    os.environ.pop("BSB200_NOFULL")
    finally:
        os.environ.pop("BSB200_LAT", None)
        os.environ.pop("BSB200_NOFULL", None)


def test_every_forward_instantiation(ctx):
    """The forward kernel is a family of template instantiations chosen by gap model, band width and batch shape; the env hooks of
    the host side force each of them on inputs the oracle finishes quickly: literal arithmetic (BSB200_NOFAST) for affine and
    two-piece gaps, sub-lane anchors with every gap model, one-warp CTAs with 1-3 groups per warp (BSB200_GPW), with and without LAT."""
    keys = ("BSB200_NOFAST", "BSB200_GPW", "BSB200_LAT", "BSB200_NOFULL")
    try:
        cases = [(0, 0, 1200, 5), (1, 1100, 1500, 5), (2, 0, 1300, 4), (1, 96, 600, 12)]
        for mode, bw, qlen, n in cases:
            b = synth.make_pairs(n, qlen, seed=17 * qlen + mode)
            for gaps in [(-3, -2, 0, 0), (0, -2, 0, 0), (-3, -2, -8, -1)]:
                exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, M26, gaps, nthreads=8)
                for env in [{}, {"BSB200_NOFAST": "1"}, {"BSB200_GPW": "1"}, {"BSB200_GPW": "3", "BSB200_LAT": "0"}, {"BSB200_GPW": "2", "BSB200_LAT": "1", "BSB200_NOFULL": "1"},
                            {"BSB200_NOFAST": "1", "BSB200_GPW": "2"}]:
                    for k in keys:
                        os.environ.pop(k, None)
                    os.environ.update(env)
                    assert_same(ctx.epi8_batch(b, mode, bw, M26, *gaps), exp, ecg, tag=("inst", mode, bw, qlen, gaps, sorted(env.items())))
    finally:
        for k in keys:
            os.environ.pop(k, None)


def test_asymmetric_and_long_narrow_pairs(ctx):
    """Shapes that drive the band steering hard: a query against a window of it and against a target with long flanks (global mode
    hurries the band to the end with shifts of many cells, overlap / extend end inside the target), bands wider than the query,
    and long pairs under the narrowest bands.  Pairs on which the reference reads out of bounds must come back flagged."""
    rng = np.random.default_rng(11)
    pairs = []
    for k in range(24):
        ql = int(rng.integers(150, 1200))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        if k % 3 == 0:      # target = a mutated window of the query
            a_ = int(rng.integers(0, ql // 2)); w = q[a_:a_ + max(20, ql // 3)]
            t, _ = synth.mutate_batch(rng, w[None, :], 0.02, 0.02, 0.02)
        elif k % 3 == 1:    # target = mutated query with random flanks
            m, _ = synth.mutate_batch(rng, q[None, :], 0.03, 0.03, 0.03)
            t = np.concatenate([rng.integers(0, 4, int(rng.integers(0, 900))).astype(np.uint8), m, rng.integers(0, 4, int(rng.integers(0, 900))).astype(np.uint8)])
        else:               # unrelated
            t = rng.integers(0, 4, int(rng.integers(50, 2000))).astype(np.uint8)
        pairs.append((q, np.ascontiguousarray(t, dtype=np.uint8)))
    b = synth.PairBatch.from_lists(pairs)
    for mode in (0, 1, 2):
        for bw, gaps in [(16, (-3, -2, 0, 0)), (64, (-3, -2, -8, -1)), (256, (-3, -2, 0, 0)), (2048, (0, -2, 0, 0)), (0, (-3, -2, 0, 0))]:
            errs = np.zeros(b.n, np.int32)
            exp, ecg, _ = ck.oracle_batch("epi8", b, mode, bw, M26, gaps, errs=errs, nthreads=8)
            assert_same(ctx.epi8_batch(b, mode, bw, M26, *gaps), exp, ecg, errs=errs, tag=("asym", mode, bw, gaps))
        for bwe in (0, 64, 320):
            errs = np.zeros(b.n, np.int32)
            exp, ecg, _ = ck.oracle_batch("edit", b, mode, bwe, errs=errs, nthreads=8)
            assert_same(ctx.edit_batch(b, mode, bwe), exp, ecg, errs=errs, tag=("asym-edit", mode, bwe))
    long_b = synth.make_pairs(3, 30000, seed=77, p_sub=0.01, p_ins=0.01, p_del=0.01)
    for mode, bw in [(0, 16), (1, 32), (0, 128)]:
        errs = np.zeros(long_b.n, np.int32)
        exp, ecg, _ = ck.oracle_batch("epi8", long_b, mode, bw, M26, (-3, -2, 0, 0), errs=errs, nthreads=3)
        assert_same(ctx.epi8_batch(long_b, mode, bw, M26, -3, -2, 0, 0), exp, ecg, errs=errs, tag=("long", mode, bw))


def test_dense_fetch_equals_scattered_fetch(ctx):
    b = synth.make_pairs(500, 200, seed=99)
    a = ctx.epi8_batch(b, 1, 64, M26, -3, -2, 0, 0)
    d = ctx.epi8_batch(b, 1, 64, M26, -3, -2, 0, 0, dense=True)
    assert np.array_equal(a.results, d.results) and np.array_equal(a.ncigar, d.ncigar) and np.array_equal(a.status, d.status)
    assert all(np.array_equal(x, y) for x, y in zip(a.cigars(), d.cigars()))
    e1 = ctx.edit_batch(b, 0, 64)
    e2 = ctx.edit_batch(b, 0, 64, dense=True)
    assert np.array_equal(e1.results, e2.results) and all(np.array_equal(x, y) for x, y in zip(e1.cigars(), e2.cigars()))


def cigar_score(q, t, res, cig, mtx, go, ge):
    """Affine score of the path a CIGAR describes (global): independent of the DP."""
    x, y, s = int(res[1]), int(res[3]), 0
    for w in cig:
        op, ln = int(w) & 15, int(w) >> 4
        if op == 0:
            s += int(mtx[q[x:x + ln].astype(np.int64) * 4 + t[y:y + ln]].sum()); x += ln; y += ln
        elif op == 1:
            s += go + ge * ln; x += ln
        else:
            s += go + ge * ln; y += ln
    return s, x, y


def test_full_size_properties(ctx):
    """BASELINE config 2 shape at 20k pairs (beyond what the oracle finishes in seconds): every CIGAR must be a
    complete global path with the reported operation counts, scores must be consistent with it, and a second run must
    reproduce the first bit for bit.  Config 4 shape at 200k pairs: edit score = mismatches + indels."""
    b = synth.make_pairs(20000, 1000, seed=1234)
    r1 = ctx.epi8_batch(b, 0, 0, M26, -3, -2, 0, 0)
    r2 = ctx.epi8_batch(b, 0, 0, M26, -3, -2, 0, 0)
    assert np.array_equal(r1.results, r2.results) and np.array_equal(r1.cigar_arena, r2.cigar_arena)
    res = r1.results
    assert not r1.status.any()
    assert np.array_equal(res[:, 9], res[:, 5] + res[:, 6] + res[:, 7] + res[:, 8])
    assert np.array_equal(res[:, 2] - res[:, 1], res[:, 5] + res[:, 6] + res[:, 7])
    assert np.array_equal(res[:, 4] - res[:, 3], res[:, 5] + res[:, 6] + res[:, 8])
    assert (res[:, 1] == 0).all() and (res[:, 3] == 0).all() and np.array_equal(res[:, 2], b.qlen.astype(np.int32)) and np.array_equal(res[:, 4], b.tlen.astype(np.int32))
    # the CIGAR is a complete global path with the reported op counts; its affine score equals the reported score
    # except where the reference's own re-derived traceback is not score-consistent (about 2% of these pairs: the
    # compiled reference shows the same, e.g. pair 1164 of this batch), so equality is required of >= 95% only
    same = tot = 0
    for i in range(0, b.n, 97):
        cg = r1.cigar(i)
        s, x, y = cigar_score(b.query(i), b.target(i), res[i], cg, M26, -3, -2)
        assert (x, y) == (int(b.qlen[i]), int(b.tlen[i])), i
        ops, lens = cg & 15, cg >> 4
        assert int(lens[ops == 0].sum()) == res[i, 5] + res[i, 6] and int(lens[ops == 1].sum()) == res[i, 7] and int(lens[ops == 2].sum()) == res[i, 8], i
        assert s <= int(res[i, 0]), i
        same += int(s == int(res[i, 0])); tot += 1
    assert same >= 0.95 * tot, (same, tot)
    sub = np.arange(0, b.n, 400)
    exp, ecg, _ = ck.oracle_batch("epi8", b.subset(sub), 0, 0, M26, (-3, -2, 0, 0), nthreads=8)
    for k, i in enumerate(sub):
        assert np.array_equal(res[i], exp[k]) and np.array_equal(r1.cigar(i), ecg[k])
    b = synth.make_pairs(200000, 300, seed=4321, p_sub=0.02, p_ins=0.02, p_del=0.02)
    e1 = ctx.edit_batch(b, 0, 64)
    res = e1.results
    assert not e1.status.any()
    assert np.array_equal(res[:, 0], res[:, 6] + res[:, 7] + res[:, 8])
    assert np.array_equal(res[:, 2] - res[:, 1], res[:, 5] + res[:, 6] + res[:, 7])
    assert np.array_equal(res[:, 4] - res[:, 3], res[:, 5] + res[:, 6] + res[:, 8])
    sub = np.arange(0, b.n, 2000)
    exp, ecg, _ = ck.oracle_batch("edit", b.subset(sub), 0, 64, nthreads=8)
    for k, i in enumerate(sub):
        assert np.array_equal(res[i], exp[k]) and np.array_equal(e1.cigar(i), ecg[k])


def test_real_ont_example_all_pairs(ctx):
    """All 12,477 ONT read pairs of the reference's own example (example/real.ont.b10M.txt) under the three configurations of
    example/run.sh:5-9, against the answers of the unmodified reference (tests/golden/real_ont.npz): ten result ints, cigar word
    counts and checksums of every pair."""
    import real_ont
    batch, cfgs, mtx, gaps = real_ont.load()
    for cfg in cfgs:
        if cfg["kind"] == "epi8":
            got = ctx.epi8_batch(batch, cfg["mode"], cfg["bandwidth"], mtx, *gaps, dense=True)
        else:
            got = ctx.edit_batch(batch, cfg["mode"], cfg["bandwidth"], dense=True)
        assert not got.status.any(), (cfg["kind"], cfg["bandwidth"], np.nonzero(got.status)[0][:5])
        total = int(got.ncigar.astype(np.int64).sum())
        assert real_ont.compare(cfg, np.arange(batch.n), got.results, got.ncigar, got.cigar_arena[:total]) == [], (cfg["kind"], cfg["bandwidth"])


def test_pairwise_compat_header_runs_against_reference():
    """include/bsalign_b200_compat.h EXECUTED: a C program built from the reference's own headers calls the reference functions and the
    re-bodied ones (epi8, edit, k-mer guided edit) on the same pairs (oracle/pairwise_dropin_test.c; built by oracle/Makefile where the reference tree exists, travels
    as a binary) - identical seqalign_result_t + cigar vectors, CIGRESV appends, empty edit input, and a flagged pair reaches the
    status hook instead of looking valid."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "pairwise_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/pairwise_dropin was not built (no reference tree at build time)")
    for args in (["40", "500", "3"], ["12", "2300", "9"]):
        out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        n = 3 * int(args[0])   # epi8, edit and the k-mer guided edit of every pair
        assert "identical=%d/%d" % (n, n) in out.stdout and "flagged_calls=0" not in out.stdout, out.stdout


def test_packed_2bit_upload_equals_byte_upload(ctx):
    """bsb200_batch_upload_bits: sequences in the reference's BaseBank word layout (dna.h:63), unpacked on the device."""
    for kind, n, qlen, mode, bw in (("epi8", 97, 333, 0, 0), ("epi8", 50, 1001, 1, 64), ("edit", 300, 300, 0, 64)):
        b = synth.make_pairs(n, qlen, seed=qlen + n)
        bits = api.pack_bits(b.seqs)
        # spot-check the layout against the reference's accessor: base i = bits[i >> 5] >> (((~i) & 31) << 1) & 3
        for i in (0, 1, 31, 32, 33, len(b.seqs) - 1):
            assert (int(bits[i >> 5]) >> (((~i) & 31) << 1)) & 3 == int(b.seqs[i])
        rb = ctx.upload_bits(kind, bits, b, mode, bw, M26, (-3, -2, 0, 0))
        rb.run()
        got = rb.fetch()
        rb.free()
        exp, ecg, _ = ck.oracle_batch(kind, b, mode, bw, M26, (-3, -2, 0, 0), nthreads=8)
        assert_same(got, exp, ecg, tag=("bits", kind, mode, bw))


def test_pointer_array_and_multi_device_entries(ctx):
    """bsb200_pairwise_batch_ptrs (one pointer per sequence, the reference's calling convention) and bsb200_pairwise_batch_multi (every
    GPU of the box from one process; one context twice when the box has a single GPU) give the arena entry point's answers."""
    import torch
    b = synth.make_pairs(120, 350, seed=21)
    exp, ecg, _ = ck.oracle_batch("epi8", b, 1, 64, M26, (-3, -2, 0, 0), nthreads=8)
    res, cgs, st = api.pairwise_batch_ptrs(ctx, "epi8", [b.query(i) for i in range(b.n)], [b.target(i) for i in range(b.n)], 1, 64, M26, (-3, -2, 0, 0))
    assert np.array_equal(res, exp) and not st.any() and all(np.array_equal(x, y) for x, y in zip(cgs, ecg))
    ndev = min(2, torch.cuda.device_count())
    ctxs = [ctx] + [api.Context(d) for d in range(1, ndev)]
    if len(ctxs) == 1:
        ctxs.append(api.Context(0))   # two contexts on the one device: still the multi-context code path
    for kind, mode, bw in (("epi8", 0, 0), ("edit", 0, 64)):
        exp, ecg, _ = ck.oracle_batch(kind, b, mode, bw, M26, (-3, -2, 0, 0), nthreads=8)
        assert_same(api.pairwise_batch_multi(ctxs, kind, b, mode, bw, M26, (-3, -2, 0, 0)), exp, ecg, tag=("multi", kind))
    for c in ctxs[1:]:
        c.close()


def test_command_line_text_equals_reference_on_real_example(tmp_path):
    """The whole command, files in, text out: bsalign_b200_cli (tools/bsalign_b200_cli.c -> bsb200_align_file: reader, BaseBank words, GPU
    batches, formatter) on the reference's own example under the three configurations of example/run.sh prints byte for byte what the
    reference prints (md5 + size of 31 MB of text each: tests/golden/real_ont_cli.json, equal to BASELINE.md's fingerprints)."""
    import hashlib
    import json
    import subprocess
    import real_ont
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "bsalign_b200", "bsalign_b200_cli")
    if not os.path.exists(exe):
        pytest.skip("bsalign_b200/bsalign_b200_cli not built (python __graft_entry__.py build)")
    gold = json.load(open(os.path.join(root, "tests", "golden", "real_ont_cli.json")))
    batch, _, _, _ = real_ont.load()
    fa = tmp_path / "real_ont.fa"
    with open(fa, "w") as f:
        for i in range(batch.n):
            f.write(">%d.1\n%s\n>%d.2\n%s\n" % (i, "".join("ACGT"[c] for c in batch.query(i)), i, "".join("ACGT"[c] for c in batch.target(i))))
    for name, g in gold.items():
        out = subprocess.run([exe] + g["args"] + [str(fa)], capture_output=True, timeout=600)
        assert out.returncode == 0, out.stderr.decode()
        assert len(out.stdout) == g["bytes"] and hashlib.md5(out.stdout).hexdigest() == g["md5"], name


def test_kmer_edit_matches_golden_and_oracle(ctx, monkeypatch):
    """The k-mer guided edit (bsb200_kmer_edit_batch, replaces kmer_striped_seqedit_pairwise bsalign.h:1209): the reference's golden answers,
    then seeded batches against the oracle - short and long pairs, pairs without anchors (plain global edit), the dense output form, a
    starved pool for large gap traces (retry rounds) and few warp slots."""
    from test_oracle import kmer_golden
    batch, cases = kmer_golden()
    for k, res, cigs in cases:
        got = ctx.kmer_edit_batch(batch, k, dense=(k == 9))
        assert np.array_equal(got.results, res), k
        assert all(np.array_equal(a, b) for a, b in zip(got.cigars(), cigs)), k
        st = got.status.copy()
        st[batch.qlen == 0] &= ~16
        assert not st.any()
    for seed, k in ((1, 13), (2, 3), (3, 15), (4, 2), (5, 7), (6, 1), (7, 9)):   # homopolymers, tandem repeats, reverse complements ...
        rb = ck.repetitive_pairs(seed, n=120)
        got = ctx.kmer_edit_batch(rb, k)
        exp, ecg, _ = ck.kmer_batch("oracle", rb, k, nthreads=4)
        assert np.array_equal(got.results, exp), (seed, k)
        assert all(np.array_equal(x, y) for x, y in zip(got.cigars(), ecg)), (seed, k)
    rng = np.random.default_rng(123)

    def related(n, qlen, p):
        q = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
        t, tl = synth.mutate_batch(rng, q, p, p, p)
        out, o = [], 0
        for i in range(n):
            out.append((q[i].copy(), t[o + (i % 4) * 5:o + tl[i]].copy()))
            o += tl[i]
        return out

    def check(b, k, **kw):
        got = ctx.kmer_edit_batch(b, k, **kw)
        exp, ecg, _ = ck.kmer_batch("oracle", b, k, nthreads=8)
        assert np.array_equal(got.results, exp), k
        assert all(np.array_equal(x, y) for x, y in zip(got.cigars(), ecg)), k
        return got
    pairs = related(3000, 300, .03) + related(300, 1000, .05) + related(500, 80, .1) + related(6, 15000, .03)
    pairs += [(rng.integers(0, 4, 90).astype(np.uint8), rng.integers(0, 4, 70).astype(np.uint8)) for _ in range(40)]
    b = synth.PairBatch.from_lists(pairs)
    check(b, 13)
    assert ctx.timing()["waves"] >= 40          # the unrelated pairs took the plain global edit
    check(b, 13, dense=True)
    check(b, 6)
    core = rng.integers(0, 4, 1500).astype(np.uint8)
    gap = np.concatenate([core[:200], rng.integers(0, 4, 900).astype(np.uint8), core[1200:]])
    big = synth.PairBatch.from_lists([(core, gap), (gap, core)] * 6 + related(4, 15000, .04))
    check(big, 13)
    monkeypatch.setenv("BSB200_KMER_POOL", "2000000")   # a few 900 x 900 gaps at a time
    check(big, 13)
    monkeypatch.setenv("BSB200_KMER_SLOTS", "4")
    check(b, 13)
    monkeypatch.delenv("BSB200_KMER_SLOTS")
    monkeypatch.delenv("BSB200_KMER_POOL")
    for grp in ("32", "7", "1"):   # pairs a warp works on at a time (serial phases: one lane per pair)
        monkeypatch.setenv("BSB200_KMER_GROUP", grp)
        check(b, 13)
        check(big, 13)
    monkeypatch.setenv("BSB200_KMER_POOL", "2000000")
    monkeypatch.setenv("BSB200_KMER_GROUP", "32")
    check(big, 13)


def test_pipelined_one_call_equals_plain_call(ctx, monkeypatch):
    """bsb200_pairwise_batch_dense / _dense_bits cut large edit batches into chunks that two streams work on (copies, host planning and
    kernels overlap): same results, same dense cigars as the unpipelined call, from bytes and from 2-bit words."""
    rng = np.random.default_rng(9)
    n = 300000
    q = rng.integers(0, 4, (n, 48)).astype(np.uint8)
    t, tl = synth.mutate_batch(rng, q, .03, .03, .03)
    qoff = np.arange(n, dtype=np.uint64) * 48
    toff = n * 48 + np.concatenate([[0], np.cumsum(tl[:-1].astype(np.uint64))]).astype(np.uint64)
    qlen = np.full(n, 48, np.uint32)
    qlen[[5, 37499, 37500, n - 1]] = 0          # empty pairs, also on chunk borders
    b = synth.PairBatch(np.concatenate([q.ravel(), t]), qoff, qlen, toff, tl.astype(np.uint32))
    monkeypatch.setenv("BSB200_NOPIPE", "1")
    plain = ctx.edit_batch(b, 0, 64, dense=True)
    assert ctx.last_timing["waves"] == 1
    monkeypatch.delenv("BSB200_NOPIPE")
    piped = ctx.edit_batch(b, 0, 64, dense=True)
    assert ctx.last_timing["waves"] > 1
    bits = api.pack_bits(b.seqs)
    packed = ctx.dense_bits("edit", bits, b, 0, 64)
    for got in (piped, packed):
        assert np.array_equal(got.results, plain.results) and np.array_equal(got.ncigar, plain.ncigar) and np.array_equal(got.status, plain.status)
        tot = int(plain.ncigar.sum())
        assert np.array_equal(got.cigar_arena[:tot], plain.cigar_arena[:tot])
    m = 2000
    sub = synth.PairBatch(b.seqs, b.qoff[:m], b.qlen[:m], b.toff[:m], b.tlen[:m])
    exp, ecg, _ = ck.oracle_batch("edit", sub, 0, 64)
    gc = piped.cigars()
    assert np.array_equal(piped.results[:m], exp) and all(np.array_equal(gc[i], ecg[i]) for i in range(m))
    # the k-mer guided edit pipelines its dense call the same way (from 400k pairs on)
    n2 = 420000
    idx = np.arange(n2) % n
    b2 = synth.PairBatch(b.seqs, b.qoff[idx], b.qlen[idx], b.toff[idx], b.tlen[idx])
    monkeypatch.setenv("BSB200_NOPIPE", "1")
    kplain = ctx.kmer_edit_batch(b2, 7, dense=True)
    monkeypatch.delenv("BSB200_NOPIPE")
    kpiped = ctx.kmer_edit_batch(b2, 7, dense=True)
    assert np.array_equal(kpiped.results, kplain.results) and np.array_equal(kpiped.ncigar, kplain.ncigar) and np.array_equal(kpiped.status, kplain.status)
    tot = int(kplain.ncigar.sum())
    assert np.array_equal(kpiped.cigar_arena[:tot], kplain.cigar_arena[:tot])
    exp, ecg, _ = ck.kmer_batch("oracle", sub, 7)
    gc = kpiped.cigars()
    assert np.array_equal(kpiped.results[:m], exp) and all(np.array_equal(gc[i], ecg[i]) for i in range(m))


def test_edit_pairs_beyond_the_kernel_band_limit(ctx):
    """striped_seqedit_pairwise has no length limit: unbanded edit alignments whose band exceeds edit_kernel's 16384 cells take
    edit_long_kernel (one thread per pair on scratch in HBM), next to ordinary pairs of the same batch; a moving band that wide is
    reported as BSB200_ST_UNSUPPORTED instead of failing the batch."""
    rng = np.random.default_rng(31)
    q = rng.integers(0, 4, (1, 17000)).astype(np.uint8)
    t, tl = synth.mutate_batch(rng, q, .03, .03, .03)
    small = synth.make_pairs(40, 300, seed=8)
    pairs = [(small.query(i), small.target(i)) for i in range(small.n)]
    pairs.insert(7, (q[0], t[50:int(tl[0]) - 30]))
    pairs.append((t[:int(tl[0])], q[0, 100:]))
    b = synth.PairBatch.from_lists(pairs)
    for mode in (0, 1, 2):
        got = ctx.edit_batch(b, mode, 0)
        exp, ecg, _ = ck.oracle_batch("edit", b, mode, 0, nthreads=4)
        assert np.array_equal(got.results, exp), mode
        assert all(np.array_equal(x, y) for x, y in zip(got.cigars(), ecg)), mode
        assert not got.status.any()
        dense = ctx.edit_batch(b, mode, 0, dense=True)
        assert np.array_equal(dense.results, exp) and all(np.array_equal(x, y) for x, y in zip(dense.cigars(), ecg)), mode
    # a band of 16448 cells that moves along a 40 kb query: not built, flagged, the rest of the batch is unaffected
    big = rng.integers(0, 4, 40000).astype(np.uint8)
    b2 = synth.PairBatch.from_lists(pairs[:5] + [(big, big[:39000].copy())])
    got = ctx.edit_batch(b2, 0, 16400)
    assert got.status[5] == 32 and not got.results[5].any() and not got.status[:5].any()
    exp, _, _ = ck.oracle_batch("edit", synth.PairBatch.from_lists(pairs[:5]), 0, 16400)
    assert np.array_equal(got.results[:5], exp)
