"""ctypes loaders for the two CHECKERS (test infrastructure): our CPU restatement (oracle/libbsoracle.so)
and, when it was built in the container, the unmodified reference (oracle/_ref/libbsref.so)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libbsoracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbsref.so")

_P = ctypes.c_void_p
_I8 = ctypes.c_int8


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "bsalign_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libbsoracle.so"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None
last_call_seconds = 0.0  # wall time of the most recent C call made by run_batch (excludes numpy/python packing)


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        _oracle = ctypes.CDLL(ORACLE_SO)
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_SO)
    return _ref


def cigar_caps(batch):
    cap = (batch.qlen.astype(np.uint64) + batch.tlen.astype(np.uint64) + 8)
    off = np.zeros(batch.n + 1, dtype=np.uint64)
    np.cumsum(cap, out=off[1:])
    return off


def split_cigars(arena, off, ncg):
    return [arena[int(off[i]):int(off[i]) + int(ncg[i])].copy() for i in range(len(ncg))]


def run_batch(lib, prefix, kind, batch, mode, bandwidth, mtx=None, gaps=(0, 0, 0, 0), nthreads=1, repeat=1, want_cigar=True, errs=None):
    """kind: 'epi8' or 'edit'.  Returns (results[n,10] int32, list of cigar arrays or None, rc).
    errs: optional int32[n] array receiving per-pair anomaly flags (oracle only)."""
    n = batch.n
    extra = () if errs is None else (_ptr(errs),)
    res = np.zeros((n, 10), dtype=np.int32)
    off = cigar_caps(batch) if want_cigar else None
    arena = np.zeros(int(off[-1]), dtype=np.uint32) if want_cigar else None
    ncg = np.zeros(n, dtype=np.uint32)
    global last_call_seconds
    import time as _time
    fn = getattr(lib, "%s_%s_batch%s" % (prefix, kind, "" if errs is None else "_ex"))
    _t0 = _time.perf_counter()
    if kind == "epi8":
        m = np.ascontiguousarray(mtx, dtype=np.int8)
        rc = fn(ctypes.c_uint64(n), _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                ctypes.c_int(mode), ctypes.c_uint32(bandwidth), _ptr(m), _I8(gaps[0]), _I8(gaps[1]), _I8(gaps[2]), _I8(gaps[3]),
                _ptr(res), _ptr(arena), _ptr(off), _ptr(ncg), ctypes.c_int(nthreads), ctypes.c_int(repeat), *extra)
    else:
        rc = fn(ctypes.c_uint64(n), _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                ctypes.c_int(mode), ctypes.c_uint32(bandwidth),
                _ptr(res), _ptr(arena), _ptr(off), _ptr(ncg), ctypes.c_int(nthreads), ctypes.c_int(repeat), *extra)
    last_call_seconds = _time.perf_counter() - _t0
    cigs = split_cigars(arena, off, ncg) if want_cigar else None
    return res, cigs, rc


def kmer_caps(batch):
    cap = 2 * (batch.qlen.astype(np.uint64) + batch.tlen.astype(np.uint64)) + 8
    off = np.zeros(batch.n + 1, dtype=np.uint64)
    np.cumsum(cap, out=off[1:])
    return off


def kmer_batch(which, batch, ksz, nthreads=1, repeat=1, errs=None):
    """kmer_striped_seqedit_pairwise (bsalign.h:1209) over a batch: which = 'ref' (the compiled reference; it reverses sequence prefixes in
    place and restores them, so it gets its own copy of the arena) or 'oracle'.  Returns (results, cigars, rc)."""
    n = batch.n
    res = np.zeros((n, 10), dtype=np.int32)
    off = kmer_caps(batch)
    arena = np.zeros(int(off[-1]), dtype=np.uint32)
    ncg = np.zeros(n, dtype=np.uint32)
    seqs = batch.seqs.copy()
    global last_call_seconds
    import time as _time
    _t0 = _time.perf_counter()
    if which == "ref":
        rc = ref().bsref_kmer_edit_batch(ctypes.c_uint64(n), _ptr(seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                         ctypes.c_uint32(ksz), _ptr(res), _ptr(arena), _ptr(off), _ptr(ncg), ctypes.c_int(nthreads), ctypes.c_int(repeat))
    else:
        rc = oracle().bso_kmer_edit_batch(ctypes.c_uint64(n), _ptr(seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                          ctypes.c_uint32(ksz), _ptr(res), _ptr(arena), _ptr(off), _ptr(ncg), ctypes.c_int(nthreads), ctypes.c_int(repeat),
                                          _ptr(errs) if errs is not None else None)
    last_call_seconds = _time.perf_counter() - _t0
    return res, split_cigars(arena, off, ncg), rc


def ref_batch(kind, batch, mode, bandwidth, mtx=None, gaps=(0, 0, 0, 0), **kw):
    return run_batch(ref(), "bsref", kind, batch, mode, bandwidth, mtx, gaps, **kw)


def oracle_batch(kind, batch, mode, bandwidth, mtx=None, gaps=(0, 0, 0, 0), **kw):
    return run_batch(oracle(), "bso", kind, batch, mode, bandwidth, mtx, gaps, **kw)


def rows_dump(lib, fname, q, t, mode, bandwidth, mtx, gaps, extra_args=()):
    qlen, tlen = len(q), len(t)
    bw = bandwidth if bandwidth else qlen
    bw = (bw + 15) // 16 * 16
    res = np.zeros(10, dtype=np.int32)
    begs = np.zeros(tlen, dtype=np.int32)
    ub = np.zeros((tlen, 17), dtype=np.int32)
    u = np.zeros((tlen, bw), dtype=np.int8)
    e = np.zeros((tlen, bw), dtype=np.int8)
    qq = np.zeros((tlen, bw), dtype=np.int8)
    m = np.ascontiguousarray(mtx, dtype=np.int8)
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    fn = getattr(lib, fname)
    fn(_ptr(q), ctypes.c_uint32(qlen), _ptr(t), ctypes.c_uint32(tlen), ctypes.c_int(mode), ctypes.c_uint32(bandwidth), _ptr(m),
       _I8(gaps[0]), _I8(gaps[1]), _I8(gaps[2]), _I8(gaps[3]), _ptr(res), *extra_args, _ptr(begs), _ptr(ub), _ptr(u), _ptr(e), _ptr(qq))
    return res, begs, ub, u, e, qq


def forked(fn, *args, timeout_s=120, **kw):
    """Run fn in a forked child so that a crash or a hang of the checker (the reference reads out of bounds, or
    never leaves its traceback loop, on some adversarial inputs) is reported as None instead of taking the
    test process with it."""
    import pickle
    import signal
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        signal.alarm(int(timeout_s))
        try:
            data = pickle.dumps(fn(*args, **kw))
            with os.fdopen(w, "wb") as f:
                f.write(data)
        finally:
            os._exit(0)
    os.close(w)
    with os.fdopen(r, "rb") as f:
        data = f.read()
    _, status = os.waitpid(pid, 0)
    if status != 0 or not data:
        return None
    return pickle.loads(data)


def repetitive_pairs(seed, n=60):
    """Pairs that stress the k-mer anchoring (kmer_striped_seqedit_pairwise): homopolymers, tandem repeats, two-letter sequences, internal
    duplications, reversed and reverse-complemented partners, lengths around the k-mer size, a few point differences."""
    from bsalign_b200 import synth
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        kind = int(rng.integers(0, 6))
        L = int(rng.integers(5, 400))
        if kind == 0:
            a = np.zeros(L, np.uint8)
        elif kind == 1:
            a = np.full(L, 3, np.uint8)
        elif kind == 2:
            u = rng.integers(0, 4, int(rng.integers(1, 7))).astype(np.uint8)
            a = np.tile(u, L // len(u) + 1)[:L]
        elif kind == 3:
            a = (rng.integers(0, 2, L) * 3).astype(np.uint8)
        elif kind == 4:
            a = rng.integers(0, 4, L).astype(np.uint8)
            m = L // 3
            a[m:2 * m] = a[:m]
        else:
            a = rng.integers(0, 4, L).astype(np.uint8)
        b = a.copy()
        for _ in range(int(rng.integers(0, 6))):
            p = int(rng.integers(0, len(b)))
            op = int(rng.integers(0, 3))
            if op == 0:
                b[p] = (b[p] + 1) & 3
            elif op == 1:
                b = np.insert(b, p, rng.integers(0, 4)).astype(np.uint8)
            elif len(b) > 2:
                b = np.delete(b, p)
        if rng.integers(0, 4) == 0:
            b = b[::-1].copy()
        if rng.integers(0, 5) == 0:
            b = (3 - b[::-1]).astype(np.uint8)
        out.append((a, b))
    return synth.PairBatch.from_lists(out)
