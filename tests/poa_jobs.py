"""POA sweep jobs for the parity tests (test infrastructure).

A *sweep job* is what one call of the reference's align_rd_bspoacore (bspoa.h:2515-2618) consumes and produces:
the read (query), the selected sub-graph in the reference's own edge order (CSR), per-node band offset / base /
bonus / in-degree, the scoring parameters -- and every node row plus the best end (maxscr, maxidx, maxoff).
`ref_dump` drives a whole BSPOA job through the UNMODIFIED reference (oracle/_ref/libbsref.so, bsref_poa_dump in
oracle/ref_harness.c) and parses the dump; it only works where the reference tree was present at build time.
"""
import ctypes

import numpy as np

import checkers as ck

HDR_WORDS = 24
MAGIC = 0x504F4131


class SweepJob:
    __slots__ = ("bw", "pw", "slen", "alnmode", "M", "X", "O", "E", "Q", "P", "T", "refbonus", "nnode", "head", "tail", "nedge",
                 "query", "base", "bonus", "rpos", "nct", "eoff", "edst", "maxscr", "maxidx", "maxoff", "rows", "ub", "done",
                 "reoff", "resrc", "recov", "rs", "merged", "qb")

    def params(self):
        return np.array([self.bw, self.alnmode, self.M, self.X, self.O, self.E, self.Q, self.P, self.T, self.refbonus], dtype=np.int32)

    def to_api(self):
        """The product-side job object (bsalign_b200.poa.SweepJob) with the same content."""
        from bsalign_b200 import poa
        return poa.SweepJob(self.params(), self.query, self.base, self.bonus, self.rpos, self.nct, self.eoff, self.edst, self.head, self.tail,
                            reoff=self.reoff, resrc=self.resrc, recov=self.recov)


def parse_dump(blob, with_rows=True):
    """blob: uint8 array written by bsref_poa_dump -> list of SweepJob."""
    jobs = []
    pos = 0
    n = len(blob)
    pad4 = lambda x: (x + 3) & ~3
    while pos < n:
        hdr = blob[pos:pos + 4 * HDR_WORDS].view(np.int32)
        assert hdr[0] == MAGIC, "bad magic at %d" % pos
        pos += 4 * HDR_WORDS
        j = SweepJob()
        (j.bw, j.pw, j.slen, j.alnmode, j.M, j.X, j.O, j.E, j.Q, j.P, j.T, j.refbonus, j.nnode, j.head, j.tail, j.nedge,
         j.maxscr, j.maxidx, j.maxoff) = [int(x) for x in hdr[1:20]]
        j.query = blob[pos:pos + j.slen].copy(); pos += pad4(j.slen)
        node = blob[pos:pos + 20 * j.nnode].view(np.int32).reshape(j.nnode, 5); pos += 20 * j.nnode
        j.base = node[:, 0].astype(np.uint8); j.bonus = node[:, 1].astype(np.uint8)
        j.rpos = node[:, 2].astype(np.int32); j.nct = node[:, 3].astype(np.int32)
        j.eoff = blob[pos:pos + 4 * (j.nnode + 1)].view(np.int32).copy(); pos += 4 * (j.nnode + 1)
        j.edst = blob[pos:pos + 4 * j.nedge].view(np.int32).copy(); pos += 4 * j.nedge
        nre = int(hdr[21]); j.qb = int(hdr[23])
        j.reoff = blob[pos:pos + 4 * (j.nnode + 1)].view(np.int32).copy(); pos += 4 * (j.nnode + 1)
        j.resrc = blob[pos:pos + 4 * nre].view(np.int32).copy(); pos += 4 * nre
        j.recov = blob[pos:pos + 4 * nre].view(np.int32).copy(); pos += 4 * nre
        rec = 3 * j.bw + 68
        assert rec % 4 == 0
        body = blob[pos:pos + rec * j.nnode].reshape(j.nnode, rec); pos += rec * j.nnode
        if with_rows:
            j.rows = body[:, :3 * j.bw].view(np.int8).reshape(j.nnode, 3, j.bw).copy()
            j.ub = np.ascontiguousarray(body[:, 3 * j.bw:]).view(np.int32).reshape(j.nnode, 17).copy()
        else:
            j.rows = None; j.ub = None
        j.done = blob[pos:pos + j.nnode].copy(); pos += pad4(j.nnode)
        j.done[j.tail] = 0  # the tail is only counted (v->vst++), it has no row
        j.rs = blob[pos:pos + 40].view(np.int32).copy(); pos += 40          # seqalign_result_t of alignment2graph_bspoa
        j.merged = blob[pos:pos + 4 * j.slen].view(np.int32).copy(); pos += 4 * j.slen   # per read position: node it was merged into
        jobs.append(j)
    return jobs


def ref_dump(reads, par_override=None, with_rows=True):
    """reads: list of uint8 arrays (bases 0..3).  par_override: None or 10 ints
    (bandwidth, M, X, O, E, Q, P, T, refbonus, alnmode).  Returns the list of sweep jobs of the whole POA job."""
    return parse_dump(ref_dump_blob(reads, par_override), with_rows)


def load_golden():
    """Sweep jobs of tests/golden/poa_golden.npz (made by tests/golden/make_poa_golden.py from the reference)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poa_golden.npz"))
    jobs = []
    for ci in range(int(z["ncase"])):
        jobs += parse_dump(z["blob%d" % ci])
    return jobs


def ref_dump_blob(reads, par_override=None):
    lib = ck.ref()
    lens = np.array([len(r) for r in reads], dtype=np.uint32)
    off = np.zeros(len(reads), dtype=np.uint64)
    off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    seqs = np.ascontiguousarray(np.concatenate(reads), dtype=np.uint8)
    out = ctypes.c_void_p()
    out_len = ctypes.c_uint64()
    po = None if par_override is None else np.ascontiguousarray(par_override, dtype=np.int32)
    lib.bsref_poa_dump.restype = ctypes.c_int64
    nj = lib.bsref_poa_dump(ctypes.c_uint32(len(reads)), ck._ptr(seqs), ck._ptr(off), ck._ptr(lens), ck._ptr(po),
                            ctypes.byref(out), ctypes.byref(out_len))
    blob = np.ctypeslib.as_array(ctypes.cast(out, ctypes.POINTER(ctypes.c_uint8)), shape=(out_len.value,)).copy() if out_len.value else np.zeros(0, np.uint8)
    lib.bsref_free.argtypes = [ctypes.c_void_p]
    lib.bsref_free(out)
    return blob


def make_reads(nreads, tlen, seed, p_sub=0.03, p_ins=0.03, p_del=0.04):
    """nreads reads mutated independently from one random template."""
    from bsalign_b200 import synth
    rng = np.random.default_rng(seed)
    tmpl = rng.integers(0, 4, size=(1, tlen), dtype=np.uint8)
    reads = []
    for _ in range(nreads):
        flat, _l = synth.mutate_batch(rng, tmpl, p_sub, p_ins, p_del)
        reads.append(flat.astype(np.uint8))
    return reads


def oracle_sweep(j):
    """Run oracle/bsalign_oracle.c:bso_poa_sweep on one SweepJob -> (rows[nnode,3,bw] int8 linear, ub[nnode,17], done, best[3], ops[2], rc)."""
    lib = ck.oracle()
    par = j.params()
    rows = np.zeros((j.nnode, 3, j.bw), dtype=np.int8)
    ub = np.zeros((j.nnode, 17), dtype=np.int32)
    done = np.zeros(j.nnode, dtype=np.uint8)
    best = np.zeros(3, dtype=np.int32)
    ops = np.zeros(2, dtype=np.uint64)
    q = np.ascontiguousarray(j.query, dtype=np.uint8)
    base = np.ascontiguousarray(j.base, dtype=np.uint8)
    bonus = np.ascontiguousarray(j.bonus, dtype=np.uint8)
    rpos = np.ascontiguousarray(j.rpos, dtype=np.int32)
    nct = np.ascontiguousarray(j.nct, dtype=np.int32)
    eoff = np.ascontiguousarray(j.eoff, dtype=np.int32)
    edst = np.ascontiguousarray(j.edst, dtype=np.int32)
    rc = lib.bso_poa_sweep(ck._ptr(par), ck._ptr(q), ctypes.c_uint32(j.slen), ctypes.c_uint32(j.nnode), ck._ptr(base), ck._ptr(bonus),
                           ck._ptr(rpos), ck._ptr(nct), ck._ptr(eoff), ck._ptr(edst), ctypes.c_uint32(j.head), ctypes.c_uint32(j.tail),
                           ck._ptr(rows), ck._ptr(ub), ck._ptr(done), ck._ptr(best), ck._ptr(ops))
    return rows, ub, done, best, ops, rc


def compare_rows(j, rows, ub, done, best, pw=None):
    """Mismatch description (or None) between a sweep result and the reference dump held by job j."""
    pw = j.pw if pw is None else pw
    if (int(best[0]), int(best[1]), int(best[2])) != (j.maxscr, j.maxidx, j.maxoff):
        return "best %s != ref %s" % (best.tolist(), (j.maxscr, j.maxidx, j.maxoff))
    if not np.array_equal(done.astype(bool), j.done.astype(bool)):
        return "visited sets differ"
    m = j.done.astype(bool)
    if not np.array_equal(ub[m], j.ub[m]):
        bad = np.nonzero((ub != j.ub).any(axis=1) & m)[0]
        return "ubegs differ at nodes %s" % bad[:5].tolist()
    for a in range(pw + 1):
        if not np.array_equal(rows[m, a], j.rows[m, a]):
            bad = np.nonzero((rows[:, a] != j.rows[:, a]).any(axis=1) & m)[0]
            return "array %d differs at nodes %s" % (a, bad[:5].tolist())
    return None


def oracle_sweep_batch(batch, nthreads=1, want_rows=True):
    """oracle/bsalign_oracle.c:bso_poa_sweep_batch on a packed bsalign_b200.poa.SweepBatch.
    Returns dict(rows=list of [nnode,3,bw] linear arrays or None, ub, done, best, ops, seconds)."""
    import time
    lib = ck.oracle()
    n = batch.n
    nn = (batch.node_off[1:] - batch.node_off[:-1]).astype(np.uint64)
    bw = batch.par[:, 0].astype(np.uint64) if n else np.zeros(0, np.uint64)
    roff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(nn * bw * np.uint64(3), out=roff[1:])
    rows = np.zeros(int(roff[-1]), dtype=np.int8) if want_rows else None
    tot = int(batch.node_off[-1]) if n else 0
    ub = np.zeros((tot, 17), dtype=np.int32) if want_rows else None
    done = np.zeros(tot, dtype=np.uint8) if want_rows else None
    best = np.zeros((n, 3), dtype=np.int32)
    ops = np.zeros((n, 2), dtype=np.uint64)
    p = ck._ptr
    t0 = time.perf_counter()
    rc = lib.bso_poa_sweep_batch(ctypes.c_uint32(n), p(batch.par), p(batch.queries), p(batch.qoff), p(batch.slen), p(batch.node_off), p(batch.base), p(batch.bonus),
                                 p(batch.rpos), p(batch.nct), p(batch.eoff), p(batch.edge_off), p(batch.edst), p(batch.head), p(batch.tail),
                                 p(rows), p(roff), p(ub), p(done), p(best), p(ops), ctypes.c_int(nthreads))
    dt = time.perf_counter() - t0
    out_rows = None
    if want_rows:
        out_rows = [rows[int(roff[i]):int(roff[i + 1])].reshape(int(nn[i]), 3, int(bw[i])) for i in range(n)]
    return dict(rows=out_rows, ub=ub, done=done, best=best, ops=ops, seconds=dt, rc=rc)


def ref_time_jobs(read_sets, nthreads, par_override=None):
    """Run whole BSPOA jobs through the unmodified reference (bsref_poa_time), one job per call on `nthreads` host threads.
    Returns dict(dp_seconds = CPU seconds inside align_rd_bspoacore summed over jobs, total_seconds, nupd, nmrg, cells, wall)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    lib = ck.ref()
    po = None if par_override is None else np.ascontiguousarray(par_override, dtype=np.int32)

    def one(reads):
        lens = np.array([len(r) for r in reads], dtype=np.uint32)
        off = np.zeros(len(reads), dtype=np.uint64)
        off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        seqs = np.ascontiguousarray(np.concatenate(reads), dtype=np.uint8)
        dp, tot = ctypes.c_double(), ctypes.c_double()
        nu, nm, ce = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        lib.bsref_poa_time(ctypes.c_uint32(len(reads)), ck._ptr(seqs), ck._ptr(off), ck._ptr(lens), ck._ptr(po),
                           ctypes.byref(dp), ctypes.byref(tot), ctypes.byref(nu), ctypes.byref(nm), ctypes.byref(ce))
        return dp.value, tot.value, nu.value, nm.value, ce.value
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=nthreads) as ex:
        rs = list(ex.map(one, read_sets))
    wall = time.perf_counter() - t0
    a = np.array(rs, dtype=np.float64)
    return dict(dp_seconds=float(a[:, 0].sum()), total_seconds=float(a[:, 1].sum()), nupd=int(a[:, 2].sum()), nmrg=int(a[:, 3].sum()),
                cells=int(a[:, 4].sum()), wall=wall, jobs=len(read_sets))


def oracle_backtrace(j, rows, ub, midx, xe):
    """oracle/bsalign_oracle.c:bso_poa_backtrace on one dumped job with the given (linear) rows -> (match[slen], out[8])."""
    lib = ck.oracle()
    par = j.params()
    match = np.zeros(max(1, j.slen), dtype=np.int32)
    out = np.zeros(8, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.int8); ub = np.ascontiguousarray(ub, dtype=np.int32)
    q = np.ascontiguousarray(j.query, dtype=np.uint8); base = np.ascontiguousarray(j.base, dtype=np.uint8); bonus = np.ascontiguousarray(j.bonus, dtype=np.uint8)
    rpos = np.ascontiguousarray(j.rpos, dtype=np.int32)
    lib.bso_poa_backtrace(ck._ptr(par), ck._ptr(q), ctypes.c_uint32(j.slen), ctypes.c_uint32(j.nnode), ck._ptr(base), ck._ptr(bonus), ck._ptr(rpos),
                          ck._ptr(j.reoff), ck._ptr(j.resrc), ck._ptr(j.recov), ctypes.c_uint32(j.head), ctypes.c_uint32(j.tail),
                          ck._ptr(rows), ck._ptr(ub), ctypes.c_int32(midx), ctypes.c_int32(xe), ck._ptr(match), ck._ptr(out))
    return match[:j.slen], out


def compare_backtrace(j, match, out):
    """Mismatch description (or None) between a walk result and what the reference's alignment2graph_bspoa did (dump of job j)."""
    rs = j.rs   # score, qb, qe, tb, te, mat, mis, ins, del, aln
    if int(out[7]) != 0:
        return "flags %d" % int(out[7])
    if (int(out[2]), int(out[3]), int(out[4]), int(out[5])) != (int(rs[5]), int(rs[6]), int(rs[7]), int(rs[8])):
        return "counts %s != ref %s" % (out[2:6].tolist(), rs[5:9].tolist())
    if int(out[0]) + j.qb != int(rs[1]):
        return "qb %d != ref %d" % (int(out[0]) + j.qb, int(rs[1]))
    # every position the reference merged into a graph node must be matched to that node; matched-but-not-merged = mismatches
    m = j.merged >= 0
    if not np.array_equal(match[m], j.merged[m]):
        bad = np.nonzero(m & (match != j.merged))[0]
        return "merged nodes differ at read positions %s" % bad[:5].tolist()
    if int((match >= 0).sum()) != int(rs[5]) + int(rs[6]):
        return "matched positions %d != mat + mis" % int((match >= 0).sum())
    return None
