"""POA read-vs-graph sweep: host-side mirror of the reference's align_rd_bspoacore (bspoa.h:2515-2618) as a BATCH call.

A sweep job is what one call of align_rd_bspoacore consumes: the read (g->qseq + g->qb, g->slen), the selected
sub-graph (g->sels, out-edges restricted to selected nodes, in edge-list order), per node .base/.bonus/.rpos/.nct, the
scoring parameters of BSPOAPar (bspoa.h:55-76) with the real bandwidth g->bandwidth -- and what it produces: every
node's row block in the g->memp layout (dpalign_row_prepare_data, bspoa.h:1787-1793) and (maxscr, maxidx, maxoff).
There is NO CPU fallback: everything here calls libbsalign_b200.so (include/bsalign_b200.h).
"""
import ctypes

import numpy as np

from . import api

PAR_FIELDS = ("bandwidth", "alnmode", "M", "X", "O", "E", "Q", "P", "T", "refbonus")
DEFAULT_BSPOA_PAR = dict(bandwidth=128, alnmode=api.SEQALIGN_MODE_OVERLAP, M=2, X=-6, O=-3, E=-2, Q=-8, P=-1, T=20, refbonus=1)  # bspoa.h:78-80


def piecewise(O, E, Q, P, bw):
    """banded_striped_epi8_seqalign_get_piecewise, bsalign.h:2084-2092 (C division truncates toward zero)."""
    if Q < O and P > E and Q + P < O + E and int((O - Q) / (E - P)) < bw:
        return 2
    return 1 if O else 0


def block_bytes(par):
    """g->mmblk, bspoa.h:2217."""
    bw = int(par[0])
    pw = piecewise(int(par[4]), int(par[5]), int(par[6]), int(par[7]), bw)
    return (bw * (pw + 1) + 68 + 15) // 16 * 16


class SweepJob:
    """One sweep job; arrays are over LOCAL node ids (position in g->sels)."""
    __slots__ = ("par", "query", "base", "bonus", "rpos", "nct", "eoff", "edst", "head", "tail", "reoff", "resrc", "recov")

    def __init__(self, par, query, base, bonus, rpos, nct, eoff, edst, head, tail, reoff=None, resrc=None, recov=None):
        self.par = np.ascontiguousarray(par, dtype=np.int32)
        self.query = np.ascontiguousarray(query, dtype=np.uint8)
        self.base = np.ascontiguousarray(base, dtype=np.uint8)
        self.bonus = np.ascontiguousarray(bonus, dtype=np.uint8)
        self.rpos = np.ascontiguousarray(rpos, dtype=np.int32)
        self.nct = np.ascontiguousarray(nct, dtype=np.int32)
        self.eoff = np.ascontiguousarray(eoff, dtype=np.int32)
        self.edst = np.ascontiguousarray(edst, dtype=np.int32)
        self.head, self.tail = int(head), int(tail)
        # reverse edges (bspoanode_t.erev lists in list order, restricted to selected nodes) with bspoaedge_t.cov: only the
        # device-side traceback (alignment2graph_bspoa's walk) needs them
        self.reoff = None if reoff is None else np.ascontiguousarray(reoff, dtype=np.int32)
        self.resrc = None if resrc is None else np.ascontiguousarray(resrc, dtype=np.int32)
        self.recov = None if recov is None else np.ascontiguousarray(recov, dtype=np.int32)

    @property
    def nnode(self):
        return len(self.base)


class SweepBatch:
    """Packed arenas of many sweep jobs: the argument layout of bsb200_poa_upload / bsb200_poa_rows_batch."""

    def __init__(self, jobs):
        n = len(jobs)
        self.n = n
        self.par = np.ascontiguousarray(np.stack([j.par for j in jobs]) if n else np.zeros((0, 10)), dtype=np.int32)
        self.slen = np.array([len(j.query) for j in jobs], dtype=np.uint32)
        self.qoff = np.zeros(n, dtype=np.uint64)
        if n:
            self.qoff[1:] = np.cumsum(self.slen[:-1], dtype=np.uint64)
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
        self.queries = cat([j.query for j in jobs], np.uint8)
        nn = np.array([j.nnode for j in jobs], dtype=np.uint64)
        self.node_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(nn, out=self.node_off[1:])
        self.base = cat([j.base for j in jobs], np.uint8)
        self.bonus = cat([j.bonus for j in jobs], np.uint8)
        self.rpos = cat([j.rpos for j in jobs], np.int32)
        self.nct = cat([j.nct for j in jobs], np.int32)
        self.eoff = cat([j.eoff for j in jobs], np.int32)
        ne = np.array([len(j.edst) for j in jobs], dtype=np.uint64)
        self.edge_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(ne, out=self.edge_off[1:])
        self.edst = cat([j.edst for j in jobs], np.int32)
        self.head = np.array([j.head for j in jobs], dtype=np.uint32)
        self.tail = np.array([j.tail for j in jobs], dtype=np.uint32)
        self.has_rev = n > 0 and all(j.reoff is not None for j in jobs)
        if self.has_rev:
            self.reoff = cat([j.reoff for j in jobs], np.int32)
            nre = np.array([len(j.resrc) for j in jobs], dtype=np.uint64)
            self.redge_off = np.zeros(n + 1, dtype=np.uint64)
            np.cumsum(nre, out=self.redge_off[1:])
            self.resrc = cat([j.resrc for j in jobs], np.int32)
            self.recov = cat([j.recov for j in jobs], np.int32)
        self.blk = np.array([block_bytes(j.par) for j in jobs], dtype=np.uint64)
        self.row_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(nn * self.blk, out=self.row_off[1:])

    def args(self):
        p = api._ptr
        return (self.n, p(self.par), p(self.queries), p(self.qoff), p(self.slen), p(self.node_off), p(self.base), p(self.bonus),
                p(self.rpos), p(self.nct), p(self.eoff), p(self.edge_off), p(self.edst), p(self.head), p(self.tail))

    def rev_args(self):
        if not self.has_rev:
            raise ValueError("these sweep jobs carry no reverse edges (SweepJob(reoff=, resrc=, recov=))")
        p = api._ptr
        return (p(self.reoff), p(self.redge_off), p(self.resrc), p(self.recov))


class SweepResult:
    def __init__(self, batch, rows, best, status, ops, match=None, trace=None):
        self.batch, self.rows, self.best, self.status, self.ops = batch, rows, best, status, ops
        self.match_arena, self.trace = match, trace     # device-side walk of alignment2graph_bspoa (when reverse edges were given)

    def match(self, i):
        """Per read position of job i: local id of the node it is aligned to, or -1."""
        o = int(self.batch.qoff[i])
        return self.match_arena[o:o + int(self.batch.slen[i])]

    def blocks(self, i):
        """Row blocks of job i as (nnode, mmblk) uint8 -- the bytes of g->memp from block 2 on."""
        b = self.batch
        nn = int(b.node_off[i + 1] - b.node_off[i])
        return self.rows[int(b.row_off[i]):int(b.row_off[i + 1])].reshape(nn, int(b.blk[i]))

    def linear(self, i):
        """(rows[nnode, 3, bw] int8 in linear band order, ubegs[nnode, 17] int32) of job i (arrays beyond pw are zero)."""
        b = self.batch
        par = b.par[i]
        bw = int(par[0]); W = bw // 16
        pw = piecewise(int(par[4]), int(par[5]), int(par[6]), int(par[7]), bw)
        blk = self.blocks(i)
        nn = blk.shape[0]
        out = np.zeros((nn, 3, bw), dtype=np.int8)
        p = np.arange(bw)
        idx = (p % W) * 16 + p // W                       # banded_striped_epi8_pos2idx, bsalign.h:321
        for a in range(pw + 1):
            out[:, a, :] = blk[:, a * bw:(a + 1) * bw].view(np.int8)[:, idx]
        ub = np.ascontiguousarray(blk[:, bw * (pw + 1):bw * (pw + 1) + 68]).view(np.int32).reshape(nn, 17)
        return out, ub


def _bind(L):
    if getattr(L, "_poa_bound", False):
        return
    P = ctypes.c_void_p
    job_args = [ctypes.c_uint32] + [P] * 14
    L.bsb200_poa_upload.restype = P
    L.bsb200_poa_upload.argtypes = [P] + job_args
    L.bsb200_poa_run.argtypes = [P, P]
    L.bsb200_poa_rows_bytes.restype = ctypes.c_uint64
    L.bsb200_poa_rows_bytes.argtypes = [P, P]
    L.bsb200_poa_fetch.argtypes = [P, P, P, P, P, P]
    L.bsb200_poa_free.argtypes = [P, P]
    L.bsb200_poa_rows_batch.argtypes = [P] + job_args + [P, P, P, P]
    L.bsb200_poa_block_bytes.restype = ctypes.c_uint32
    L.bsb200_poa_block_bytes.argtypes = [P]
    L.bsb200_poa_attach_reverse.argtypes = [P, P, P, P, P, P]
    L.bsb200_poa_fetch_trace.argtypes = [P, P, P, P]
    L.bsb200_poa_align_batch.argtypes = [P] + job_args + [P, P, P, P] + [P, P, P, P, P, P]
    L._poa_bound = True


def _alloc(batch, want_rows=True, rows=None):
    if rows is None and want_rows:
        rows = np.zeros(int(batch.row_off[-1]), dtype=np.uint8)
    return rows, np.zeros((batch.n, 3), dtype=np.int32), np.zeros(batch.n, dtype=np.int32), np.zeros((batch.n, 2), dtype=np.uint64)


def poa_rows_batch(ctx, batch, want_rows=True, rows=None):
    """One-shot: host buffers in, host buffers out (bsb200_poa_rows_batch)."""
    L = ctx._lib
    _bind(L)
    rows, best, status, ops = _alloc(batch, want_rows, rows)
    rc = L.bsb200_poa_rows_batch(ctx._h, *batch.args(), api._ptr(rows), api._ptr(best), api._ptr(status), api._ptr(ops))
    ctx._check(rc, "bsb200_poa_rows_batch")
    return SweepResult(batch, rows, best, status, ops)


def _alloc_trace(batch, match=None):
    if match is None:
        match = np.zeros(max(1, int(batch.slen.sum())), dtype=np.int32)
    return match, np.zeros((batch.n, 8), dtype=np.int32)


def poa_align_batch(ctx, batch, want_rows=False, rows=None, match=None):
    """One-shot sweep + device-side walk (bsb200_poa_align_batch): the row blocks stay in HBM unless want_rows."""
    L = ctx._lib
    _bind(L)
    rows, best, status, ops = _alloc(batch, want_rows, rows)
    match, trace = _alloc_trace(batch, match)
    rc = L.bsb200_poa_align_batch(ctx._h, *batch.args(), *batch.rev_args(), api._ptr(rows), api._ptr(best), api._ptr(status), api._ptr(ops),
                                  api._ptr(match), api._ptr(trace))
    ctx._check(rc, "bsb200_poa_align_batch")
    return SweepResult(batch, rows, best, status, ops, match, trace)


class ResidentSweeps:
    """Staged form: jobs resident in HBM; run() = kernels only."""

    def __init__(self, ctx, batch):
        _bind(ctx._lib)
        self.ctx, self.batch = ctx, batch
        self._h = ctx._lib.bsb200_poa_upload(ctx._h, *batch.args())
        if not self._h:
            raise RuntimeError("bsb200_poa_upload failed: %s" % ctx._lib.bsb200_last_error(ctx._h).decode())

    def attach_reverse(self):
        self.ctx._check(self.ctx._lib.bsb200_poa_attach_reverse(self.ctx._h, self._h, *self.batch.rev_args()), "bsb200_poa_attach_reverse")
        self.has_rev = True

    def run(self):
        self.ctx._check(self.ctx._lib.bsb200_poa_run(self.ctx._h, self._h), "bsb200_poa_run")

    def fetch(self, want_rows=True, rows=None):
        rows, best, status, ops = _alloc(self.batch, want_rows, rows)
        rc = self.ctx._lib.bsb200_poa_fetch(self.ctx._h, self._h, api._ptr(rows), api._ptr(best), api._ptr(status), api._ptr(ops))
        self.ctx._check(rc, "bsb200_poa_fetch")
        match = trace = None
        if getattr(self, "has_rev", False):
            match, trace = _alloc_trace(self.batch)
            rc = self.ctx._lib.bsb200_poa_fetch_trace(self.ctx._h, self._h, api._ptr(match), api._ptr(trace))
            self.ctx._check(rc, "bsb200_poa_fetch_trace")
        return SweepResult(self.batch, rows, best, status, ops, match, trace)

    def free(self):
        if self._h:
            self.ctx._lib.bsb200_poa_free(self.ctx._h, self._h)
            self._h = None
