"""Host-side mirror of the reference's pairwise API on top of libbsalign_b200.so (C ABI: include/bsalign_b200.h).

Names, argument meaning and result layout follow the reference (ruanjue/bsalign):
  banded_striped_epi8_seqalign_pairwise   bsalign.h:3854 / :399
  striped_seqedit_pairwise                bsalign.h:1046 / :232
  banded_striped_epi8_seqalign_set_score_matrix  bsalign.h:323
  seqalign_cigar2alnstr                   bsalign.h:531
plus the batch forms a GPU needs.  There is NO CPU fallback: a missing library or device raises.
"""
import ctypes
import os

import numpy as np

from .synth import PairBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSB200_LIB") or os.path.join(_HERE, "libbsalign_b200.so")   # BSB200_LIB: development builds

SEQALIGN_MODE_GLOBAL = 0   # bsalign.h:30
SEQALIGN_MODE_OVERLAP = 1  # bsalign.h:31
SEQALIGN_MODE_EXTEND = 2   # bsalign.h:32

ST_RANGE, ST_LOOP, ST_CIGCAP, ST_REFBUG, ST_EMPTY = 1, 2, 4, 8, 16

RESULT_FIELDS = ("score", "qb", "qe", "tb", "te", "mat", "mis", "ins", "del", "aln")  # seqalign_result_t, bsalign.h:213-218

_P = ctypes.c_void_p
_I8 = ctypes.c_int8


class Timing(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("forward_ms", ctypes.c_float), ("traceback_ms", ctypes.c_float),
                ("d2h_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("forward_launches", ctypes.c_uint32), ("traceback_launches", ctypes.c_uint32),
                ("other_launches", ctypes.c_uint32), ("waves", ctypes.c_uint32),
                ("cells", ctypes.c_uint64), ("trace_bytes", ctypes.c_uint64),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("run_ms", ctypes.c_float), ("reserved", ctypes.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Load the CUDA library; fail loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("bsalign_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.bsb200_create.restype = _P
        L.bsb200_create.argtypes = [ctypes.c_int, ctypes.c_uint64]
        L.bsb200_destroy.argtypes = [_P]
        L.bsb200_last_error.restype = ctypes.c_char_p
        L.bsb200_last_error.argtypes = [_P]
        L.bsb200_version.restype = ctypes.c_char_p
        L.bsb200_get_timing.argtypes = [_P, ctypes.POINTER(Timing)]
        L.bsb200_batch_upload.restype = _P
        L.bsb200_batch_upload.argtypes = [_P, ctypes.c_int, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_uint32, _P,
                                          _I8, _I8, _I8, _I8, ctypes.c_int]
        L.bsb200_batch_run.argtypes = [_P, _P]
        L.bsb200_batch_sync.argtypes = [_P]
        L.bsb200_batch_fetch.argtypes = [_P, _P, _P, _P, _P, _P, _P]
        L.bsb200_batch_fetch_dense.argtypes = [_P, _P, _P, _P, ctypes.c_uint64, _P, _P, _P]
        L.bsb200_batch_free.argtypes = [_P, _P]
        L.bsb200_epi8_pairwise_batch.argtypes = [_P, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_uint32, _P,
                                                 _I8, _I8, _I8, _I8, _P, _P, _P, _P, _P]
        L.bsb200_edit_pairwise_batch.argtypes = [_P, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_uint32,
                                                 _P, _P, _P, _P, _P]
        L.bsb200_kmer_edit_batch.argtypes = [_P, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_uint32, _P, _P, _P, _P, _P]
        L.bsb200_kmer_edit_batch_dense.argtypes = [_P, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_uint32, _P, _P, ctypes.c_uint64, _P, _P, _P]
        L.bsb200_kmer_edit_pairwise.argtypes = [_P, ctypes.c_uint32, _P, ctypes.c_uint32, _P, ctypes.c_uint32, _P, _P, ctypes.c_uint32, _P, _P]
        L.bsb200_batch_upload_dev.restype = _P
        L.bsb200_batch_upload_dev.argtypes = L.bsb200_batch_upload.argtypes
        L.bsb200_pairwise_batch_dense.argtypes = [_P, ctypes.c_int, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_uint32, _P,
                                                  _I8, _I8, _I8, _I8, _P, _P, ctypes.c_uint64, _P, _P, _P]
        L.bsb200_pairwise_batch_dense_bits.argtypes = L.bsb200_pairwise_batch_dense.argtypes
        L.bsb200_remsa_batch.argtypes = [_P, ctypes.c_uint32, _P, _P, _P, ctypes.c_uint64, _P, _P, ctypes.c_uint64, _P, _P, _P]
        L.bsb200_pack_pairs_dev.argtypes = [_P, _P, _P, _P, _P, _P, _P, ctypes.c_uint64, _P]
        L.bsb200_pairwise_batch_ptrs.argtypes = [_P, ctypes.c_int, ctypes.c_uint64, _P, _P, _P, _P, ctypes.c_int, ctypes.c_uint32, _P,
                                                 _I8, _I8, _I8, _I8, _P, _P, _P, _P, ctypes.c_int]
        L.bsb200_pairwise_batch_multi.argtypes = [_P, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_uint32, _P,
                                                  _I8, _I8, _I8, _I8, _P, _P, _P, _P, _P]
        L.bsb200_seqfile_read.restype = _P
        L.bsb200_seqfile_read.argtypes = [ctypes.c_char_p]
        L.bsb200_seqfile_error.restype = ctypes.c_char_p
        L.bsb200_seqfile_error.argtypes = [_P]
        for fn in ("bsb200_seqfile_nseq", "bsb200_seqfile_nbases"):
            getattr(L, fn).restype = ctypes.c_uint64
            getattr(L, fn).argtypes = [_P]
        for fn in ("bsb200_seqfile_bits", "bsb200_seqfile_offsets", "bsb200_seqfile_lengths"):
            getattr(L, fn).restype = _P
            getattr(L, fn).argtypes = [_P]
        L.bsb200_seqfile_name.restype = ctypes.c_char_p
        L.bsb200_seqfile_name.argtypes = [_P, ctypes.c_uint64]
        L.bsb200_seqfile_free.argtypes = [_P]
        L.bsb200_format_pair_text.restype = ctypes.c_uint64
        L.bsb200_format_pair_text.argtypes = [_P, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, _P, _P, ctypes.c_uint64, ctypes.c_uint64, _P, ctypes.c_uint32]
        L.bsb200_msa_read.restype = _P
        L.bsb200_msa_read.argtypes = [_P]
        L.bsb200_msa_write.argtypes = [_P, ctypes.c_uint32, ctypes.c_uint32, _P, _P, _P, ctypes.c_char_p, ctypes.c_uint32]
        for fn in ("bsb200_msa_nseq", "bsb200_msa_mlen"):
            getattr(L, fn).restype = ctypes.c_uint32
            getattr(L, fn).argtypes = [_P]
        for fn in ("bsb200_msa_cols", "bsb200_msa_qlt", "bsb200_msa_alt"):
            getattr(L, fn).restype = _P
            getattr(L, fn).argtypes = [_P]
        L.bsb200_msa_meta.restype = _P
        L.bsb200_msa_meta.argtypes = [_P, ctypes.POINTER(ctypes.c_uint32)]
        L.bsb200_msa_free.argtypes = [_P]
        L.bsb200_batch_upload_bits.restype = _P
        L.bsb200_batch_upload_bits.argtypes = L.bsb200_batch_upload.argtypes
        L.bsb200_batch_fetch_dense_dev.argtypes = [_P, _P, _P, _P, ctypes.c_uint64, _P, _P, _P]
        L.bsb200_pack_pairs.restype = ctypes.c_uint64
        L.bsb200_pack_pairs.argtypes = [_P, _P, _P, _P, _P, _P, ctypes.c_uint64, _P, _P, _P, ctypes.c_int]
        L.bsb200_scatter_words.argtypes = [_P, _P, _P, _P, _P, ctypes.c_uint64, ctypes.c_int]
        L.bsb200_trim.argtypes = [_P]
        L.bsb200_default_context.restype = _P
        L.bsb200_epi8_bandwidth.restype = ctypes.c_uint32
        L.bsb200_epi8_bandwidth.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.bsb200_edit_bandwidth.restype = ctypes.c_uint32
        L.bsb200_edit_bandwidth.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32]
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return a.ctypes.data_as(_P)


def cigar_offsets(batch):
    """Per-pair capacity qlen+tlen+2 words (an alignment never has more runs than columns)."""
    cap = batch.qlen.astype(np.uint64) + batch.tlen.astype(np.uint64) + np.uint64(2)
    off = np.zeros(batch.n + 1, dtype=np.uint64)
    np.cumsum(cap, out=off[1:])
    return off


class Context:
    """One CUDA device + stream (bsb200_ctx)."""

    def __init__(self, device=0, trace_budget_bytes=0):
        self._lib = lib()
        self._h = self._lib.bsb200_create(int(device), int(trace_budget_bytes))
        if not self._h:
            raise RuntimeError("bsalign_b200: cannot create a context on CUDA device %d (no GPU?); there is no CPU fallback" % device)
        self.device = device

    def close(self):
        if self._h:
            self._lib.bsb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self._lib.bsb200_last_error(self._h).decode()))

    def timing(self):
        t = Timing()
        self._lib.bsb200_get_timing(self._h, ctypes.byref(t))
        return t.as_dict()

    # ---- one-shot batch calls (host buffers in, host buffers out) ---------------------------------
    def epi8_batch(self, batch, mode, bandwidth, matrix, gapo1, gape1, gapo2=0, gape2=0, want_cigar=True, out=None, dense=False):
        if dense:
            return self._staged("epi8", batch, mode, bandwidth, matrix, (gapo1, gape1, gapo2, gape2), out)
        n = batch.n
        m = np.ascontiguousarray(matrix, dtype=np.int8)
        res, cg, off, ncg, st = out if out is not None else _alloc_out(batch, want_cigar)
        rc = self._lib.bsb200_epi8_pairwise_batch(self._h, n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                                  int(mode), int(bandwidth), _ptr(m), gapo1, gape1, gapo2, gape2,
                                                  _ptr(res), _ptr(cg), _ptr(off), _ptr(ncg), _ptr(st))
        self._check(rc, "bsb200_epi8_pairwise_batch")
        return BatchResult(res, cg, off, ncg, st)

    def _staged(self, kind, batch, mode, bandwidth, matrix, gaps, out):
        """ONE C-ABI call (bsb200_pairwise_batch_dense): host buffers in, results + dense pair-ordered cigars out."""
        m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
        res, cg, off, ncg, st = out if out is not None else _alloc_out(batch, True)
        total = ctypes.c_uint64(0)
        rc = self._lib.bsb200_pairwise_batch_dense(self._h, 0 if kind == "epi8" else 1, batch.n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen),
                                                   _ptr(batch.toff), _ptr(batch.tlen), int(mode), int(bandwidth), _ptr(m), gaps[0], gaps[1], gaps[2], gaps[3],
                                                   _ptr(res), _ptr(cg), 0 if cg is None else cg.size, ctypes.byref(total), _ptr(ncg), _ptr(st))
        self._check(rc, "bsb200_pairwise_batch_dense")
        self.last_timing = self.timing()
        return BatchResult(res, cg, None, ncg, st)

    def dense_bits(self, kind, bits, batch, mode, bandwidth, matrix=None, gaps=(0, 0, 0, 0), out=None):
        """ONE C-ABI call (bsb200_pairwise_batch_dense_bits) with the sequences 2-bit packed in BaseBank words (pack_bits); batch.qoff / toff are
        base offsets.  Results + dense pair-ordered cigars."""
        m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
        res, cg, off, ncg, st = out if out is not None else _alloc_out(batch, True)
        total = ctypes.c_uint64(0)
        rc = self._lib.bsb200_pairwise_batch_dense_bits(self._h, 0 if kind == "epi8" else 1, batch.n, _ptr(bits), _ptr(batch.qoff), _ptr(batch.qlen),
                                                        _ptr(batch.toff), _ptr(batch.tlen), int(mode), int(bandwidth), _ptr(m), gaps[0], gaps[1], gaps[2], gaps[3],
                                                        _ptr(res), _ptr(cg), 0 if cg is None else cg.size, ctypes.byref(total), _ptr(ncg), _ptr(st))
        self._check(rc, "bsb200_pairwise_batch_dense_bits")
        self.last_timing = self.timing()
        return BatchResult(res, cg, None, ncg, st)

    def edit_batch(self, batch, mode, bandwidth, want_cigar=True, out=None, dense=False):
        if dense:
            return self._staged("edit", batch, mode, bandwidth, None, (0, 0, 0, 0), out)
        n = batch.n
        res, cg, off, ncg, st = out if out is not None else _alloc_out(batch, want_cigar)
        rc = self._lib.bsb200_edit_pairwise_batch(self._h, n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                                  int(mode), int(bandwidth), _ptr(res), _ptr(cg), _ptr(off), _ptr(ncg), _ptr(st))
        self._check(rc, "bsb200_edit_pairwise_batch")
        return BatchResult(res, cg, off, ncg, st)

    def kmer_edit_batch(self, batch, ksz, dense=False, out=None):
        """kmer_striped_seqedit_pairwise (bsalign.h:1209) over a batch; dense=True: cigars come back dense and in pair order."""
        n = batch.n
        if not dense:
            res, cg, off, ncg, st = out if out is not None else _alloc_out(batch, True)
            rc = self._lib.bsb200_kmer_edit_batch(self._h, n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                                  int(ksz), _ptr(res), _ptr(cg), _ptr(off), _ptr(ncg), _ptr(st))
            self._check(rc, "bsb200_kmer_edit_batch")
            return BatchResult(res, cg, off, ncg, st)
        res, cg, _off, ncg, st = out if out is not None else _alloc_out(batch, True)
        total = ctypes.c_uint64(0)
        rc = self._lib.bsb200_kmer_edit_batch_dense(self._h, n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                                    int(ksz), _ptr(res), _ptr(cg), len(cg), ctypes.byref(total), _ptr(ncg), _ptr(st))
        self._check(rc, "bsb200_kmer_edit_batch_dense")
        return BatchResult(res, cg, None, ncg, st)

    # ---- staged form: inputs stay resident in HBM between runs -------------------------------------
    def upload(self, kind, batch, mode, bandwidth, matrix=None, gaps=(0, 0, 0, 0), want_cigar=True):
        m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
        h = self._lib.bsb200_batch_upload(self._h, 0 if kind == "epi8" else 1, batch.n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen),
                                          _ptr(batch.toff), _ptr(batch.tlen), int(mode), int(bandwidth), _ptr(m),
                                          gaps[0], gaps[1], gaps[2], gaps[3], 1 if want_cigar else 0)
        if not h:
            raise RuntimeError("bsb200_batch_upload failed: %s" % self._lib.bsb200_last_error(self._h).decode())
        return ResidentBatch(self, h, batch, want_cigar)


    def upload_dev(self, kind, d_seqs_ptr, batch, mode, bandwidth, matrix=None, gaps=(0, 0, 0, 0), want_cigar=True):
        """Like upload(), but the sequence arena is already in this device's memory (int device pointer; the caller keeps it alive
        until the batch is freed); batch.qoff/qlen/toff/tlen are host arrays that index it (batch.seqs is not read)."""
        m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
        h = self._lib.bsb200_batch_upload_dev(self._h, 0 if kind == "epi8" else 1, batch.n, ctypes.c_void_p(int(d_seqs_ptr)), _ptr(batch.qoff), _ptr(batch.qlen),
                                              _ptr(batch.toff), _ptr(batch.tlen), int(mode), int(bandwidth), _ptr(m),
                                              gaps[0], gaps[1], gaps[2], gaps[3], 1 if want_cigar else 0)
        if not h:
            raise RuntimeError("bsb200_batch_upload_dev failed: %s" % self._lib.bsb200_last_error(self._h).decode())
        return ResidentBatch(self, h, batch, want_cigar)

    def upload_bits(self, kind, bits, batch, mode, bandwidth, matrix=None, gaps=(0, 0, 0, 0), want_cigar=True):
        """Like upload(), with the sequences 2-bit packed in the reference's BaseBank word layout (pack_bits); batch.qoff / toff are base offsets."""
        m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
        h = self._lib.bsb200_batch_upload_bits(self._h, 0 if kind == "epi8" else 1, batch.n, _ptr(bits), _ptr(batch.qoff), _ptr(batch.qlen),
                                               _ptr(batch.toff), _ptr(batch.tlen), int(mode), int(bandwidth), _ptr(m),
                                               gaps[0], gaps[1], gaps[2], gaps[3], 1 if want_cigar else 0)
        if not h:
            raise RuntimeError("bsb200_batch_upload_bits failed: %s" % self._lib.bsb200_last_error(self._h).decode())
        return ResidentBatch(self, h, batch, want_cigar)

    def trim(self):
        self._lib.bsb200_trim(self._h)


def pack_pairs(batch, idx, out_seqs=None, nthreads=8):
    """Compact arena of the pairs idx of a batch (query then target of each pair, in idx order).  Returns (PairBatch over the compact
    arena, bytes).  out_seqs: optional uint8 array (e.g. pinned) of sufficient size to pack into."""
    L = lib()
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    m = len(idx)
    qoff = np.zeros(m, dtype=np.uint64)
    toff = np.zeros(m, dtype=np.uint64)
    nbytes = int(batch.qlen[idx.astype(np.int64)].astype(np.int64).sum() + batch.tlen[idx.astype(np.int64)].astype(np.int64).sum()) if m else 0
    if out_seqs is None:
        out_seqs = np.empty(max(nbytes, 1), dtype=np.uint8)
    assert out_seqs.size >= nbytes
    L.bsb200_pack_pairs(_ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen), _ptr(idx), m,
                        _ptr(out_seqs), _ptr(qoff), _ptr(toff), int(nthreads))
    i64 = idx.astype(np.int64)
    pb = PairBatch.__new__(PairBatch)
    pb.seqs, pb.qoff, pb.qlen, pb.toff, pb.tlen = out_seqs[:max(nbytes, 1)], qoff, np.ascontiguousarray(batch.qlen[i64]), toff, np.ascontiguousarray(batch.tlen[i64])
    return pb, nbytes


class RemsaBatch:
    """The arenas of bsb200_remsa_batch for a list of jobs (mlen, bw, mbeg, mend, rdlen, seqs0, seqs1 of sz1 bytes with bw / 2 bytes of
    padding in front, mats (2, 4, sz1)): the ten arrays of a job back to back, the reference's own layout (bspoa.h:4209-4229)."""

    def __init__(self, jobs):
        n = len(jobs)
        self.jobs, self.n = jobs, n
        self.hdr = np.zeros((n, 8), dtype=np.int32)
        self.in_off, self.match_off, self.mat_off = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        o = mo = xo = 0
        for k, j in enumerate(jobs):
            self.hdr[k, :5] = (j.mlen, j.bw, j.mbeg, j.mend, j.rdlen)
            self.in_off[k] = o; o += (10 * len(j.seqs0) + 15) // 16 * 16
            self.match_off[k] = mo; mo += j.rdlen
            self.mat_off[k] = xo; xo += 2 * (2 * j.mlen + 1) * (j.bw + 2)
        self.in_bytes, self.match_ints, self.mat_bytes = o, mo, xo
        self.arena = np.zeros(max(o, 1), dtype=np.uint8)
        for k, j in enumerate(jobs):
            b0, sz1 = int(self.in_off[k]), len(j.seqs0)
            self.arena[b0:b0 + sz1] = j.seqs0
            self.arena[b0 + sz1:b0 + 2 * sz1] = j.seqs1
            self.arena[b0 + 2 * sz1:b0 + 10 * sz1] = j.mats.reshape(-1)


def remsa_batch(ctx, jobs, want_matrices=False, out=None):
    """bsb200_remsa_batch: the DP + walk of remsa_pedit_rd_bspoacore (bspoa.h:3916) for a batch of jobs (a list, or a RemsaBatch).
    Returns (list of match arrays, out[n, 4] = score / status / matched / 0, list of (M0, M1) or None).  out: optional (match, out) arrays."""
    rb = jobs if isinstance(jobs, RemsaBatch) else RemsaBatch(jobs)
    n = rb.n
    match, res = out if out is not None else (np.zeros(max(rb.match_ints, 1), dtype=np.int32), np.zeros((n, 4), dtype=np.int32))
    mats = np.zeros(max(rb.mat_bytes, 1), dtype=np.uint8) if want_matrices else None
    rc = ctx._lib.bsb200_remsa_batch(ctx._h, n, _ptr(rb.hdr), _ptr(rb.arena), _ptr(rb.in_off), int(rb.in_bytes), _ptr(match), _ptr(rb.match_off),
                                     int(rb.match_ints), _ptr(res), _ptr(mats), _ptr(rb.mat_off) if want_matrices else None)
    ctx._check(rc, "bsb200_remsa_batch")
    ms = [match[int(rb.match_off[k]):int(rb.match_off[k]) + rb.jobs[k].rdlen] for k in range(n)]
    mm = None
    if want_matrices:
        mm = []
        for k, j in enumerate(rb.jobs):
            szm = (2 * j.mlen + 1) * (j.bw + 2)
            b0 = int(rb.mat_off[k])
            mm.append((mats[b0:b0 + szm], mats[b0 + szm:b0 + 2 * szm]))
    return ms, res, mm


def pack_pairs_dev(ctx, d_src_ptr, batch, idx, d_dst_ptr):
    """bsb200_pack_pairs_dev: the batch's whole arena is at device pointer d_src_ptr; the pairs idx (query then target of each pair, in idx order)
    are gathered into the compact arena at d_dst_ptr by a kernel.  batch.qoff / qlen / toff / tlen are host arrays."""
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    rc = ctx._lib.bsb200_pack_pairs_dev(ctx._h, ctypes.c_void_p(int(d_src_ptr)), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                        _ptr(idx), len(idx), ctypes.c_void_p(int(d_dst_ptr)))
    ctx._check(rc, "bsb200_pack_pairs_dev")


def pairwise_batch_ptrs(ctx, kind, queries, targets, mode, bandwidth, matrix=None, gaps=(0, 0, 0, 0), nthreads=4):
    """bsb200_pairwise_batch_ptrs: one numpy array per sequence (one pointer each, like the reference's callers).  Returns (results, list of cigars, status)."""
    L = lib()
    n = len(queries)
    qs = [np.ascontiguousarray(x, dtype=np.uint8) for x in queries]
    ts = [np.ascontiguousarray(x, dtype=np.uint8) for x in targets]
    qp = (ctypes.c_void_p * n)(*[x.ctypes.data for x in qs])
    tp = (ctypes.c_void_p * n)(*[x.ctypes.data for x in ts])
    ql = np.array([len(x) for x in qs], np.uint32); tl = np.array([len(x) for x in ts], np.uint32)
    cgs = [np.zeros(int(a) + int(b) + 2, np.uint32) for a, b in zip(ql, tl)]
    cp = (ctypes.c_void_p * n)(*[x.ctypes.data for x in cgs])
    res = np.zeros((n, 10), np.int32); ncg = np.zeros(n, np.uint32); st = np.zeros(n, np.int32)
    m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
    rc = L.bsb200_pairwise_batch_ptrs(ctx._h, 0 if kind == "epi8" else 1, n, qp, _ptr(ql), tp, _ptr(tl), int(mode), int(bandwidth), _ptr(m),
                                      gaps[0], gaps[1], gaps[2], gaps[3], _ptr(res), cp, _ptr(ncg), _ptr(st), int(nthreads))
    ctx._check(rc, "bsb200_pairwise_batch_ptrs")
    return res, [c[:int(k)] for c, k in zip(cgs, ncg)], st


def pairwise_batch_multi(ctxs, kind, batch, mode, bandwidth, matrix=None, gaps=(0, 0, 0, 0)):
    """bsb200_pairwise_batch_multi: every context's GPU from this one process."""
    L = lib()
    hs = (ctypes.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    m = np.ascontiguousarray(matrix if matrix is not None else np.zeros(16), dtype=np.int8)
    res, cg, off, ncg, st = _alloc_out(batch, True)
    rc = L.bsb200_pairwise_batch_multi(hs, len(ctxs), 0 if kind == "epi8" else 1, batch.n, _ptr(batch.seqs), _ptr(batch.qoff), _ptr(batch.qlen), _ptr(batch.toff), _ptr(batch.tlen),
                                       int(mode), int(bandwidth), _ptr(m), gaps[0], gaps[1], gaps[2], gaps[3], _ptr(res), _ptr(cg), _ptr(off), _ptr(ncg), _ptr(st))
    if rc != 0:
        raise RuntimeError("bsb200_pairwise_batch_multi failed: " + "; ".join(L.bsb200_last_error(c._h).decode() for c in ctxs))
    return BatchResult(res, cg, off, ncg, st)


class SeqFile:
    """bsb200_seqfile_read: FASTA / FASTQ (plain or .gz) into the reference's BaseBank words (readseq_filereader + seq2basebank)."""

    def __init__(self, path):
        L = lib()
        h = L.bsb200_seqfile_read(os.fsencode(path))
        err = L.bsb200_seqfile_error(h).decode()
        if err:
            L.bsb200_seqfile_free(h)
            raise IOError(err)
        n = int(L.bsb200_seqfile_nseq(h))
        self.nbases = int(L.bsb200_seqfile_nbases(h))
        nw = (self.nbases + 31) // 32 + 1
        self.bits = np.ctypeslib.as_array(ctypes.cast(L.bsb200_seqfile_bits(h), ctypes.POINTER(ctypes.c_uint64)), shape=(nw,)).copy()
        self.off = np.ctypeslib.as_array(ctypes.cast(L.bsb200_seqfile_offsets(h), ctypes.POINTER(ctypes.c_uint64)), shape=(max(n, 1),))[:n].copy()
        self.len = np.ctypeslib.as_array(ctypes.cast(L.bsb200_seqfile_lengths(h), ctypes.POINTER(ctypes.c_uint32)), shape=(max(n, 1),))[:n].copy()
        self.names = [L.bsb200_seqfile_name(h, i).decode() for i in range(n)]
        L.bsb200_seqfile_free(h)

    def bases(self, i):
        idx = np.arange(int(self.off[i]), int(self.off[i]) + int(self.len[i]), dtype=np.uint64)
        return ((self.bits[(idx >> np.uint64(5)).astype(np.int64)] >> (((~idx) & np.uint64(31)) << np.uint64(1))) & np.uint64(3)).astype(np.uint8)

    def pairs(self):
        """Consecutive records as a PairBatch whose offsets are BASE offsets into self.bits (for Context.upload_bits)."""
        n = len(self.len) // 2
        pb = PairBatch.__new__(PairBatch)
        pb.seqs = None
        pb.qoff, pb.toff = np.ascontiguousarray(self.off[0:2 * n:2]), np.ascontiguousarray(self.off[1:2 * n:2])
        pb.qlen, pb.tlen = np.ascontiguousarray(self.len[0:2 * n:2]), np.ascontiguousarray(self.len[1:2 * n:2])
        return pb


def format_pair_text(sf, k, result, cigar):
    """The text `bsalign align` / `bsalign edit` print for pair k of a SeqFile (records 2k, 2k+1)."""
    L = lib()
    rs = np.ascontiguousarray(result, dtype=np.int32)
    cg = np.ascontiguousarray(cigar, dtype=np.uint32)
    args = (sf.names[2 * k].encode(), int(sf.len[2 * k]), sf.names[2 * k + 1].encode(), int(sf.len[2 * k + 1]), _ptr(rs), _ptr(sf.bits),
            int(sf.off[2 * k]), int(sf.off[2 * k + 1]), _ptr(cg), len(cg))
    need = int(L.bsb200_format_pair_text(None, 0, *args))
    buf = ctypes.create_string_buffer(need + 1)
    L.bsb200_format_pair_text(buf, need, *args)
    return buf.raw[:need]


_libc = ctypes.CDLL(None)
_libc.fopen.restype = ctypes.c_void_p
_libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
_libc.fclose.argtypes = [ctypes.c_void_p]


def read_binary_msa(path):
    """All MSAs of a file in the reference's binary MSA format (bspoa.h:1555-1643) as dicts: nseq, mlen, cols (mlen, nseq + 1), qlt, alt, meta."""
    L = lib()
    fp = _libc.fopen(os.fsencode(path), b"rb")
    if not fp:
        raise IOError("cannot open %s" % path)
    out = []
    try:
        while True:
            m = L.bsb200_msa_read(fp)
            if not m:
                break
            nseq, mlen = int(L.bsb200_msa_nseq(m)), int(L.bsb200_msa_mlen(m))
            grab = lambda p, n: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(max(n, 1),))[:n].copy()
            ln = ctypes.c_uint32(0)
            mp = L.bsb200_msa_meta(m, ctypes.byref(ln))
            out.append(dict(nseq=nseq, mlen=mlen, cols=grab(L.bsb200_msa_cols(m), mlen * (nseq + 1)).reshape(mlen, nseq + 1), qlt=grab(L.bsb200_msa_qlt(m), mlen),
                            alt=grab(L.bsb200_msa_alt(m), mlen), meta=ctypes.string_at(mp, ln.value) if ln.value else b""))
            L.bsb200_msa_free(m)
    finally:
        _libc.fclose(fp)
    return out


def write_binary_msa(path, msas, mode="wb"):
    L = lib()
    fp = _libc.fopen(os.fsencode(path), mode.encode())
    if not fp:
        raise IOError("cannot open %s" % path)
    try:
        for m in msas:
            cols = np.ascontiguousarray(m["cols"], dtype=np.uint8)
            rc = L.bsb200_msa_write(fp, int(m["nseq"]), int(m["mlen"]), _ptr(cols), _ptr(np.ascontiguousarray(m["qlt"], np.uint8)),
                                    _ptr(np.ascontiguousarray(m["alt"], np.uint8)), m.get("meta") or None, len(m.get("meta") or b""))
            if rc:
                raise IOError("bsb200_msa_write failed")
    finally:
        _libc.fclose(fp)


def pack_bits(seqs):
    """uint8 bases 0..3 -> the reference's BaseBank words (dna.h:63): base i sits in word i >> 5 at bit ((~i) & 31) << 1."""
    n = len(seqs)
    pad = (-n) % 32
    s = np.concatenate([np.asarray(seqs, np.uint8), np.zeros(pad, np.uint8)]).reshape(-1, 32).astype(np.uint64)
    sh = ((31 - np.arange(32)) * 2).astype(np.uint64)
    return np.bitwise_or.reduce(s << sh, axis=1).astype(np.uint64)


def scatter_words(dst, dst_off, src, src_off, length, nthreads=8):
    lib().bsb200_scatter_words(_ptr(dst), _ptr(np.ascontiguousarray(dst_off, dtype=np.uint64)), _ptr(src), _ptr(np.ascontiguousarray(src_off, dtype=np.uint64)),
                               _ptr(np.ascontiguousarray(length, dtype=np.uint32)), len(length), int(nthreads))


def _alloc_out(batch, want_cigar):
    n = batch.n
    res = np.zeros((n, 10), dtype=np.int32)
    off = cigar_offsets(batch) if want_cigar else None
    cg = np.zeros(int(off[-1]) if want_cigar else 0, dtype=np.uint32) if want_cigar else None
    ncg = np.zeros(n, dtype=np.uint32)
    st = np.zeros(n, dtype=np.int32)
    return res, cg, off, ncg, st


class BatchResult:
    def __init__(self, results, cigar_arena, cigar_off, ncigar, status):
        self.results = results      # (n, 10) int32, columns = RESULT_FIELDS
        self.cigar_arena = cigar_arena
        self._off = cigar_off       # None: dense pair-ordered cigars, the offsets are the running sum of ncigar (computed on first use)
        self.ncigar = ncigar
        self.status = status

    @property
    def cigar_off(self):
        if self._off is None:
            self._off = np.zeros(len(self.ncigar) + 1, dtype=np.uint64)
            np.cumsum(self.ncigar, out=self._off[1:])
        return self._off

    def cigar(self, i):
        o = int(self.cigar_off[i])
        return self.cigar_arena[o:o + int(self.ncigar[i])]

    def cigars(self):
        return [self.cigar(i) for i in range(len(self.ncigar))]


class ResidentBatch:
    def __init__(self, ctx, handle, batch, want_cigar):
        self.ctx, self._h, self.batch, self.want_cigar = ctx, handle, batch, want_cigar

    def run(self):
        self.ctx._check(self.ctx._lib.bsb200_batch_run(self.ctx._h, self._h), "bsb200_batch_run")

    def fetch(self, out=None):
        res, cg, off, ncg, st = out if out is not None else _alloc_out(self.batch, self.want_cigar)
        rc = self.ctx._lib.bsb200_batch_fetch(self.ctx._h, self._h, _ptr(res), _ptr(cg), _ptr(off), _ptr(ncg), _ptr(st))
        self.ctx._check(rc, "bsb200_batch_fetch")
        return BatchResult(res, cg, off, ncg, st)

    def fetch_dense(self, out=None):
        """Cigars dense and in pair order (pair i starts at sum(ncigar[:i])): the fast path, no host-side scatter."""
        res, cg, off, ncg, st = out if out is not None else _alloc_out(self.batch, self.want_cigar)
        total = ctypes.c_uint64(0)
        rc = self.ctx._lib.bsb200_batch_fetch_dense(self.ctx._h, self._h, _ptr(res), _ptr(cg), 0 if cg is None else cg.size, ctypes.byref(total), _ptr(ncg), _ptr(st))
        self.ctx._check(rc, "bsb200_batch_fetch_dense")
        return BatchResult(res, cg, None, ncg, st)

    def fetch_dense_dev(self, d_results, d_cigars, cigar_cap_words, d_ncigar, d_status):
        """Results into DEVICE buffers (int device pointers, any may be 0/None).  Returns the number of dense cigar words."""
        total = ctypes.c_uint64(0)
        vp = lambda x: ctypes.c_void_p(int(x)) if x else None
        rc = self.ctx._lib.bsb200_batch_fetch_dense_dev(self.ctx._h, self._h, vp(d_results), vp(d_cigars), int(cigar_cap_words), ctypes.byref(total), vp(d_ncigar), vp(d_status))
        if rc != 0 and not (d_cigars and total.value > cigar_cap_words):
            self.ctx._check(rc, "bsb200_batch_fetch_dense_dev")
        return int(total.value), rc

    def free(self):
        if self._h:
            self.ctx._lib.bsb200_batch_free(self.ctx._h, self._h)
            self._h = None


# ---- reference-named single-pair functions --------------------------------------------------------
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def banded_striped_epi8_seqalign_set_score_matrix(mat, mis):
    """bsalign.h:323: matrix[i] = mis if the two 2-bit codes of i differ else mat."""
    return np.array([mis if ((i ^ (i >> 2)) & 3) else mat for i in range(16)], dtype=np.int8)


def _single(batchfn, qseq, tseq, *args):
    b = PairBatch.from_lists([(qseq, tseq)])
    r = batchfn(b, *args)
    rs = dict(zip(RESULT_FIELDS, (int(v) for v in r.results[0])))
    return rs, r.cigar(0).copy(), int(r.status[0])


def banded_striped_epi8_seqalign_pairwise(qseq, tseq, mode, bandwidth, matrix, gapo1, gape1, gapo2, gape2, ctx=None):
    """Same arguments as bsalign.h:399 without mempool/verbose.  Returns (result dict, cigar words, status)."""
    ctx = ctx or default_context()
    return _single(ctx.epi8_batch, qseq, tseq, mode, bandwidth, matrix, gapo1, gape1, gapo2, gape2)


def striped_seqedit_pairwise(qseq, tseq, mode, bandwidth, ctx=None):
    """Same arguments as bsalign.h:232 without mempool/verbose."""
    ctx = ctx or default_context()
    return _single(ctx.edit_batch, qseq, tseq, mode, bandwidth)


def seqalign_cigar2alnstr(qseq, tseq, rs, cigars):
    """bsalign.h:531-582: three strings (query row, target row, match row) of the alignment."""
    x, y = rs["qb"], rs["tb"]
    a, b, c = [], [], []
    for w in cigars:
        op, ln = int(w) & 0xF, int(w) >> 4
        for _ in range(ln):
            if op == 0:
                qa, ta = "ACGTN"[qseq[x]], "ACGTN"[tseq[y]]
                a.append(qa); b.append(ta); c.append("|" if qseq[x] == tseq[y] else "*")
                x += 1; y += 1
            elif op == 1:
                a.append("ACGTN"[qseq[x]]); b.append("-"); c.append("-"); x += 1
            elif op == 2:
                a.append("-"); b.append("ACGTN"[tseq[y]]); c.append("-"); y += 1
    return "".join(a), "".join(b), "".join(c)
