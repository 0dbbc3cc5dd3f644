"""Records of remsa_pedit_rd_bspoacore (bspoa.h:3916) captured from the unmodified reference (oracle/ref_remsa_harness.c ->
oracle/_ref/libbsref_remsa.so) and the checks against them: test infrastructure."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REMSA_SO = os.path.join(ROOT, "oracle", "_ref", "libbsref_remsa.so")
GOLD = os.path.join(HERE, "golden", "remsa_golden.npz")


class RemsaJob:
    """One call: the read in MSA coordinates (seqs0) against the reversed consensus (seqs1) and the profile scores (mats)."""
    pass


def have_ref():
    return os.path.exists(REMSA_SO)


def _a8(x):
    return (x + 7) // 8 * 8


def parse(blob, ncall):
    jobs, o = [], 0
    for _ in range(ncall):
        hdr = blob[o:o + 64].view(np.int32); o += 64
        j = RemsaJob()
        magic, j.rid, j.mlen, j.bw, j.mbeg, j.mend, j.rdlen, j.nev, j.sz1, j.szm = [int(v) for v in hdr[:10]]
        assert magic == 0x52454d53
        j.seqs0 = blob[o:o + j.sz1].copy(); o += _a8(j.sz1)
        j.seqs1 = blob[o:o + j.sz1].copy(); o += _a8(j.sz1)
        j.mats = np.zeros((2, 4, j.sz1), np.uint8)
        for a in range(2):
            for b in range(4):
                j.mats[a, b] = blob[o:o + j.sz1]; o += _a8(j.sz1)
        j.M0 = blob[o:o + j.szm].copy(); o += _a8(j.szm)
        j.M1 = blob[o:o + j.szm].copy(); o += _a8(j.szm)
        j.match = blob[o:o + 4 * j.rdlen].view(np.int32).copy(); o += _a8(4 * j.rdlen)
        jobs.append(j)
    assert o == len(blob)
    return jobs


def reference_dump(reads, realn=1, editbw=0):
    """Run one BSPOA job through the reference's end_bspoa and return the records of every remsa_pedit_rd_bspoacore call."""
    L = ctypes.CDLL(REMSA_SO)
    L.bsref_remsa_dump.restype = ctypes.c_int64
    seqs = np.concatenate(reads).astype(np.uint8)
    ln = np.array([len(r) for r in reads], np.uint32)
    off = np.concatenate([[0], np.cumsum(ln[:-1])]).astype(np.uint64)
    out, nc = ctypes.c_void_p(), ctypes.c_uint32()
    n = L.bsref_remsa_dump(ctypes.c_uint32(len(reads)), seqs.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p), ln.ctypes.data_as(ctypes.c_void_p),
                           ctypes.c_int(realn), ctypes.c_int(editbw), ctypes.byref(out), ctypes.byref(nc))
    blob = np.frombuffer(ctypes.string_at(out, n), dtype=np.uint8)
    L.bsref_remsa_free(out)
    return parse(blob, nc.value)


def oracle_core(j):
    """bso_remsa_core on one record -> (M0, M1, match, scr, err)"""
    import checkers as ck
    L = ck.oracle()
    hw = j.bw // 2
    P = lambda a, o=0: ctypes.c_void_p(a.ctypes.data + o)
    M0 = np.zeros(j.szm, np.uint8); M1 = np.zeros(j.szm, np.uint8)
    match = np.zeros(max(j.rdlen, 1), np.int32)
    scr = ctypes.c_int32(0)
    err = L.bso_remsa_core(j.mlen, j.bw, j.mbeg, j.mend, j.rdlen, P(j.seqs0, hw), P(j.seqs1, hw),
                           *[P(j.mats[a, b], hw) for a in range(2) for b in range(4)], P(M0), P(M1), P(match), ctypes.byref(scr))
    return M0, M1, match[:j.rdlen], scr.value, err


def compare(j, M0, M1, match):
    """the rows of this call's diagonals (the reference's matrices keep older calls' rows elsewhere) and the matched columns"""
    rl = j.bw + 2
    lo, hi = rl * 2 * j.mbeg, rl * 2 * j.mend
    if not np.array_equal(M0[lo:hi], j.M0[lo:hi]):
        return "matrix 0 differs"
    if not np.array_equal(M1[lo:hi], j.M1[lo:hi]):
        return "matrix 1 differs"
    if not np.array_equal(match, j.match):
        return "matched columns differ at %s" % np.nonzero(match != j.match)[0][:5]
    return None


def save_golden(jobs, path=GOLD):
    z = {"n": np.array([len(jobs)])}
    for k, j in enumerate(jobs):
        z["h%d" % k] = np.array([j.rid, j.mlen, j.bw, j.mbeg, j.mend, j.rdlen, j.nev, j.sz1, j.szm], np.int32)
        z["s0_%d" % k] = j.seqs0; z["s1_%d" % k] = j.seqs1; z["m_%d" % k] = j.mats; z["match_%d" % k] = j.match
        rl = j.bw + 2
        lo, hi = rl * 2 * j.mbeg, rl * 2 * j.mend
        z["M0_%d" % k] = j.M0[lo:hi]; z["M1_%d" % k] = j.M1[lo:hi]
    np.savez_compressed(path, **z)


def load_golden(path=GOLD):
    z = np.load(path)
    jobs = []
    for k in range(int(z["n"][0])):
        j = RemsaJob()
        j.rid, j.mlen, j.bw, j.mbeg, j.mend, j.rdlen, j.nev, j.sz1, j.szm = [int(v) for v in z["h%d" % k]]
        j.seqs0 = z["s0_%d" % k]; j.seqs1 = z["s1_%d" % k]; j.mats = z["m_%d" % k]; j.match = z["match_%d" % k]
        rl = j.bw + 2
        lo, hi = rl * 2 * j.mbeg, rl * 2 * j.mend
        j.M0 = np.zeros(j.szm, np.uint8); j.M1 = np.zeros(j.szm, np.uint8)
        j.M0[lo:hi] = z["M0_%d" % k]; j.M1[lo:hi] = z["M1_%d" % k]
        jobs.append(j)
    return jobs
