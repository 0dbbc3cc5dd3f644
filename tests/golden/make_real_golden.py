"""Golden answers of the UNMODIFIED reference (oracle/_ref/libbsref.so) for ALL 12,477 ONT read pairs of the reference's own
example/real.ont.b10M.txt under the three configurations of example/run.sh:5-9.  Run in the build container only:

    python tests/golden/make_real_golden.py

  NoBand : bsalign align -M 2 -X 2 -O 4 -E 2 -Q 0 -P 0            (mode overlap - main.c:262 - band = roundup16(qlen), main.c:315)
  Band64 : bsalign align -W 64 -M 2 -X 2 -O 4 -E 2 -Q 0 -P 0 -m overlap
  Edit0  : bsalign edit -W 0                                       (mode global, main.c:142)

Output (committed): tests/golden/real_ont.npz - the sequences 2-bit packed (4 bases per byte), per configuration the ten result ints of
every pair, its number of cigar words and two position-weighted 64-bit checksums of the words (the words themselves would be 30 MB).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck
from bsalign_b200 import synth

CONFIGS = [("NoBand", "epi8", 1, 0), ("Band64", "epi8", 1, 64), ("Edit0", "edit", 0, 0)]
MATRIX, GAPS = (2, -2), (-4, -2, 0, 0)


def read_pairs(path):
    seqs = []
    cur = []
    for line in open(path):
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(line.strip())
    if cur:
        seqs.append("".join(cur))
    tab = np.full(256, 255, np.uint8)
    for i, c in enumerate("ACGT"):
        tab[ord(c)] = i; tab[ord(c.lower())] = i
    codes = [tab[np.frombuffer(s.encode(), np.uint8)] for s in seqs]
    assert all((c < 4).all() for c in codes)
    return [(codes[i], codes[i + 1]) for i in range(0, len(codes) - 1, 2)]


def cigar_checksums(arena, off, ncg):
    """Two position-weighted sums (mod 2^64) of every pair's words; vectorised over the concatenated words."""
    n = len(ncg)
    tot = int(ncg.sum())
    words = np.concatenate([arena[int(off[i]):int(off[i]) + int(ncg[i])] for i in range(n)]).astype(np.uint64) if tot else np.zeros(0, np.uint64)
    return checksums_dense(words, ncg)


def checksums_dense(words, ncg):
    n = len(ncg)
    start = np.zeros(n + 1, np.int64); np.cumsum(ncg.astype(np.int64), out=start[1:])
    local = np.arange(len(words), dtype=np.uint64) - np.repeat(start[:-1].astype(np.uint64), ncg.astype(np.int64))
    with np.errstate(over="ignore"):
        a = (words.astype(np.uint64) + np.uint64(1)) * (local * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD1B54A32D192ED03))
        b = (words.astype(np.uint64) ^ np.uint64(0xA5A5A5A5)) * ((local + np.uint64(7)) * (local + np.uint64(13)) * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(1))
        h1 = np.zeros(n, np.uint64); h2 = np.zeros(n, np.uint64)
        nz = np.nonzero(ncg)[0]
        if len(nz):
            h1[nz] = np.add.reduceat(a, start[nz])
            h2[nz] = np.add.reduceat(b, start[nz])
    return h1, h2


def main():
    assert ck.have_ref(), "oracle/_ref/libbsref.so missing: run `make -C oracle` where /root/reference exists"
    pairs = read_pairs("/root/reference/example/real.ont.b10M.txt")
    batch = synth.PairBatch.from_lists(pairs)
    print("pairs", batch.n, "bases", batch.seqs.size, flush=True)
    pad = (-batch.seqs.size) % 4
    s = np.concatenate([batch.seqs, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    packed = (s[:, 0] | (s[:, 1] << 2) | (s[:, 2] << 4) | (s[:, 3] << 6)).astype(np.uint8)
    out = {"packed": packed, "nbases": np.int64(batch.seqs.size), "qoff": batch.qoff, "qlen": batch.qlen, "toff": batch.toff, "tlen": batch.tlen,
           "configs": np.array([[k == "edit", mode, bw] for _, k, mode, bw in CONFIGS], dtype=np.int32), "matrix": np.array(MATRIX, np.int32), "gaps": np.array(GAPS, np.int32)}
    mtx = synth.score_matrix(*MATRIX)
    for ci, (name, kind, mode, bw) in enumerate(CONFIGS):
        errs = np.zeros(batch.n, np.int32)
        ck.oracle_batch(kind, batch, mode, bw, mtx, GAPS, nthreads=8, errs=errs, want_cigar=False)
        assert not errs.any(), "the reference's own example hits one of its undefined-behaviour paths?"
        res, cigs, _ = ck.ref_batch(kind, batch, mode, bw, mtx, GAPS, nthreads=8)
        ncg = np.array([len(c) for c in cigs], np.uint32)
        words = np.concatenate(cigs).astype(np.uint64)
        h1, h2 = checksums_dense(words, ncg)
        out["res%d" % ci] = res; out["ncig%d" % ci] = ncg; out["h1_%d" % ci] = h1; out["h2_%d" % ci] = h2
        print(name, "score sum", int(res[:, 0].astype(np.int64).sum()), "cigar words", int(ncg.sum()), flush=True)
    np.savez_compressed(os.path.join(HERE, "real_ont.npz"), **out)
    print("wrote", os.path.join(HERE, "real_ont.npz"), os.path.getsize(os.path.join(HERE, "real_ont.npz")), "bytes")


if __name__ == "__main__":
    main()
