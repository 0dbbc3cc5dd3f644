"""Golden TEXT of the reference's own command line (oracle/_ref/bsalign_ref = /root/reference/main.c compiled by oracle/Makefile).
Run in the build container only:   python tests/golden/make_cli_golden.py

  real_ont_cli.json   md5 + size of the three outputs of example/run.sh:5-9 on example/real.ont.b10M.txt, regenerated from the committed
                      2-bit fixture (tests/golden/real_ont.npz; names k.1 / k.2 like the original file) - they equal BASELINE.md's md5s
  cli_small.fa / cli_small.fq.gz / cli_small.*.txt
                      a few records with what a reader must cope with (lower case, N, wrapped lines, descriptions, an empty record,
                      FASTQ, gzip) and the reference's text for them
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.path.join(ROOT, "oracle", "_ref", "bsalign_ref")
RUN_SH = {"NoBand": ["align", "-M", "2", "-X", "2", "-O", "4", "-E", "2", "-Q", "0", "-P", "0"],
          "Band64": ["align", "-W", "64", "-M", "2", "-X", "2", "-O", "4", "-E", "2", "-Q", "0", "-P", "0", "-m", "overlap"],
          "Edit0": ["edit", "-W", "0"],
          "Kmer13": ["edit", "-m", "kmer", "-k", "13"]}   # not in run.sh: the k-mer guided edit (main.c:196) on the same file


def write_real_fasta(path):
    import real_ont
    batch, _, _, _ = real_ont.load()
    with open(path, "w") as f:
        for i in range(batch.n):
            f.write(">%d.1\n%s\n>%d.2\n%s\n" % (i, "".join("ACGT"[c] for c in batch.query(i)), i, "".join("ACGT"[c] for c in batch.target(i))))


def main():
    assert os.path.exists(REF), "oracle/_ref/bsalign_ref missing: make -C oracle refcli"
    tmp = "/tmp/real_ont_regen.fa"
    write_real_fasta(tmp)
    out = {}
    for name, args in RUN_SH.items():
        txt = subprocess.run([REF] + args + [tmp], capture_output=True, check=True).stdout
        out[name] = {"args": args, "md5": hashlib.md5(txt).hexdigest(), "bytes": len(txt)}
        print(name, out[name]["md5"], len(txt))
    # the original file gives the same text (the fixture round-trips)
    for name, args in RUN_SH.items():
        txt = subprocess.run([REF] + args + ["/root/reference/example/real.ont.b10M.txt"], capture_output=True, check=True).stdout
        assert hashlib.md5(txt).hexdigest() == out[name]["md5"], name
    json.dump(out, open(os.path.join(HERE, "real_ont_cli.json"), "w"), indent=1)
    rng = np.random.default_rng(7)
    q = "".join("ACGT"[c] for c in rng.integers(0, 4, 230))
    t = q[:60] + "ac" + q[60:120].lower() + q[125:]
    fa = ">r1 first read\n%s\n%s\n>r2\tdesc\n%s\n>empty\n>r3\n%s\n>r4\n%s\n>odd_one_out\nACGT\n" % (
        q[:100], q[100:], t, q[:50] + "NNNN" + q[54:], q[:30] + q[40:])
    open(os.path.join(HERE, "cli_small.fa"), "w").write(fa)
    fq = "@a1 x\n%s\n+\n%s\n@a2\n%s\n+a2\n%s\n" % (q, "I" * len(q), t.upper(), "I" * len(t))
    with gzip.open(os.path.join(HERE, "cli_small.fq.gz"), "wt") as f:
        f.write(fq)
    for src in ("cli_small.fa", "cli_small.fq.gz"):
        for name, args in (("align", ["align", "-m", "global"]), ("edit", ["edit"]), ("kmer", ["edit", "-m", "kmer", "-k", "9"])):
            txt = subprocess.run([REF] + args + [os.path.join(HERE, src)], capture_output=True, check=True).stdout
            open(os.path.join(HERE, "%s.%s.txt" % (src.split(".")[0] + "_" + src.split(".")[1], name)), "wb").write(txt)
            print(src, name, len(txt))


def msa_golden():
    """msa_small.bin: two small BSPOA jobs through the unmodified end_bspoa + dump_binary_msa_bspoa (oracle/_ref/libbsref.so: bsref_poa_msa_bytes)."""
    import ctypes
    import checkers as ck
    from bsalign_b200 import synth
    L = ck.ref()
    blob = b""
    for seed, nreads, tlen, meta in ((3, 6, 300, b"job-a"), (4, 9, 450, b"")):
        rng = np.random.default_rng(seed)
        tmpl = rng.integers(0, 4, (1, tlen)).astype(np.uint8)
        reads = []
        for _ in range(nreads):
            t, _l = synth.mutate_batch(rng, tmpl, 0.03, 0.03, 0.04)
            reads.append(t)
        seqs = np.concatenate(reads); ln = np.array([len(r) for r in reads], np.uint32)
        off = np.concatenate([[0], np.cumsum(ln[:-1])]).astype(np.uint64)
        out = ctypes.c_void_p()
        L.bsref_poa_msa_bytes.restype = ctypes.c_int64
        n = L.bsref_poa_msa_bytes(ctypes.c_uint32(nreads), seqs.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p), ln.ctypes.data_as(ctypes.c_void_p),
                                  meta or None, ctypes.c_uint32(len(meta)), ctypes.byref(out))
        blob += ctypes.string_at(out, n)
        L.bsref_free(out)
    open(os.path.join(HERE, "msa_small.bin"), "wb").write(blob)
    print("msa_small.bin", len(blob))


if __name__ == "__main__":
    main()
    msa_golden()
