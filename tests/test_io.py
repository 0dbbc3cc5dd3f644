"""Ingest / egress next to the path, on the CPU: the sequence-file reader against hand-made files (FASTA with lower case,
N, wrapped lines, descriptions, an empty record; gzipped FASTQ) and the text formatter against the reference's own command line
(tests/golden/cli_small_*.txt, written by oracle/_ref/bsalign_ref: tests/golden/make_cli_golden.py).  The alignment results fed to the
formatter come from the oracle (a test)."""
import os

import numpy as np

import checkers as ck
from bsalign_b200 import api, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reader_matches_reference_conventions():
    sf = api.SeqFile(os.path.join(GOLD, "cli_small.fa"))
    assert sf.names == ["r1", "r2", "r3", "r4", "odd_one_out"]          # tag = header up to the first blank; the empty record is dropped
    assert list(sf.len) == [230, 227, 230, 220, 4]
    r1, r3 = sf.bases(0), sf.bases(2)
    assert np.array_equal(r1[:50], r3[:50]) and not r3[50:54].any() and np.array_equal(r1[54:], r3[54:])   # N -> base_bit_table & 3 = A
    assert np.array_equal(sf.bases(4), [0, 1, 2, 3])
    fq = api.SeqFile(os.path.join(GOLD, "cli_small.fq.gz"))
    assert fq.names == ["a1", "a2"] and np.array_equal(fq.bases(0), r1) and np.array_equal(fq.bases(1), sf.bases(1))   # lower case = upper case
    # the words are the reference's BaseBank layout (dna.h:63)
    assert np.array_equal(sf.bits[:len(api.pack_bits(np.concatenate([sf.bases(i) for i in range(5)])))], api.pack_bits(np.concatenate([sf.bases(i) for i in range(5)])))


def test_formatter_reproduces_reference_text():
    mtx = synth.score_matrix(2, -6)
    for src, tag in (("cli_small.fa", "cli_small_fa"), ("cli_small.fq.gz", "cli_small_fq")):
        sf = api.SeqFile(os.path.join(GOLD, src))
        npair = len(sf.len) // 2
        batch = synth.PairBatch.from_lists([(sf.bases(2 * k), sf.bases(2 * k + 1)) for k in range(npair)])
        for name, kind, mode in (("align", "epi8", 0), ("edit", "edit", 0)):
            res, cigs, _ = ck.oracle_batch(kind, batch, mode, 0, mtx, (-3, -2, 0, 0))
            text = b"".join(api.format_pair_text(sf, k, res[k], cigs[k]) for k in range(npair))
            assert text == open(os.path.join(GOLD, "%s.%s.txt" % (tag, name)), "rb").read(), (src, name)
        res, cigs, _ = ck.kmer_batch("oracle", batch, 9)   # `edit -m kmer -k 9` (main.c:196)
        text = b"".join(api.format_pair_text(sf, k, res[k], cigs[k]) for k in range(npair))
        assert text == open(os.path.join(GOLD, "%s.kmer.txt" % tag), "rb").read(), (src, "kmer")


def test_binary_msa_round_trips_the_reference_bytes(tmp_path):
    """tests/golden/msa_small.bin was written by the reference's own dump_binary_msa_bspoa (two jobs, one with metadata): the library reads
    it and writes the same bytes back."""
    src = os.path.join(GOLD, "msa_small.bin")
    msas = api.read_binary_msa(src)
    assert len(msas) == 2 and msas[0]["meta"] == b"job-a" and msas[1]["meta"] == b""
    assert msas[0]["nseq"] == 7 and msas[1]["nseq"] == 10 and msas[0]["cols"].shape == (msas[0]["mlen"], 8)   # 6 / 9 reads; the reference's realignment round leaves one all-5 row in front
    assert msas[0]["cols"].max() <= 5 and 250 < msas[0]["mlen"] < 400
    dst = tmp_path / "copy.bin"
    api.write_binary_msa(str(dst), msas)
    assert open(dst, "rb").read() == open(src, "rb").read()
