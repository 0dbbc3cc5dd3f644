"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/bsalign_b200.h declares,
band-width rules, the no-CPU-fallback contract, synthetic batch determinism, the compat header compiles against
the reference's own headers (when they are present)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import checkers as ck
from bsalign_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bsalign_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libbsalign_b200.so does not export %s" % n
    assert b"sm_100a" in L.bsb200_version()


def test_bandwidth_rules_match_reference_rules():
    L = api.lib()
    for qlen, bw, exp in [(1000, 0, 1008), (1000, 128, 128), (5, 0, 16), (300, 500, 512), (16, 16, 16)]:
        assert L.bsb200_epi8_bandwidth(qlen, bw) == exp      # bsalign.h:3861-3862
    # bsalign.h:1055-1067
    for qlen, tlen, mode, bw, exp in [(300, 300, 0, 64, 64), (300, 300, 0, 0, 320), (300, 300, 1, 64, 320), (300, 300, 2, 64, 320),
                                      (300, 2, 0, 64, 192), (100, 100, 0, 128, 128), (64, 64, 0, 0, 64)]:
        assert L.bsb200_edit_bandwidth(qlen, tlen, mode, bw) == exp


def test_no_cpu_fallback_without_gpu():
    L = api.lib()
    if L.bsb200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        api.Context(0)
    m = api.banded_striped_epi8_seqalign_set_score_matrix(2, -6)
    with pytest.raises(RuntimeError):
        api.banded_striped_epi8_seqalign_pairwise(np.array([0, 1, 2], np.uint8), np.array([0, 1, 2], np.uint8), 0, 0, m, -3, -2, 0, 0)


def test_score_matrix_matches_reference_layout():
    m = api.banded_striped_epi8_seqalign_set_score_matrix(2, -6)   # bsalign.h:323
    assert m.dtype == np.int8 and list(m[[0, 5, 10, 15]]) == [2, 2, 2, 2] and m[1] == -6 and m.sum() == 4 * 2 - 12 * 6


def test_synth_is_deterministic_and_packed():
    a = synth.make_pairs(50, 200, seed=7)
    b = synth.make_pairs(50, 200, seed=7)
    assert np.array_equal(a.seqs, b.seqs) and np.array_equal(a.tlen, b.tlen)
    assert a.seqs.max() <= 3 and int(a.qoff[-1] + a.qlen[-1]) <= len(a.seqs)
    assert abs(float(a.tlen.mean()) - 200 * (1 + 0.03 - 0.04)) < 6


def test_cigar2alnstr():
    q = np.array([0, 1, 2, 3, 0], np.uint8)
    t = np.array([0, 1, 3, 0], np.uint8)
    rs = dict(qb=0, tb=0)
    a, b, c = api.seqalign_cigar2alnstr(q, t, rs, [(2 << 4) | 0, (1 << 4) | 1, (2 << 4) | 0])
    assert a == "ACGTA" and b == "AC-TA" and c == "||-||"


@pytest.mark.skipif(not os.path.exists("/root/reference/bsalign.h"), reason="reference headers not on this box")
def test_compat_header_compiles_against_reference_headers(tmp_path):
    src = tmp_path / "compat_check.c"
    src.write_text('#include "bsalign.h"\n#define BSALIGN_B200_OVERRIDE\n#include "bsalign_b200_compat.h"\n'
                   'int main(void){ b1i m[16]; u4v *cg = init_u4v(8); u1i q[4] = {0,1,2,3};\n'
                   ' banded_striped_epi8_seqalign_set_score_matrix(m, 2, -6);\n'
                   ' seqalign_result_t r = banded_striped_epi8_seqalign_pairwise(q, 4, q, 4, NULL, cg, SEQALIGN_MODE_GLOBAL, 0, m, -3, -2, 0, 0, 0);\n'
                   ' seqalign_result_t e = striped_seqedit_pairwise(q, 4, q, 4, SEQALIGN_MODE_GLOBAL, 0, NULL, cg, 0);\n'
                   ' return r.score + e.score; }\n')
    obj = tmp_path / "compat_check.o"
    subprocess.check_call(["gcc", "-c", "-O1", "-w", "-msse4.2", "-mpopcnt", "-D_GNU_SOURCE", "-I/root/reference", "-I" + os.path.join(ROOT, "include"),
                           str(src), "-o", str(obj)])
    syms = subprocess.check_output(["nm", str(obj)]).decode()
    assert "U bsb200_epi8_pairwise" in syms and "U bsb200_edit_pairwise" in syms
