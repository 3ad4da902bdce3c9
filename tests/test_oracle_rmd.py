"""The RMD oracle (oracle/rmd_oracle.c) against per-mode SAD / mode bits / candidate lists printed
by the reference encoder itself (oracle/_ref/TAppEncoder_trace, DEBUG_INTRA_SEARCH_COSTS) together
with its reconstruction -- fixture made by tools/gen_golden.py."""
import math
import os

import numpy as np

from conftest import GOLDEN


def _load():
    return np.load(os.path.join(GOLDEN, "rmd_trace_192x128_qp32.npz"))


def _pus(oracle, g):
    H, W = g["Y"].shape
    out = []
    for a, lab in enumerate(g["labels"]):
        out += [tuple(int(v) for v in p) for p in oracle.enum_ctu_pus(lab, a % (W // 64), a // (W // 64), W, H)]
    return out


def test_pu_enumeration_matches_encoder_visit_count(oracle):
    g = _load()
    pus = _pus(oracle, g)
    assert len(pus) == g["sad"].shape[0]
    assert {p[2] for p in pus} == {64, 32, 16, 8, 4}


def test_satd_bit_exact_vs_reference_trace(oracle):
    """Every 2Nx2N PU (64..8) and the first NxN PU see only final reconstruction -> must match the
    reference exactly.  NxN PUs 2..4 read the NxN trial's own reconstruction, which survives in the
    output only when the CU finally chose NxN; those are compared when they match and counted."""
    g = _load()
    Y, rec = g["Y"], g["rec"]
    exact, later_ok, later = 0, 0, 0
    for i, (x, y, n, part) in enumerate(_pus(oracle, g)):
        s = oracle.pu_satd35(Y, x, y, n, oracle.build_ref_line(rec, x, y, n))
        if part <= 1:
            assert (s == g["sad"][i]).all(), (i, x, y, n, part)
            exact += 1
        else:
            later += 1
            later_ok += int((s == g["sad"][i]).all())
    assert exact == 81 + 28 and later == 84 and later_ok >= 30, (exact, later, later_ok)


def test_candidate_list_matches_reference(oracle):
    g = _load()
    sl = math.sqrt(0.57 * 2 ** ((int(g["qp"]) - 12) / 3.0))       # TEncSlice.cpp:487
    for i, (x, y, n, part) in enumerate(_pus(oracle, g)):
        keep = 3 if n >= 16 else 8
        ref = g["cand"][i, :g["ncand"][i]]
        lst = oracle.cand_list(g["sad"][i], g["bits"][i], sl, n, np.zeros(3, np.int32), 0)
        assert (lst == ref[:keep]).all(), i
        mpms = {m for m in range(35) if g["bits"][i][m] < g["bits"][i].max()}
        assert all(int(m) in mpms for m in ref[keep:]), i           # appended entries are MPMs


def test_mpm_rules(oracle):
    m, k = oracle.mpm(-1, -1)
    assert list(m) == [0, 1, 26] and k == 1
    m, k = oracle.mpm(10, 10)
    assert list(m) == [10, 9, 11] and k == 1
    m, k = oracle.mpm(2, 2)
    assert list(m) == [2, 33, 3] and k == 1
    m, k = oracle.mpm(0, 26)
    assert list(m) == [0, 26, 1] and k == 2
    m, k = oracle.mpm(5, 7)
    assert list(m) == [5, 7, 0] and k == 2
    m, k = oracle.mpm(0, 1)
    assert list(m) == [0, 1, 26] and k == 2


def test_mpm_append(oracle):
    satd = np.arange(35, dtype=np.uint32) * 10 + 100
    bits = np.full(35, 6, np.uint32)
    lst = oracle.cand_list(satd, bits, 7.6, 32, np.array([20, 1, 30], np.int32), 2)
    assert list(lst) == [0, 1, 2, 20]          # 1 already listed, 30 beyond the first two MPMs


def test_planar_dc_flat(oracle):
    line = np.full(4 * 8 + 1, 77, np.int16)
    for mode in range(35):
        assert (oracle.predict(line, 8, mode) == 77).all()


def test_no_neighbours_gives_128(oracle):
    pic = np.zeros((64, 64), np.uint8)
    assert (oracle.build_ref_line(pic, 0, 0, 16) == 128).all()
