"""The transform / quantisation oracle (oracle/tq_oracle.c) against outputs of the reference itself
(tools/gen_golden_tq.py): its own xTrMxN / xITrMxN on random blocks, and per-TU dumps of the reference encoder built
with its DEBUG_TRANSFORM_AND_QUANTISE switch (RDOQ and sign-bit hiding off: the flat quantiser)."""
import os

import numpy as np

from conftest import GOLDEN


def test_core_matrices_are_the_hevc_ones(oracle):
    t4, t8 = oracle.tq_matrix(4), oracle.tq_matrix(8)
    assert t4.tolist() == [[64, 64, 64, 64], [83, 36, -36, -83], [64, -64, -64, 64], [36, -83, 83, -36]]       # TComRom.cpp:376-382
    assert t8[1].tolist() == [89, 75, 50, 18, -18, -50, -75, -89] and t8[7].tolist() == [18, -50, 75, -89, 89, -75, 50, -18]
    t16, t32 = oracle.tq_matrix(16), oracle.tq_matrix(32)
    assert (t32[::2, :16] == t16).all() and (t16[::2, :8] == t8).all() and (t8[::2, :4] == t4).all()           # nested even rows
    assert t32[1].tolist()[:16] == [90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4]
    assert t32[31].tolist()[:4] == [4, -13, 22, -31]


def test_transforms_equal_the_references_own_functions(oracle):
    g = np.load(os.path.join(GOLDEN, "tq_transform_ref.npz"))
    for i, n in enumerate(g["sizes"]):
        n = int(n)
        a, b = int(g["off"][i]), int(g["off"][i + 1])
        dst = bool(g["dst"][i])
        assert (oracle.tq_forward(g["resi"][a:b].reshape(n, n), dst) == g["coeff"][a:b].reshape(n, n)).all(), (i, n, dst)
        assert (oracle.tq_inverse(g["icoeff"][a:b].reshape(n, n), dst) == g["iresi"][a:b].reshape(n, n)).all(), (i, n, dst)


def test_tu_pipeline_equals_the_reference_encoder_dump(oracle):
    g = np.load(os.path.join(GOLDEN, "tq_trace_192x128_qp32.npz"))
    seen = set()
    for i, n in enumerate(g["sizes"]):
        n = int(n)
        a, b = int(g["off"][i]), int(g["off"][i + 1])
        c, q, d, r, s = oracle.tq_tu(g["resi"][a:b].reshape(n, n), int(g["qp"][i]), int(g["flags"][i]))
        assert (c.ravel() == g["coeff"][a:b]).all(), ("coeff", i, n)
        assert (q.ravel() == g["level"][a:b]).all(), ("level", i, n)
        assert s == int(np.abs(g["level"][a:b]).sum())
        if g["has_inv"][i]:
            assert (d.ravel() == g["deq"][a:b]).all(), ("deq", i, n)
            assert (r.ravel() == g["rec"][a:b]).all(), ("rec", i, n)
        else:
            assert s == 0                          # the reference skips the inverse path only for all-zero levels
        seen.add((n, int(g["chan"][i]), int(g["flags"][i])))
    assert {(4, 0, 1), (4, 0, 2), (8, 0, 0), (16, 0, 0), (32, 0, 0), (4, 1, 0), (8, 2, 0), (16, 1, 0)} <= seen


def test_rdoq_equals_the_reference_encoder_calls(oracle):
    """oracle/rdoq_oracle.c against calls of the reference's own xRateDistOptQuant (RDOQ + sign-bit hiding, every TU size,
    luma and chroma, the three scan types, transform skip), inputs and outputs dumped by the reference encoder itself."""
    g = np.load(os.path.join(GOLDEN, "tq_rdoq_192x128_qp32.npz"))
    seen = set()
    for i, h in enumerate(g["hdr"]):
        n = int(h[1])
        a, b = int(g["off"][i]), int(g["off"][i + 1])
        lev, s = oracle.rdoq(g["src"][a:b].reshape(n, n), 0 if h[2] == 0 else 1, int(h[6]), int(h[3]), int(h[7]), float(g["lam"][i]),
                             g["est"][i], int(h[8]), int(h[9]), int(h[13]), int(h[10]))
        assert s == int(g["abs_sum"][i]) and (lev.ravel() == g["dst"][a:b]).all(), (i, n, h[2], h[6], h[7])
        seen.add((n, int(h[2] != 0), int(h[6]), int(h[7])))
    assert {(4, 0, 0, 0), (4, 0, 1, 1), (4, 0, 2, 0), (8, 0, 1, 0), (8, 0, 2, 0), (8, 1, 0, 0), (16, 0, 0, 0), (16, 1, 0, 0), (32, 0, 0, 0)} <= seen
