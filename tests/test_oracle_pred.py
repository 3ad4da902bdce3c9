"""oracle.predict (oracle/rmd_oracle.c: oracle_predict_ex) against reference samples / predicted blocks of the reference's own
TComPrediction::predIntraAng, dumped by oracle/_ref/TAppEncoder_predtrace during an encode (tests/golden/
pred_trace_192x128_qp32.npz, tools/gen_golden_tq.py): luma with the edge filters and chroma without them."""
import os

import numpy as np

from conftest import GOLDEN


def test_prediction_equals_the_references_own(oracle):
    z = np.load(os.path.join(GOLDEN, "pred_trace_192x128_qp32.npz"))
    hdr = z["hdr"]
    seen = set()
    for i, h in enumerate(hdr):
        comp, mode, n, filt, edge = (int(v) for v in h)
        line = z["line"][z["line_off"][i]:z["line_off"][i + 1]]
        want = z["pred"][z["pred_off"][i]:z["pred_off"][i + 1]].reshape(n, n)
        got = oracle.predict(line, n, mode, edge=(comp == 0 and edge == 1))
        assert (got == want).all(), (i, h.tolist())
        seen.add((comp != 0, n, mode))
    assert len(hdr) > 1000
    assert {n for c, n, m in seen if not c} == {4, 8, 16, 32, 64} and {m for c, n, m in seen if not c} == set(range(35))
    assert {n for c, n, m in seen if c} >= {4, 8, 16} and {m for c, n, m in seen if c} >= {0, 1, 10, 26}


def test_chroma_blocks_have_no_edge_filters(oracle):
    """DC, pure vertical and pure horizontal blocks of 4..16 differ between luma and chroma only in their first row / column."""
    rng = np.random.default_rng(2)
    for n in (4, 8, 16):
        line = rng.integers(0, 256, 4 * n + 1).astype(np.int16)
        for mode in (1, 10, 26):
            a, b = oracle.predict(line, n, mode, edge=True), oracle.predict(line, n, mode, edge=False)
            assert (a[1:, 1:] == b[1:, 1:]).all() and (a != b).any()
    line = rng.integers(0, 256, 129).astype(np.int16)
    assert (oracle.predict(line, 32, 1, edge=True) == oracle.predict(line, 32, 1, edge=False)).all()
