"""Stage-by-stage parity of the tcgen05 CNN kernels (K1..K4) with the CPU replay of the same
formulation (tools/tc_emulate.py) and with the fp32 oracle.  Run on the GPU box: pytest -m gpu."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from conftest import unsafe_label_mismatches  # noqa: E402

pytestmark = pytest.mark.gpu


def _f(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def _close(a, b, what):
    a, b = _f(a), _f(b)
    tol = 0.02 + 0.02 * np.abs(b)                       # a few bf16 ulps: summation order differs
    bad = np.abs(a - b) > tol
    assert not bad.any(), "%s: %d of %d differ, worst |d|=%.4f at %d (gpu %.4f ref %.4f)" % (
        what, int(bad.sum()), bad.size, float(np.abs(a - b).max()), int(np.abs(a - b).argmax()),
        float(a.ravel()[np.abs(a - b).argmax()]), float(b.ravel()[np.abs(a - b).argmax()]))


@pytest.mark.parametrize("size", [(128, 64), (416, 240)])
def test_tc_intermediates_vs_emulation(built, host, oracle, weights, pkg, size):
    import tc_emulate
    import tc_pack
    W, H = size
    blob = tc_emulate.Blob(tc_pack.pack(tc_pack.load_hdlw(host.DEFAULT_WEIGHTS)))
    Y, U, V = pkg.synth.synth_frame(W, H, 3)
    dp = host.DepthPredictor(W, H, precision=host.PREC_BF16_TC, rmd=False)
    lab, lg = dp.predict_frame(Y, U, V, want_logits=True)
    cat, a2, feats = dp.debug_copy(0), dp.debug_copy(1), dp.debug_copy(2)
    dp.close()
    cw = (W + 63) // 64
    nctu = dp.nctu
    ecat, ea2, efeat = [], [], np.zeros((nctu * 4, 2048), np.float32)
    for a in range(nctu):
        c, _ = tc_emulate.k1_ctu(oracle.stage_ctu_rgb(Y, U, V, a % cw, a // cw), blob)
        ecat.append(c)
        ea2.append(tc_emulate.k2_ctu(c, blob))
        efeat[4 * a:4 * a + 4] = tc_emulate.k3_ctu(ea2[-1], blob)
    _close(cat, np.concatenate(ecat), "K1 cat")
    _close(a2, np.concatenate(ea2), "K2 a2")
    npad = (nctu * 4 + 127) // 128 * 128
    _close(feats, tc_emulate.feats_layout(efeat, npad), "K3 features")
    olab, olg, mar = oracle.frame_labels(weights, Y, U, V, want_logits=True)
    assert np.abs(lg - olg).max() < 0.25, np.abs(lg - olg).max()
    assert unsafe_label_mismatches(lab, olab, mar, 0.1)[0] == 0
