"""Intra prediction of single blocks on the device (hevcdl_intra_pred, csrc/pred.cuh) through the C-ABI: identical to the
reference's own TComPrediction::predIntraAng on the calls dumped from the reference encoder (tests/golden/
pred_trace_192x128_qp32.npz: luma and chroma, every size and mode it used, filtered and unfiltered references), to the oracle
on random reference lines of every size / mode / edge-filter setting, and -- inside the real encoder (HEVCDL_PRED=1) --
byte-identical bitstreams."""
import os
import re

import numpy as np
import pytest

import hm_util
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp(built, host):
    d = host.DepthPredictor(64, 64, precision=host.PREC_FP32, rmd=False, outputs=0)
    yield d
    d.close()


def test_prediction_vs_the_references_own_calls(dp):
    z = np.load(os.path.join(GOLDEN, "pred_trace_192x128_qp32.npz"))
    hdr = z["hdr"]
    lines = [z["line"][z["line_off"][i]:z["line_off"][i + 1]] for i in range(len(hdr))]
    got = dp.intra_pred(lines, hdr[:, 1], [(h[0] == 0 and h[4] == 1) for h in hdr])
    for i, h in enumerate(hdr):
        want = z["pred"][z["pred_off"][i]:z["pred_off"][i + 1]].reshape(got[i].shape)
        assert (got[i] == want).all(), (i, h.tolist(), int((got[i] != want).sum()))
    assert len(hdr) > 1000 and (hdr[:, 0] != 0).sum() > 100


def test_prediction_vs_oracle_random(dp, oracle):
    """Every size 4..64 x mode 0..34 x edge filters on / off, random and extreme (0 / 255 alternating) reference lines."""
    rng = np.random.default_rng(6)
    lines, modes, edge = [], [], []
    for n in (4, 8, 16, 32, 64):
        for m in range(35):
            for e in (False, True):
                for kind in range(3):
                    if kind == 0:
                        l = rng.integers(0, 256, 4 * n + 1)
                    elif kind == 1:
                        l = (np.arange(4 * n + 1) % 2) * 255
                    else:
                        l = np.clip(np.cumsum(rng.integers(-9, 10, 4 * n + 1)) + 128, 0, 255)
                    lines.append(l.astype(np.int16)); modes.append(m); edge.append(e)
    got = dp.intra_pred(lines, modes, edge)
    for i, l in enumerate(lines):
        n = (len(l) - 1) // 4
        want = oracle.predict(l, n, modes[i], edge=edge[i])
        assert (got[i] == want).all(), (n, modes[i], edge[i], int((got[i] != want).sum()))
    assert dp.intra_pred([], [], []) == []


def test_prediction_rejects_bad_requests(dp, host):
    with pytest.raises(host.HevcdlError):
        dp.intra_pred([np.zeros(17, np.int16)], [35], [False])            # mode out of range
    with pytest.raises(host.HevcdlError):
        dp.intra_pred([np.zeros(9, np.int16)], [0], [False])              # 2x2 blocks do not exist


@pytest.mark.skipif(not hm_util.have("ref", "dec", "hevcdl"), reason="reference / drop-in encoder binaries not built")
@pytest.mark.parametrize("w,h,qp", [(192, 128, 32)])
def test_dropin_prediction_on_the_device_keeps_the_bitstream(tmp_path, built, host, pkg, w, h, qp):
    """HEVCDL_PRED=1: every TComPrediction::predIntraAng call of the encode (first pass and RD pass, luma and chroma) is
    computed on the B200 from HM's own reference samples: the bitstream must stay byte-identical to the reference's."""
    frames = [pkg.synth.synth_frame(w, h, 130 + i) for i in range(1)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dpx = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dpx.predict_frame(Y, U, V, frame=f))
    dpx.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, 1, qp)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 1, qp, env={"HEVCDL_PRECISION": "fp32", "HEVCDL_PRED": "1", "HEVCDL_VERBOSE": "1"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-600:])
    m = re.search(r"blocks predicted on the device (\d+) / by the reference's code (\d+)", rb["stderr"])
    assert m and int(m.group(1)) > 5000 and int(m.group(2)) == 0, rb["stderr"][-600:]
    assert ra["sha1"] == rb["sha1"]
    ok, out = hm_util.decode_ok(str(b))
    assert ok, out[-400:]
