"""Deblocking filter on the device (hevcdl_deblock_frame, csrc/dbf.cuh) through the C-ABI: identical to the reference's own
loopFilterPic on the dumped pictures (tests/golden/dbf_pictures.npz), to the oracle on synthetic pictures of other sizes, QPs
and offsets, and -- inside the real encoder (HEVCDL_DBF=1) -- byte-identical bitstreams.  The same three levels for the two SAO passes (hevcdl_sao_stats, hevcdl_sao_apply, csrc/sao.cuh,
HEVCDL_SAO=1)."""
import os
import re

import numpy as np
import pytest

import hm_util
from test_oracle_dbf import cases, sao_apply_cases, sao_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp(built, host):
    d = host.DepthPredictor(64, 64, precision=host.PREC_FP32, rmd=False, outputs=0)
    yield d
    d.close()


def test_deblocking_vs_the_references_own_filter(dp):
    n = 0
    for k, h, pre, post, tu, qp in cases():
        out = dp.deblock_frame(*pre, tu, qp, int(h[4]), int(h[5]), int(h[6]), int(h[7]))
        for a, b, name in zip(out, post, "YUV"):
            assert (a == b).all(), (k, name, int((a != b).sum()))
        n += 1
    assert n == 3


def test_deblocking_vs_oracle_synthetic(dp, oracle, pkg):
    """Random TU-size maps (consistent quadtrees of 32/16/8/4 blocks), per-unit QPs over the whole range, non-zero deblocking
    and chroma QP offsets, pictures from 8x8 to 1920x1080 with blocky content so that every filter branch fires."""
    rng = np.random.default_rng(4)
    for (W, H) in ((8, 8), (16, 8), (8, 24), (72, 200), (416, 240), (1920, 1080)):
        w4, h4 = W // 4, H // 4
        tu = np.full((h4, w4), 2, np.uint8)
        for by in range(0, H, 32):
            for bx in range(0, W, 32):
                def fill(x0, y0, lg):
                    s = 1 << lg
                    if x0 + s <= W and y0 + s <= H and (lg == 2 or rng.random() < 0.45):
                        tu[y0 // 4:(y0 + s) // 4, x0 // 4:(x0 + s) // 4] = lg
                    elif lg > 2:
                        for k in range(4):
                            if x0 + (k & 1) * (s // 2) < W and y0 + (k >> 1) * (s // 2) < H:
                                fill(x0 + (k & 1) * (s // 2), y0 + (k >> 1) * (s // 2), lg - 1)
                fill(bx, by, 5)
        qp = rng.integers(0, 52, (h4, w4)).astype(np.int8) if W < 400 else np.full((h4, w4), int(rng.integers(20, 45)), np.int8)
        base = np.kron(rng.integers(30, 226, ((H + 7) // 8, (W + 7) // 8)), np.ones((8, 8)))[:H, :W]       # blocky: edges on the 8x8 grid
        Y = np.clip(base + rng.integers(-3, 4, (H, W)), 0, 255).astype(np.uint8)
        U = np.clip(np.kron(rng.integers(60, 196, ((H + 15) // 16, (W + 15) // 16)), np.ones((8, 8)))[:H // 2, :W // 2] + rng.integers(-2, 3, (H // 2, W // 2)), 0, 255).astype(np.uint8)
        V = np.clip(U.astype(np.int32)[::-1, ::-1] + 7, 0, 255).astype(np.uint8)
        offs = (int(rng.integers(-2, 3)), int(rng.integers(-2, 3)), int(rng.integers(-4, 5)), int(rng.integers(-4, 5)))
        want = oracle.deblock_frame(Y, U, V, tu, qp, *offs)
        got = dp.deblock_frame(Y, U, V, tu, qp, *offs)
        for a, b, name in zip(got, want, "YUV"):
            assert (a == b).all(), (W, H, name, int((a != b).sum()))
        if W >= 72:
            assert (want[0] != Y).sum() > 50


@pytest.mark.skipif(not hm_util.have("ref", "dec", "hevcdl"), reason="reference / drop-in encoder binaries not built")
@pytest.mark.parametrize("w,h,qp", [(192, 128, 32), (416, 240, 27)])
def test_dropin_deblocking_on_the_device_keeps_the_bitstream(tmp_path, built, host, pkg, w, h, qp):
    """HEVCDL_DBF=1: TComLoopFilter::loopFilterPic of every picture runs on the B200.  The deblocked reconstruction feeds SAO and
    the picture-hash SEI, so any differing sample would change the bitstream: it must stay byte-identical to the reference's."""
    frames = [pkg.synth.synth_frame(w, h, 95 + i) for i in range(2)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dpx = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dpx.predict_frame(Y, U, V, frame=f))
    dpx.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, 2, qp)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 2, qp, env={"HEVCDL_DBF": "1", "HEVCDL_VERBOSE": "1"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-400:])
    m = re.search(r"pictures deblocked on the device (\d+) / by the reference's filter (\d+)", rb["stderr"])
    assert m and int(m.group(1)) == 2 and int(m.group(2)) == 0, rb["stderr"][-300:]
    assert ra["sha1"] == rb["sha1"] and (ra["kbps"], ra["psnr_y"]) == (rb["kbps"], rb["psnr_y"])
    if w % 64 == 0 and h % 64 == 0:
        ok, out = hm_util.decode_ok(str(b))
        assert ok, out[-400:]


def test_sao_statistics_vs_the_references_own(dp):
    n = 0
    for k, org, src, want in sao_cases():
        got = dp.sao_stats(org, src)
        assert got.shape == want.shape and (got == want).all(), (k, int((got != want).sum()))
        n += 1
    assert n == 2


def test_sao_statistics_vs_oracle_synthetic(dp, oracle):
    """Picture sizes with partial CTUs on the right and bottom, one-CTU pictures, and 1080p; the 'deblocked' picture is the
    original plus noise and a smoothing pass, so that every edge class and most bands are populated."""
    rng = np.random.default_rng(9)
    for (W, H) in ((64, 64), (8, 8), (72, 40), (200, 136), (416, 240), (1920, 1080)):
        org, src = [], []
        for c in range(3):
            w, h = (W, H) if c == 0 else (W // 2, H // 2)
            o = np.clip(np.kron(rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8)), np.ones((8, 8)))[:h, :w] + rng.integers(-20, 21, (h, w)), 0, 255)
            s = o + rng.integers(-6, 7, (h, w))
            s[:, 1:-1] = (s[:, :-2] + 2 * s[:, 1:-1] + s[:, 2:]) // 4 if w > 2 else s[:, 1:-1]
            org.append(o.astype(np.uint8))
            src.append(np.clip(s, 0, 255).astype(np.uint8))
        want = oracle.sao_stats(org, src)
        got = dp.sao_stats(org, src)
        assert got.shape == want.shape and (got == want).all(), (W, H, int((got != want).sum()))
        if W >= 200:
            assert (want[:, :, :, 1] > 0).sum() > 100


@pytest.mark.skipif(not hm_util.have("ref", "dec", "hevcdl"), reason="reference / drop-in encoder binaries not built")
@pytest.mark.parametrize("w,h,qp", [(192, 128, 37), (416, 240, 32)])
def test_dropin_sao_statistics_on_the_device_keep_the_bitstream(tmp_path, built, host, pkg, w, h, qp):
    """HEVCDL_SAO=1: TEncSampleAdaptiveOffset::getStatistics of every picture and the application of the decided offsets
    (TComSampleAdaptiveOffset::offsetCTU, one deferred pass per picture) run on the B200; the offsets SAO signals are decided from
    these sums and the decoded-picture hash is taken from the offset picture, so any differing count, difference or sample would
    change the bitstream: it must stay byte-identical."""
    frames = [pkg.synth.synth_frame(w, h, 120 + i) for i in range(2)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dpx = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dpx.predict_frame(Y, U, V, frame=f))
    dpx.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, 2, qp)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 2, qp,
                        env={"HEVCDL_PRECISION": "fp32", "HEVCDL_SAO": "1", "HEVCDL_DBF": "1", "HEVCDL_VERBOSE": "1"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-600:])
    assert re.search(r"SAO statistics passes on the device 2 / by the reference's code 0", rb["stderr"]), rb["stderr"][-600:]
    assert re.search(r"SAO offsets applied on the device for 2 pictures / by the reference's code for 0", rb["stderr"]), rb["stderr"][-600:]
    assert re.search(r"in-loop passes of 2 pictures shared one upload", rb["stderr"]), rb["stderr"][-800:]
    # ... and with the three passes kept apart (three round trips per picture): the same bitstream
    c = tmp_path / "sep"
    c.mkdir()
    hm_util.write_yuv(str(c / "in.yuv"), frames)
    rc_ = hm_util.encode("hevcdl", str(c), "in.yuv", w, h, 2, qp,
                         env={"HEVCDL_PRECISION": "fp32", "HEVCDL_SAO": "1", "HEVCDL_DBF": "1", "HEVCDL_INLOOP_FUSE": "0", "HEVCDL_VERBOSE": "1"})
    assert rc_["rc"] == 0 and re.search(r"in-loop passes of 0 pictures shared one upload", rc_["stderr"]) and rc_["sha1"] == ra["sha1"]
    assert ra["sha1"] == rb["sha1"]
    ok, out = hm_util.decode_ok(str(b))
    assert ok, out[-400:]


@pytest.mark.parametrize("W,H", [(3840, 2160), (7680, 4320)])
def test_inloop_passes_at_baseline_sizes(dp, oracle, W, H):
    """BASELINE configs[3] / [4] picture sizes: deblocking and SAO statistics equal the oracle on the whole picture, and the
    size-independent properties hold: a flat picture passes the filter unchanged; per CTU and component the band-offset
    counts add up to the samples SAO may touch (block minus the not-yet-deblocked right / bottom lines) and the band-offset
    differences to the plain sum of (original - reconstructed) over the same samples."""
    rng = np.random.default_rng(W)
    base = np.kron(rng.integers(30, 226, (H // 8, W // 8)), np.ones((8, 8)))
    Y = np.clip(base + rng.integers(-3, 4, (H, W)), 0, 255).astype(np.uint8)
    U = np.clip(np.kron(rng.integers(60, 196, (H // 16, W // 16)), np.ones((8, 8))) + rng.integers(-2, 3, (H // 2, W // 2)), 0, 255).astype(np.uint8)
    V = U[::-1, ::-1].copy()
    tu = np.kron(rng.integers(2, 6, (H // 32 + 1, W // 32 + 1)), np.ones((8, 8), np.int64))[:H // 4, :W // 4].astype(np.uint8)
    qp = np.kron(rng.integers(18, 46, (H // 64 + 1, W // 64 + 1)), np.ones((16, 16), np.int64))[:H // 4, :W // 4].astype(np.int8)
    want = oracle.deblock_frame(Y, U, V, tu, qp, 1, -1, 2, -2)
    got = dp.deblock_frame(Y, U, V, tu, qp, 1, -1, 2, -2)
    for a, b, name in zip(got, want, "YUV"):
        assert (a == b).all(), (name, int((a != b).sum()))
    assert (got[0] != Y).sum() > W * H // 100
    flat = [np.full_like(Y, 97), np.full_like(U, 140), np.full_like(V, 99)]
    for a, b in zip(dp.deblock_frame(*flat, tu, qp), flat):
        assert (a == b).all()
    st = dp.sao_stats((Y, U, V), got)
    assert (st == oracle.sao_stats((Y, U, V), got)).all()
    cw, chh = (W + 63) // 64, (H + 63) // 64
    for c, (o, r, s, skr, skb) in enumerate(((Y, got[0], 64, 5, 4), (U, got[1], 32, 3, 2), (V, got[2], 32, 3, 2))):
        h, w = o.shape
        d = o.astype(np.int64) - r
        cnt = np.zeros((chh, cw), np.int64)
        dif = np.zeros((chh, cw), np.int64)
        for cy in range(chh):
            y1 = min(h, (cy + 1) * s) - (skb if cy < chh - 1 else 0)
            for cx in range(cw):
                x1 = min(w, (cx + 1) * s) - (skr if cx < cw - 1 else 0)
                cnt[cy, cx] = (y1 - cy * s) * (x1 - cx * s)
                dif[cy, cx] = d[cy * s:y1, cx * s:x1].sum()
        assert (st[:, c, 4, 1, :].sum(axis=1).reshape(chh, cw) == cnt).all(), c
        assert (st[:, c, 4, 0, :].sum(axis=1).reshape(chh, cw) == dif).all(), c


def test_sao_application_vs_the_references_own(dp):
    n = 0
    for k, src, t, o, res in sao_apply_cases():
        got = dp.sao_apply(src, t, o)
        for a, b, name in zip(got, res, "YUV"):
            assert (a == b).all(), (k, name, int((a != b).sum()))
        n += 1
    assert n == 3


def test_sao_application_vs_oracle_synthetic(dp, oracle):
    """Random types and offsets (-7..7) per CTU and component on pictures from one partial CTU to 1920x1080 and 3840x2160:
    every type meets every border configuration; all-off parameters return the picture unchanged."""
    rng = np.random.default_rng(12)
    for (W, H) in ((8, 8), (64, 64), (72, 40), (200, 136), (416, 240), (1920, 1080), (3840, 2160)):
        src = [rng.integers(0, 256, s).astype(np.uint8) if W < 1000 else
               np.clip(np.kron(rng.integers(0, 256, ((s[0] + 3) // 4, (s[1] + 3) // 4)), np.ones((4, 4)))[:s[0], :s[1]] + rng.integers(-9, 10, s), 0, 255).astype(np.uint8)
               for s in ((H, W), (H // 2, W // 2), (H // 2, W // 2))]
        n = ((W + 63) // 64) * ((H + 63) // 64)
        t = rng.integers(-1, 5, (n, 3)).astype(np.int8)
        o = rng.integers(-7, 8, (n, 3, 32)).astype(np.int8)
        want = oracle.sao_apply(src, t, o)
        got = dp.sao_apply(src, t, o)
        for a, b, name in zip(got, want, "YUV"):
            assert (a == b).all(), (W, H, name, int((a != b).sum()))
        same = dp.sao_apply(src, np.full((n, 3), -1, np.int8), o)
        assert all((a == b).all() for a, b in zip(same, src))


def test_fused_inloop_call_and_resident_offsets(dp, oracle, host):
    """hevcdl_inloop_frame = deblocking + SAO statistics in one round trip, then hevcdl_sao_apply with no source = the picture
    that call left on the device: all three results equal the oracle's; another in-loop call drops the resident picture."""
    rng = np.random.default_rng(31)
    for (W, H) in ((72, 40), (416, 240), (1920, 1080)):
        base = np.kron(rng.integers(30, 226, ((H + 7) // 8, (W + 7) // 8)), np.ones((8, 8)))[:H, :W]
        Y = np.clip(base + rng.integers(-3, 4, (H, W)), 0, 255).astype(np.uint8)
        U = np.clip(np.kron(rng.integers(60, 196, ((H + 15) // 16, (W + 15) // 16)), np.ones((8, 8)))[:H // 2, :W // 2] + rng.integers(-2, 3, (H // 2, W // 2)), 0, 255).astype(np.uint8)
        V = U[::-1, ::-1].copy()
        org = [np.clip(p.astype(np.int32) + rng.integers(-5, 6, p.shape), 0, 255).astype(np.uint8) for p in (Y, U, V)]
        tu = np.kron(rng.integers(2, 6, (H // 32 + 1, W // 32 + 1)), np.ones((8, 8), np.int64))[:H // 4, :W // 4].astype(np.uint8)
        qp = np.full(tu.shape, 33, np.int8)
        want_rec = oracle.deblock_frame(Y, U, V, tu, qp, 0, 1, -1, 2)
        want_st = oracle.sao_stats(org, want_rec)
        rec, st = dp.inloop_frame(Y, U, V, tu, qp, org, 0, 1, -1, 2)
        assert all((a == b).all() for a, b in zip(rec, want_rec)) and (st == want_st).all(), (W, H)
        n = ((W + 63) // 64) * ((H + 63) // 64)
        t = rng.integers(-1, 5, (n, 3)).astype(np.int8)
        o = rng.integers(-7, 8, (n, 3, 32)).astype(np.int8)
        want = oracle.sao_apply(want_rec, t, o)
        got = dp.sao_apply(None, t, o, shape=(H, W))
        assert all((a == b).all() for a, b in zip(got, want)), (W, H)
        again = dp.sao_apply(None, t, o, shape=(H, W))                 # the source is untouched: still resident
        assert all((a == b).all() for a, b in zip(again, want))
    n2 = ((W + 63) // 64) * ((H // 2 // 8 * 8 + 63) // 64)
    with pytest.raises(host.HevcdlError):
        dp.sao_apply(None, np.full((n2, 3), -1, np.int8), np.zeros((n2, 3, 32), np.int8), shape=(H // 2 // 8 * 8, W))   # no resident picture of that size
    dp.sao_stats(org, rec)                                              # any other in-loop call drops it
    with pytest.raises(host.HevcdlError):
        dp.sao_apply(None, t, o, shape=(H, W))


def test_page_locked_planes_take_the_direct_copy_path(dp, oracle, host):
    """hevcdl_host_register: planes in page-locked memory are copied with their stride straight to / from the device, planes in
    ordinary memory through the staging buffer -- mixed in one call here (Y page-locked with a row stride larger than the width,
    chroma ordinary); results equal the oracle's either way."""
    import ctypes as C
    rng = np.random.default_rng(41)
    W, H, S = 416, 240, 480                     # luma stride 480 > width
    Yb = np.zeros((H, S), np.int16)
    Yb[:, :W] = np.clip(np.kron(rng.integers(30, 226, (H // 8, W // 8)), np.ones((8, 8))) + rng.integers(-3, 4, (H, W)), 0, 255)
    Yb[:, W:] = -77                              # must come back untouched
    U = np.clip(np.kron(rng.integers(60, 196, (H // 16, W // 16)), np.ones((8, 8))) + rng.integers(-2, 3, (H // 2, W // 2)), 0, 255).astype(np.int16)
    V = U[::-1, ::-1].copy()
    tu = np.kron(rng.integers(2, 6, (H // 32 + 1, W // 32 + 1)), np.ones((8, 8), np.int64))[:H // 4, :W // 4].astype(np.uint8)
    qp = np.full(tu.shape, 35, np.int8)
    want = oracle.deblock_frame(Yb[:, :W], U, V, tu, qp)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    assert dp.lib.hevcdl_host_register(vp(Yb), Yb.nbytes) == 0
    try:
        rc = dp.lib.hevcdl_deblock_frame(dp.h, vp(Yb), S, vp(U), vp(V), W // 2, W, H, vp(np.ascontiguousarray(tu.ravel())), vp(np.ascontiguousarray(qp.ravel())), 0, 0, 0, 0)
        assert rc == 0
        assert (Yb[:, :W] == want[0]).all() and (Yb[:, W:] == -77).all() and (U == want[1]).all() and (V == want[2]).all()
    finally:
        assert dp.lib.hevcdl_host_unregister(vp(Yb)) == 0


def test_inloop_entry_points_reject_bad_arguments(dp, host):
    """Sizes that are not multiples of 8, TU sizes / QPs / SAO types out of range, missing planes: HEVCDL_E_INVAL, nothing
    launched, and the context keeps working afterwards."""
    import ctypes as C
    vp = lambda a: C.c_void_p(a.ctypes.data)
    W, H = 64, 64
    y, u, v = np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)
    tu, qp = np.full(W * H // 16, 3, np.uint8), np.full(W * H // 16, 30, np.int8)
    L = dp.lib
    E_INVAL = L.hevcdl_deblock_frame(dp.h, vp(y), W, vp(u), vp(v), W // 2, 60, H, vp(tu), vp(qp), 0, 0, 0, 0)       # width not a multiple of 8
    assert E_INVAL != 0
    bad_tu = tu.copy(); bad_tu[5] = 6
    assert L.hevcdl_deblock_frame(dp.h, vp(y), W, vp(u), vp(v), W // 2, W, H, vp(bad_tu), vp(qp), 0, 0, 0, 0) == E_INVAL
    bad_qp = qp.copy(); bad_qp[7] = 52
    assert L.hevcdl_deblock_frame(dp.h, vp(y), W, vp(u), vp(v), W // 2, W, H, vp(tu), vp(bad_qp), 0, 0, 0, 0) == E_INVAL
    assert L.hevcdl_deblock_frame(dp.h, vp(y), W, vp(u), vp(v), W // 2, W, H, vp(tu), vp(qp), 7, 0, 0, 0) == E_INVAL    # beta offset out of range
    assert L.hevcdl_deblock_frame(dp.h, vp(y), W - 8, vp(u), vp(v), W // 2, W, H, vp(tu), vp(qp), 0, 0, 0, 0) == E_INVAL  # stride < width
    st = np.zeros((1, 3, 5, 2, 32), np.int64)
    assert L.hevcdl_sao_stats(dp.h, vp(y), vp(u), None, W, W // 2, vp(y), vp(u), vp(v), W, W // 2, W, H, vp(st)) == E_INVAL
    assert L.hevcdl_inloop_frame(dp.h, vp(y), W, vp(u), vp(v), W // 2, W, H, vp(tu), vp(qp), 0, 0, 0, 0, vp(y), vp(u), vp(v), W, W // 2, None) == E_INVAL
    prm = np.zeros((1, 3), host.SAO_PARAM_DTYPE)
    prm["type"] = 5
    assert L.hevcdl_sao_apply(dp.h, vp(y), vp(u), vp(v), W, W // 2, vp(y.copy()), vp(u.copy()), vp(v.copy()), W, W // 2, W, H, vp(prm)) == E_INVAL
    got = dp.deblock_frame(y, u, v, tu.reshape(H // 4, W // 4), qp.reshape(H // 4, W // 4))          # still alive
    assert all((a == 0).all() for a in got)
