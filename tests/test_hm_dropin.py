"""The drop-in encoder (reference HM_dl sources + this repo's TEncCu::compressCtu over libhevcdl.so,
hm_plugin/) against the UNMODIFIED reference encoder fed the same labels through its ./pred file
handshake: the two must write byte-identical bitstreams (SURVEY.md 7 'minimum slice' (b)), and the
reference decoder must accept the result (MD5 SEI)."""
import importlib
import os

import numpy as np
import pytest

import hm_util

needs_bins = pytest.mark.skipif(not hm_util.have("ref", "dec", "hevcdl"),
                                reason="reference/drop-in encoder binaries not built (need /root/reference at build time)")


@needs_bins
def test_dropin_fails_loudly_without_a_gpu(tmp_path, pkg):
    """No CPU fallback on the named path: without a B200 the drop-in aborts with the library's message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    hm_util.write_yuv(str(tmp_path / "in.yuv"), [pkg.synth.synth_frame(64, 64, 0)])
    r = hm_util.encode("hevcdl", str(tmp_path), "in.yuv", 64, 64, 1, 32)
    assert r["rc"] != 0 and "no CPU fallback" in r["stderr"]


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("w,h,nframes,qp", [(192, 128, 3, 32), (416, 240, 1, 37), (256, 192, 2, 22)])
def test_dropin_bitstream_equals_reference_fed_same_labels(tmp_path, built, host, pkg, w, h, nframes, qp):
    frames = [pkg.synth.synth_frame(w, h, 10 + i) for i in range(nframes)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dp.predict_frame(Y, U, V, frame=f))
    dp.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, nframes, qp)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, nframes, qp, env={"HEVCDL_PRECISION": "fp32"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-400:])
    assert ra["bytes"] > 0 and ra["sha1"] == rb["sha1"], (ra["bytes"], rb["bytes"])
    assert (ra["kbps"], ra["psnr_y"]) == (rb["kbps"], rb["psnr_y"])
    if w % 64 == 0 and h % 64 == 0:            # partial CTUs are non-conformant in the reference itself
        ok, out = hm_util.decode_ok(str(b))
        assert ok, out[-400:]


@needs_bins
@pytest.mark.gpu
def test_dropin_boundary_fix_decodes_cleanly(tmp_path, built, pkg):
    """With HEVCDL_BOUNDARY_FIX=1 the labels of picture-edge CTUs are raised so partial CTUs tile:
    the stream of a non-64-aligned picture passes the decoder's MD5 check, which the reference's does not."""
    w, h = 416, 240
    hm_util.write_yuv(str(tmp_path / "in.yuv"), [pkg.synth.synth_frame(w, h, 3)])
    r = hm_util.encode("hevcdl", str(tmp_path), "in.yuv", w, h, 1, 32, env={"HEVCDL_BOUNDARY_FIX": "1"})
    assert r["rc"] == 0, r["stderr"][-400:]
    ok, out = hm_util.decode_ok(str(tmp_path))
    assert ok, out[-400:]


@needs_bins
@pytest.mark.gpu
def test_dropin_tensor_core_precision_runs(tmp_path, built, pkg):
    w, h = 192, 128
    hm_util.write_yuv(str(tmp_path / "in.yuv"), [pkg.synth.synth_frame(w, h, 5)])
    r = hm_util.encode("hevcdl", str(tmp_path), "in.yuv", w, h, 1, 32, env={"HEVCDL_PRECISION": "bf16", "HEVCDL_VERBOSE": "1"})
    assert r["rc"] == 0 and "kernel launches" in r["stderr"], r["stderr"][-400:]
    ok, out = hm_util.decode_ok(str(tmp_path))
    assert ok, out[-400:]


@needs_bins
@pytest.mark.gpu
def test_dropin_first_pass_satd_served_by_the_device(tmp_path, built, host, pkg):
    """HEVCDL_RMD=1: every (PU, mode) SATD the reference's first pass asks for is found in the device's list --
    35 per PU the encoder visits, none missed -- and the stream still decodes (mode decisions then follow the
    original-reference SATDs: BD-rate clause, tools/bdrate_sweep.py)."""
    import re
    w, h = 192, 128
    frames = [pkg.synth.synth_frame(w, h, 20 + i) for i in range(2)]
    hm_util.write_yuv(str(tmp_path / "in.yuv"), frames)
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=True)
    npu = 0
    for f, (Y, U, V) in enumerate(frames):
        dp.submit(f, Y, U, V)
        npu += len(dp.view(f)["pus"])
        dp.release(f)
    dp.close()
    r = hm_util.encode("hevcdl", str(tmp_path), "in.yuv", w, h, 2, 32, env={"HEVCDL_RMD": "1", "HEVCDL_VERBOSE": "1"})
    assert r["rc"] == 0, r["stderr"][-400:]
    m = re.search(r"first-pass SATDs served (\d+) / missed (\d+)", r["stderr"])
    assert m and int(m.group(2)) == 0 and int(m.group(1)) == 35 * npu, (r["stderr"][-300:], npu)
    ok, out = hm_util.decode_ok(str(tmp_path))
    assert ok, out[-400:]


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("w,h,qp", [(192, 128, 32), (256, 192, 22)])
def test_dropin_exact_rmd_bitstream_equals_reference(tmp_path, built, host, pkg, w, h, qp):
    """HEVCDL_RMD=2: the 35 first-pass SATDs of every PU come from hevcdl_rmd_exact fed HM's own reconstructed
    reference samples.  The device code is bit-exact, so the bitstream must again be byte-identical to the unmodified
    reference encoder fed the same labels -- thousands of PUs of every size with real reconstruction, inside the encoder."""
    import re
    frames = [pkg.synth.synth_frame(w, h, 40 + i) for i in range(2)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dp.predict_frame(Y, U, V, frame=f))
    dp.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, 2, qp)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 2, qp, env={"HEVCDL_PRECISION": "fp32", "HEVCDL_RMD": "2", "HEVCDL_VERBOSE": "1"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-400:])
    m = re.search(r"exact PU calls (\d+)", rb["stderr"])
    assert m and int(m.group(1)) > 100
    assert ra["sha1"] == rb["sha1"], (ra["bytes"], rb["bytes"], ra["kbps"], rb["kbps"])


BITSTREAM_CFG = """InputFile : {yuv}
InputBitDepth : 8
InputChromaFormat : 420
FrameRate : 30
FrameSkip : 0
SourceWidth : {w}
SourceHeight : {h}
FramesToBeEncoded : {n}
Level : 3.1
"""
SHIM = """import runpy, sys
sys.path.insert(0, {root!r})
sys.argv = ["sidecar", {cmd!r}]
runpy.run_module("hevc-deep-learning-pipeline_b200.sidecar", run_name="__main__")
"""


def test_sidecar_parses_bitstream_cfg_by_line_index_and_resets_pred(tmp_path, monkeypatch):
    import importlib
    sidecar = importlib.import_module("hevc-deep-learning-pipeline_b200.sidecar")
    monkeypatch.chdir(tmp_path)
    (tmp_path / "bitstream.cfg").write_text(BITSTREAM_CFG.format(yuv="C:\\seq\\a.yuv", w=416, h=240, n=6))
    cfg = sidecar.parse_bitstream_cfg()
    assert cfg == {"input": "C:/seq/a.yuv", "frame_rate": "30", "width": 416, "height": 240, "frames": 6}   # separators turned round
    (tmp_path / "bitstream.cfg").write_text(BITSTREAM_CFG.format(yuv=".\\Flowervase_416x240_30.yuv", w=416, h=240, n=6))
    assert sidecar.parse_bitstream_cfg()["input"] == "Flowervase_416x240_30.yuv"                          # the reference's own bitstream.cfg:1
    (tmp_path / "pred").mkdir(); (tmp_path / "pred" / "stale").mkdir()
    sidecar.main(["gen_frames"])
    assert os.path.isdir("pred") and os.listdir("pred") == []


@needs_bins
@pytest.mark.gpu
def test_unmodified_reference_encoder_with_the_b200_sidecar(tmp_path, built, pkg):
    """The whole reference system, unmodified binary included: TAppEncoder_ref runs `python gen_frames.py` and
    `python use_model.py` from its working directory (encmain.cpp:53-58,105-108) and polls ./pred; the two scripts here
    are one-line shims onto this repo's sidecar module.  The stream must equal the drop-in encoder's."""
    w, h, n = 192, 128, 3
    frames = [pkg.synth.synth_frame(w, h, 50 + i) for i in range(n)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    (a / "bitstream.cfg").write_text(BITSTREAM_CFG.format(yuv="in.yuv", w=w, h=h, n=n))
    (a / "gen_frames.py").write_text(SHIM.format(root=hm_util.ROOT, cmd="gen_frames"))
    (a / "use_model.py").write_text(SHIM.format(root=hm_util.ROOT, cmd="use_model"))
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, n, 32)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, n, 32, env={"HEVCDL_PRECISION": "fp32"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-600:], rb["stderr"][-400:])
    assert sorted(os.listdir(a / "pred")) == ["0", "1", "2"] and len(os.listdir(a / "pred" / "0")) == 6
    assert ra["sha1"] == rb["sha1"]


@needs_bins
@pytest.mark.gpu
def test_dropin_lookahead_prefetches_frames_and_keeps_the_bitstream(tmp_path, built, host, pkg):
    """a reader thread in the binding preads frames n+1.. of the encoder's input file and submits them
    while HM encodes frame n, so only frame 0 waits for the device.  The stream must stay byte-identical to the reference's
    (fp32 labels), with and without lookahead; a wrong HEVCDL_INPUT must be detected (hash mismatch) and ignored."""
    import re
    w, h, n = 256, 192, 6
    frames = [pkg.synth.synth_frame(w, h, 80 + i) for i in range(n)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    hm_util.write_yuv(str(b / "other.yuv"), frames[::-1])
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dp.predict_frame(Y, U, V, frame=f))
    dp.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, n, 32)
    assert ra["rc"] == 0
    r1 = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, n, 32, out="la.bin", env={"HEVCDL_VERBOSE": "1"})
    r0 = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, n, 32, out="nola.bin", env={"HEVCDL_VERBOSE": "1", "HEVCDL_LOOKAHEAD": "0"})
    r2 = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, n, 32, out="bad.bin", env={"HEVCDL_VERBOSE": "1", "HEVCDL_INPUT": "other.yuv"})
    for r in (r0, r1, r2):
        assert r["rc"] == 0 and r["sha1"] == ra["sha1"], r["stderr"][-400:]
    m = re.search(r"lookahead (\d+): (\d+) frames were on the device before HM asked, (\d+) uploaded from HM's planes, (\d+) mismatches", r1["stderr"])
    assert m and int(m.group(1)) == 3 and int(m.group(2)) == n - 1 and int(m.group(3)) == 1 and int(m.group(4)) == 0, r1["stderr"][-300:]
    m = re.search(r"lookahead (\d+): (\d+) frames were on the device", r0["stderr"])
    assert m and int(m.group(1)) == 0 and int(m.group(2)) == 0
    assert "lookahead disabled" in r2["stderr"]


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("w,h,qp", [(192, 128, 32), (256, 192, 22)])
def test_dropin_tu_core_on_the_device_keeps_the_bitstream(tmp_path, built, host, pkg, w, h, qp):
    """HEVCDL_TQ=1: transform, flat quantiser, dequantiser and inverse transform of EVERY luma and chroma TU the reference's
    xIntraCodingTUBlock codes (all RD trials included) come from hevcdl_tu_code.  With the encoder options the device core
    covers (--RDOQ=0 --RDOQTS=0 --SignHideFlag=0) the bitstream must be byte-identical to the unmodified reference run with
    the same options, and no TU may be left to HM."""
    import re
    frames = [pkg.synth.synth_frame(w, h, 90 + i) for i in range(2)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dp.predict_frame(Y, U, V, frame=f))
    dp.close()
    flat = ("--RDOQ=0", "--RDOQTS=0", "--SignHideFlag=0")
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, 2, qp, extra=flat)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 2, qp, extra=flat, env={"HEVCDL_TQ": "1", "HEVCDL_VERBOSE": "1"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-400:])
    m = re.search(r"TUs coded on the device (\d+) / left to HM (\d+)", rb["stderr"])
    assert m and int(m.group(1)) > 1000 and int(m.group(2)) == 0, rb["stderr"][-300:]
    assert ra["sha1"] == rb["sha1"] and (ra["kbps"], ra["psnr_y"]) == (rb["kbps"], rb["psnr_y"])
    # the reference's operating point (RDOQ, RDOQTS, sign-bit hiding on): every TU goes through the device RDOQ
    # (hevcdl_tu_code_rdoq fed HM's live CABAC bit-estimate tables) and the stream again equals the reference's
    rc = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 2, qp, out="rdoq.bin", env={"HEVCDL_TQ": "1", "HEVCDL_VERBOSE": "1"})
    rd = hm_util.encode("ref", str(a), "in.yuv", w, h, 2, qp, out="rdoq.bin")
    m = re.search(r"TUs coded on the device (\d+) / left to HM (\d+)", rc["stderr"])
    assert m and int(m.group(1)) > 1000 and int(m.group(2)) == 0, rc["stderr"][-300:]
    assert rc["sha1"] == rd["sha1"] and (rc["kbps"], rc["psnr_y"]) == (rd["kbps"], rd["psnr_y"])
    # mixed: RDOQ on, sign-bit hiding off
    re_ = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, 2, qp, out="m.bin", extra=("--SignHideFlag=0",), env={"HEVCDL_TQ": "1"})
    rf = hm_util.encode("ref", str(a), "in.yuv", w, h, 2, qp, out="m.bin", extra=("--SignHideFlag=0",))
    assert re_["rc"] == 0 and re_["sha1"] == rf["sha1"]


@needs_bins
@pytest.mark.gpu
def test_sidecar_accepts_sizes_that_are_not_multiples_of_8(tmp_path, built, host, oracle, weights, pkg, monkeypatch):
    """The reference sidecar takes any even picture size (PIL pads its crops with black, use_model.py:92-93).  Ours pads the
    planes to the next multiple of 8 with video black, which converts to RGB (0,0,0): the labels must equal the oracle's on
    the UNPADDED picture (whose crops are zero-padded exactly like PIL's)."""
    import importlib
    sidecar = importlib.import_module("hevc-deep-learning-pipeline_b200.sidecar")
    w, h = 100, 70
    Y, U, V = pkg.synth.synth_frame(104, 72, 5)
    Y, U, V = np.ascontiguousarray(Y[:h, :w]), np.ascontiguousarray(U[:h // 2, :w // 2]), np.ascontiguousarray(V[:h // 2, :w // 2])
    hm_util.write_yuv(str(tmp_path / "in.yuv"), [(Y, U, V)])
    (tmp_path / "bitstream.cfg").write_text(BITSTREAM_CFG.format(yuv="in.yuv", w=w, h=h, n=1))
    monkeypatch.chdir(tmp_path)
    sidecar.main(["gen_frames"])
    sidecar.main(["use_model", "--precision", "fp32"])
    olab = oracle.frame_labels(weights, Y, U, V)
    got = np.array([[int(t) for t in open(tmp_path / "pred" / "0" / ("ctu%d.txt" % i)).read().split()] for i in range(len(olab))], np.uint8)
    assert got.shape == olab.shape == (4, 16) and (got == olab).all()


@needs_bins
@pytest.mark.gpu
def test_dropin_every_device_component_at_once(tmp_path, built, host, pkg):
    """Labels (CNN, fp32), frame lookahead, exact first-pass SATDs from HM's reconstructed references (HEVCDL_RMD=2), transform +
    RDOQ + dequantiser + inverse transform of every TU (HEVCDL_TQ=1), every intra-predicted block of the RD pass (HEVCDL_PRED=1),
    the deblocking filter (HEVCDL_DBF=1) and the SAO statistics (HEVCDL_SAO=1) all on the B200
    inside one encode at the reference's default options: the bitstream must still be byte-identical to the unmodified
    reference encoder's, and the stream must decode with matching picture hashes."""
    import re
    w, h, n, qp = 192, 128, 3, 32
    frames = [pkg.synth.synth_frame(w, h, 110 + i) for i in range(n)]
    a, b = tmp_path / "ref", tmp_path / "dl"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        hm_util.write_yuv(str(d / "in.yuv"), frames)
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(str(a / "pred"), f, dp.predict_frame(Y, U, V, frame=f))
    dp.close()
    ra = hm_util.encode("ref", str(a), "in.yuv", w, h, n, qp)
    rb = hm_util.encode("hevcdl", str(b), "in.yuv", w, h, n, qp,
                        env={"HEVCDL_PRECISION": "fp32", "HEVCDL_RMD": "2", "HEVCDL_TQ": "1", "HEVCDL_DBF": "1", "HEVCDL_SAO": "1",
                             "HEVCDL_PRED": "1", "HEVCDL_VERBOSE": "1"})
    assert ra["rc"] == 0 and rb["rc"] == 0, (ra["stderr"][-400:], rb["stderr"][-600:])
    err = rb["stderr"]
    assert int(re.search(r"exact PU calls (\d+)", err).group(1)) > 100
    m = re.search(r"TUs coded on the device (\d+) / left to HM (\d+)", err)
    assert int(m.group(1)) > 1000 and int(m.group(2)) == 0
    assert re.search(r"pictures deblocked on the device %d / by the reference's filter 0" % n, err)
    assert re.search(r"SAO statistics passes on the device %d / by the reference's code 0" % n, err)
    assert re.search(r"SAO offsets applied on the device for %d pictures / by the reference's code for 0" % n, err)
    assert re.search(r"in-loop passes of %d pictures shared one upload" % n, err)
    m = re.search(r"blocks predicted on the device (\d+) / by the reference's code (\d+)", err)
    assert int(m.group(1)) > 5000 and int(m.group(2)) == 0
    assert re.search(r"lookahead 3: %d frames were on the device before HM asked" % (n - 1), err)
    assert ra["sha1"] == rb["sha1"]
    ok, out = hm_util.decode_ok(str(b))
    assert ok, out[-400:]
