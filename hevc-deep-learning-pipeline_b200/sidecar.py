"""Drop-in for the reference's two sidecar scripts, for an UNMODIFIED TAppEncoder (file handshake).

The reference encoder runs `python gen_frames.py` (ffmpeg frame dump + reset of ./pred, gen_frames.py:4-27) and then,
in a detached thread, `python use_model.py` (per-CTU labels to ./pred/<frame>/ctu<i>.txt, use_model.py:65-127), and
polls the file system for each CTU's file (HM_dl/source/Lib/TLibEncoder/TEncCu.cpp:243-253).  Put two one-line
scripts of those names next to the encoder that call

    python -m hevc-deep-learning-pipeline_b200.sidecar gen_frames
    python -m hevc-deep-learning-pipeline_b200.sidecar use_model [--precision fp32|bf16] [--batch N]

and the same encoder binary gets its labels from the B200: same working-directory contract (bitstream.cfg parsed by line
index, ./pred/<0-based frame>/ctu<raster address>.txt, sixteen digits each followed by a space, published by rename),
no ffmpeg, no JPEG (frames are read from the YUV file named in bitstream.cfg), no torch.  There is no CPU path: without
libhevcdl.so or a B200 use_model fails, as the reference does when its model file is missing.
"""
import argparse
import os
import shutil
import sys

import numpy as np

from . import host


def parse_bitstream_cfg(path="bitstream.cfg"):
    """The reference reads bitstream.cfg BY LINE INDEX (gen_frames.py:4-16, use_model.py:65-71): line 0 InputFile,
    3 FrameRate, 5 SourceWidth, 6 SourceHeight, 7 FramesToBeEncoded; everything after the first ':' is the value."""
    out = {}
    with open(path, "r") as f:
        for i, line in enumerate(f):
            parts = line.split(":")
            val = ":".join(parts[1:]).strip(" ").strip("\n").strip()
            if i == 0:
                # the reference's own bitstream.cfg names the file Windows-style (".\\Flowervase_...yuv"); the same file
                # must open here: drop a leading ".\\" and turn the remaining separators round
                out["input"] = val[2:].replace("\\", "/") if val.startswith(".\\") else val.replace("\\", "/")
            elif i == 3:
                out["frame_rate"] = val
            elif i == 5:
                out["width"] = int(val)
            elif i == 6:
                out["height"] = int(val)
            elif i == 7:
                out["frames"] = int(val)
    return out


def gen_frames(_args):
    """gen_frames.py without the JPEG dump: only its second job, a fresh ./pred (gen_frames.py:23-27)."""
    parse_bitstream_cfg()                                  # fail like the reference if the file is missing
    shutil.rmtree("./pred", ignore_errors=True)
    os.mkdir("./pred")


def use_model(args):
    cfg = parse_bitstream_cfg()
    w, h = cfg["width"], cfg["height"]
    if w <= 0 or h <= 0 or w % 2 or h % 2:
        raise SystemExit("hevcdl sidecar: 4:2:0 needs even SourceWidth/SourceHeight, got %dx%d" % (w, h))
    fbytes = w * h * 3 // 2
    if not os.path.exists(cfg["input"]):
        # fail BEFORE the encoder starts polling ./pred (HM_dl TEncCu.cpp:245 never times out)
        raise SystemExit("hevcdl sidecar: InputFile %r of bitstream.cfg not found" % cfg["input"])
    nfile = os.path.getsize(cfg["input"]) // fbytes
    if nfile < cfg["frames"]:
        raise SystemExit("hevcdl sidecar: %s holds %d frames of %dx%d, bitstream.cfg asks for %d -- the encoder would wait "
                         "for ./pred/%d forever" % (cfg["input"], nfile, w, h, cfg["frames"], nfile))
    nframes = cfg["frames"]                                # use_model.py:73-76: stop at FramesToBeEncoded
    prec = host.PREC_BF16_TC if args.precision == "bf16" else host.PREC_FP32
    depth = max(2, 2 * args.batch)
    # The library takes multiples of 8 (HM's minimum CU, TAppEncCfg.cpp:2176).  The reference sidecar accepts any size:
    # it crops the UNPADDED picture and PIL pads with black (use_model.py:92-93).  Video black (Y=16, Cb=Cr=128) converts
    # to RGB (0,0,0) exactly (csrc/common.cuh yuv2rgb), so padding the planes with it up to the next multiple of 8 gives
    # the CNN the very tensor the reference would see; the CTU count ceil(W/64) x ceil(H/64) does not change.
    w8, h8 = (w + 7) // 8 * 8, (h + 7) // 8 * 8
    dp = host.DepthPredictor(w8, h8, device=args.device, slots=depth, precision=prec, rmd=False, batch=args.batch, outputs=0)
    yuv = np.memmap(cfg["input"], np.uint8, "r")
    inflight = []

    def publish(f):
        os.mkdir("./pred/%d" % f)                          # use_model.py:77 (0-based directory)
        dp.write_pred_files(dp.labels(f), "./pred", f)     # use_model.py:121-125: temp name, then rename
        dp.release(f)

    # all-intra frames are independent: with --world G each of G processes (one per GPU) takes frames f = rank mod G
    # and reads them at byte offset f * W * H * 3/2; they all publish into the same ./pred
    for f in host.rank_frames(nframes, args.rank, args.world):
        fr = yuv[f * fbytes:(f + 1) * fbytes]
        Y = fr[:w * h].reshape(h, w)
        U = fr[w * h:w * h * 5 // 4].reshape(h // 2, w // 2)
        V = fr[w * h * 5 // 4:].reshape(h // 2, w // 2)
        if (w8, h8) != (w, h):
            Y = np.pad(Y, ((0, h8 - h), (0, w8 - w)), constant_values=16)
            U = np.pad(U, ((0, (h8 - h) // 2), (0, (w8 - w) // 2)), constant_values=128)
            V = np.pad(V, ((0, (h8 - h) // 2), (0, (w8 - w) // 2)), constant_values=128)
        dp.submit(f, Y, U, V)
        inflight.append(f)
        if len(inflight) >= depth:
            publish(inflight.pop(0))
    for f in inflight:
        publish(f)
    dp.close()
    print("hevcdl sidecar rank %d/%d: %d of %d frames, %d CTUs each" % (args.rank, args.world, len(host.rank_frames(nframes, args.rank, args.world)), nframes, dp.nctu))


def main(argv=None):
    ap = argparse.ArgumentParser(prog="hevc-deep-learning-pipeline_b200.sidecar")
    sub = ap.add_subparsers(dest="cmd", required=True)
    sub.add_parser("gen_frames")
    um = sub.add_parser("use_model")
    um.add_argument("--precision", default=os.environ.get("HEVCDL_PRECISION", "fp32"), choices=["fp32", "bf16"])
    um.add_argument("--batch", type=int, default=1)
    um.add_argument("--device", type=int, default=int(os.environ.get("HEVCDL_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    um.add_argument("--rank", type=int, default=int(os.environ.get("RANK", "0")))
    um.add_argument("--world", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    args = ap.parse_args(argv)
    (gen_frames if args.cmd == "gen_frames" else use_model)(args)


if __name__ == "__main__":
    main()
