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
                out["input"] = val
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
    fbytes = w * h * 3 // 2
    nfile = os.path.getsize(cfg["input"]) // fbytes
    nframes = min(cfg["frames"], nfile)                    # use_model.py:73-76: stop at FramesToBeEncoded
    prec = host.PREC_BF16_TC if args.precision == "bf16" else host.PREC_FP32
    depth = max(2, 2 * args.batch)
    dp = host.DepthPredictor(w, h, device=args.device, slots=depth, precision=prec, rmd=False, batch=args.batch)
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
