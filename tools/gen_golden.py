#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF in this
container (it cannot travel to the GPU box, the fixtures can).  Needs /root/reference and
oracle/_ref (make -C oracle ref).  Re-run only when the oracle's input definition changes.

  cnn_labels_416x240.npz   labels written by the UNMODIFIED use_model.py (cwd = temp dir holding
                           ./rec/frames/1.jpg -- PNG bytes of our staged RGB, PIL sniffs the
                           format so the lossy JPEG stage drops out -- ./rec/*.pt, bitstream.cfg)
  cnn_logits.npz           ConvNet2 (class text of use_model.py:16-58, train-mode BN, batch 1)
                           logits for seeded CTUs
  rmd_trace_192x128_qp32.npz   per-PU 35x(SAD, mode bits) printed by oracle/_ref/TAppEncoder_trace
                           (DEBUG_INTRA_SEARCH_COSTS) + its unfiltered reconstruction + labels
"""
import importlib
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
from oracle import oracle  # noqa: E402


def staged_rgb_image(Y, U, V):
    """Whole-picture RGB (H,W,3) through the oracle's K0 definition."""
    H, W = Y.shape
    cw, ch = (W + 63) // 64, (H + 63) // 64
    img = np.zeros((ch * 64, cw * 64, 3), np.uint8)
    for cy in range(ch):
        for cx in range(cw):
            t = oracle.stage_ctu_rgb(Y, U, V, cx, cy)
            img[cy * 64:(cy + 1) * 64, cx * 64:(cx + 1) * 64] = t.transpose(1, 2, 0)
    return img[:H, :W]


def run_use_model(Y, U, V):
    """Run the unmodified reference sidecar on one frame; return labels [nctu,16]."""
    from PIL import Image
    H, W = Y.shape
    d = tempfile.mkdtemp(prefix="hevcdl_gold_")
    try:
        os.makedirs(os.path.join(d, "rec", "frames"))
        os.makedirs(os.path.join(d, "pred"))
        os.symlink(os.path.join(REF, "rec", "hevc_encoder_model.pt"), os.path.join(d, "rec", "hevc_encoder_model.pt"))
        Image.fromarray(staged_rgb_image(Y, U, V), "RGB").save(os.path.join(d, "rec", "frames", "1.jpg"), format="PNG")
        with open(os.path.join(d, "bitstream.cfg"), "w") as f:   # use_model.py:65-71 reads line 7
            f.write("InputFile : x.yuv\nInputBitDepth : 8\nInputChromaFormat : 420\nFrameRate : 30\n"
                    "FrameSkip : 0\nSourceWidth : %d\nSourceHeight : %d\nFramesToBeEncoded : 1\n" % (W, H))
        subprocess.check_call([sys.executable, os.path.join(REF, "use_model.py")], cwd=d,
                              stdout=subprocess.DEVNULL)
        nctu = ((W + 63) // 64) * ((H + 63) // 64)
        lab = np.zeros((nctu, 16), np.uint8)
        for i in range(nctu):
            lab[i] = [int(t) for t in open(os.path.join(d, "pred", "0", "ctu%d.txt" % i)).read().split()]
        return lab
    finally:
        shutil.rmtree(d)


def load_convnet2():
    """exec only the class definition of use_model.py (the script body has side effects)."""
    src = open(os.path.join(REF, "use_model.py"), encoding="utf-8").read()
    head = src.split("DEVICE = ")[0]
    ns = {}
    exec(compile(head, "use_model_head", "exec"), ns)
    import torch
    m = ns["ConvNet2"]()
    m.load_state_dict(torch.load(os.path.join(REF, "rec", "hevc_encoder_model.pt"), map_location="cpu"))
    return m  # NOTE: deliberately NOT .eval() -- the reference never calls it


def gen_cnn():
    import torch
    Y, U, V = pkg.synth.synth_frame(416, 240, 0)
    lab = run_use_model(Y, U, V)
    np.savez_compressed(os.path.join(GOLD, "cnn_labels_416x240.npz"), Y=Y, U=U, V=V, labels=lab)
    print("cnn_labels_416x240: label histogram", np.bincount(lab.ravel(), minlength=4))

    # logits for seeded CTUs: 24 CTUs of a 1080p synthetic frame + 4 noise + 2 flat + 2 partial
    m = load_convnet2()
    tiles = []
    Yb, Ub, Vb = pkg.synth.synth_frame(1920, 1080, 0)
    rng = np.random.default_rng(7)
    for a in rng.choice(30 * 17, 24, replace=False):
        tiles.append(oracle.stage_ctu_rgb(Yb, Ub, Vb, int(a % 30), int(a // 30)))
    for a in (29 + 16 * 30, 5 + 16 * 30):      # bottom partial row
        tiles.append(oracle.stage_ctu_rgb(Yb, Ub, Vb, a % 30, a // 30))
    Yn, Un, Vn = pkg.synth.synth_frame(128, 128, 3, "noise")
    for a in range(4):
        tiles.append(oracle.stage_ctu_rgb(Yn, Un, Vn, a % 2, a // 2))
    Yf, Uf, Vf = pkg.synth.synth_frame(128, 64, 0, "flat")
    for a in range(2):
        tiles.append(oracle.stage_ctu_rgb(Yf, Uf, Vf, a, 0))
    tiles = np.stack(tiles)
    logits = np.zeros((len(tiles), 4, 16), np.float32)
    with torch.no_grad():
        for i, t in enumerate(tiles):
            x64 = torch.from_numpy(t.astype(np.float32) / np.float32(255.0))[None]
            for q in range(4):
                oy, ox = (q // 2) * 32, (q % 2) * 32
                logits[i, q] = m(x64[:, :, oy:oy + 32, ox:ox + 32].contiguous(), x64)[0].numpy()
    np.savez_compressed(os.path.join(GOLD, "cnn_logits.npz"), rgb64=tiles, logits=logits)
    print("cnn_logits:", tiles.shape, logits.shape)


def gen_rmd(W=192, H=128, qp=32):
    Yb, Ub, Vb = pkg.synth.synth_frame(1920, 1080, 0)
    y0, x0 = 256, 512
    Y = np.ascontiguousarray(Yb[y0:y0 + H, x0:x0 + W])
    U = np.ascontiguousarray(Ub[y0 // 2:(y0 + H) // 2, x0 // 2:(x0 + W) // 2])
    V = np.ascontiguousarray(Vb[y0 // 2:(y0 + H) // 2, x0 // 2:(x0 + W) // 2])
    w = oracle.load_weights(os.path.join(ROOT, "weights", "hevc_encoder_model.hdlw"))
    labels = oracle.frame_labels(w, Y, U, V)
    # make sure every CU size occurs: force CTU 1 to one 64x64 CU and CTU 2's first quadrant to 32x32
    labels[1, :] = 0
    labels[2, [0, 1, 4, 5]] = 1
    labels[2][labels[2] == 0] = 1
    d = tempfile.mkdtemp(prefix="hevcdl_rmd_")
    try:
        os.makedirs(os.path.join(d, "pred", "0"))
        for i, l in enumerate(labels):
            open(os.path.join(d, "pred", "0", "ctu%d.txt" % i), "w").write("".join("%d " % v for v in l))
        open(os.path.join(d, "in.yuv"), "wb").write(pkg.synth.to_i420_bytes(Y, U, V))
        cmd = [os.path.join(ROOT, "oracle", "_ref", "TAppEncoder_trace"),
               "-c", os.path.join(REF, "HM_dl", "cfg", "encoder_intra_main.cfg"),
               "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", "1", "-q", str(qp),
               "-b", "out.bin", "-o", "rec.yuv", "--LoopFilterDisable=1", "--SAO=0",
               "--SEIDecodedPictureHash=1"]
        out = subprocess.run(cmd, cwd=d, capture_output=True, text=True).stdout
        rec = np.frombuffer(open(os.path.join(d, "rec.yuv"), "rb").read(), np.uint8)[:W * H].reshape(H, W).copy()
    finally:
        shutil.rmtree(d)
    sad, bits, cands = [], [], []
    for l in out.splitlines():
        m = re.match(r"1st pass mode (\d+) SAD = (\d+), mode bits = (\d+),", l)
        if m:
            if int(m.group(1)) == 0:
                sad.append([]); bits.append([]); cands.append([])
            sad[-1].append(int(m.group(2))); bits[-1].append(int(m.group(3)))
            continue
        m = re.match(r"2nd pass \[luma,chroma\] mode \[(\d+),", l)
        if m:
            cands[-1].append(int(m.group(1)))
    sad = np.array(sad, np.uint32); bits = np.array(bits, np.uint32)
    assert sad.shape[1] == 35 and len(cands) == len(sad)
    ncand = np.array([len(c) for c in cands], np.int32)
    cand = np.full((len(cands), 10), 255, np.uint8)
    for i, c in enumerate(cands):
        cand[i, :len(c)] = c
    summary = [l for l in out.splitlines() if l.startswith("POC")]
    print("rmd_trace: %d PUs; %s" % (sad.shape[0], summary[:1]))
    np.savez_compressed(os.path.join(GOLD, "rmd_trace_%dx%d_qp%d.npz" % (W, H, qp)), Y=Y, U=U, V=V, rec=rec,
                        labels=labels, sad=sad, bits=bits, cand=cand, ncand=ncand, qp=np.int32(qp))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1:] or ["cnn", "rmd"]
    if "cnn" in what:
        gen_cnn()
    if "rmd" in what:
        gen_rmd()
