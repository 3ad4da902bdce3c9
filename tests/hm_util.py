"""Helpers that drive HM encoder/decoder binaries in a scratch directory (test infrastructure).

Binaries: oracle/_ref/TAppEncoder_ref (the UNMODIFIED reference, file handshake), TAppEncoder_anchor
(pruning off = stock HM), TAppDecoder_ref, and hm_plugin/_build/TAppEncoder_hevcdl (the drop-in:
reference sources + this repo's TEncCu::compressCtu over libhevcdl.so).  None of them is rebuilt on
the GPU box; they travel with the snapshot.
"""
import hashlib
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "tests", "data", "intra_main_b200.cfg")
BIN = {
    "ref": os.path.join(ROOT, "oracle", "_ref", "TAppEncoder_ref"),
    "anchor": os.path.join(ROOT, "oracle", "_ref", "TAppEncoder_anchor"),
    "dec": os.path.join(ROOT, "oracle", "_ref", "TAppDecoder_ref"),
    "hevcdl": os.path.join(ROOT, "hm_plugin", "_build", "TAppEncoder_hevcdl"),
}


def have(*names):
    return all(os.path.exists(BIN[n]) for n in names)


def write_yuv(path, frames):
    with open(path, "wb") as f:
        for (Y, U, V) in frames:
            f.write(np.ascontiguousarray(Y, np.uint8).tobytes())
            f.write(np.ascontiguousarray(U, np.uint8).tobytes())
            f.write(np.ascontiguousarray(V, np.uint8).tobytes())


def write_pred(pred_dir, frame, labels):
    """The reference's handshake files (use_model.py:121-125): 16 digits each followed by a space."""
    d = os.path.join(pred_dir, str(frame))
    os.makedirs(d, exist_ok=True)
    for i, l in enumerate(labels):
        with open(os.path.join(d, "ctu%d.txt" % i), "w") as f:
            f.write("".join("%d " % int(v) for v in l))


_SUMMARY = re.compile(r"^\s*(\d+)\s+a\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", re.M)
_TIME = re.compile(r"Total Time:\s*([\d.]+)\s*sec")


def encode(kind, cwd, yuv, w, h, nframes, qp, out="str.bin", env=None, extra=()):
    """Run one encoder from `cwd` (which holds ./pred for the file-handshake builds).  Returns a dict
    with the bitstream bytes' sha1, the summary line (kbps, Y/U/V/YUV PSNR) and HM's own Total Time."""
    cmd = [BIN[kind], "-c", CFG, "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(nframes),
           "-q", str(qp), "-b", out, "--SEIDecodedPictureHash=1", "--InputBitDepth=8", "--InputChromaFormat=420",
           "--Level=6.2"] + list(extra)
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=3600)
    res = {"rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr}
    bs = os.path.join(cwd, out)
    if p.returncode == 0 and os.path.exists(bs):
        data = open(bs, "rb").read()
        res["sha1"] = hashlib.sha1(data).hexdigest()
        res["bytes"] = len(data)
        m = _SUMMARY.search(p.stdout)
        if m:
            res["kbps"], res["psnr_y"], res["psnr_u"], res["psnr_v"], res["psnr_yuv"] = (float(m.group(i)) for i in range(2, 7))
        t = _TIME.search(p.stdout)
        if t:
            res["seconds"] = float(t.group(1))
    return res


def decode_ok(cwd, bitstream="str.bin"):
    """Decode with the reference decoder; True iff every picture's MD5 SEI matched (conformance of the encode)."""
    p = subprocess.run([BIN["dec"], "-b", bitstream, "-o", "dec.yuv"], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=3600)
    out = p.stdout
    return p.returncode == 0 and "(OK)" in out and "ERROR" not in out and "mismatch" not in out.lower(), out
