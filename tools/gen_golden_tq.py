#!/usr/bin/env python
"""Golden vectors of the transform / quantisation core, made by RUNNING THE REFERENCE in this
container (needs /root/reference and `make -C oracle ref`); the fixtures travel, the reference does not.

  tests/golden/tq_transform_ref.npz   random and extreme blocks through the reference's OWN xTrMxN / xITrMxN
                                      (oracle/_ref/libtqref.so = tq_ref_harness.cpp linked with libhmref.a), every size + DST
  tests/golden/pred_trace_192x128_qp32.npz  reference samples + predicted blocks of the reference's own predIntraAng (luma and chroma)
  tests/golden/sao_apply.npz          deblocked pictures + per-CTU SAO parameters + the pictures after the reference's own offsetCTU
  tests/golden/sao_stats.npz          original + deblocked pictures and the statistics the reference's own SAO getStatistics made of them
  tests/golden/dbf_pictures.npz       reconstructed pictures before / after the reference's own deblocking filter + its TU / QP maps
  tests/golden/tq_rdoq_192x128_qp32.npz    calls of the reference's xRateDistOptQuant (inputs incl. the CABAC bit-estimate
                                      tables, outputs) dumped by oracle/_ref/TAppEncoder_rdoqtrace at the default options
  tests/golden/tq_trace_192x128_qp32.npz   per-TU dumps printed by oracle/_ref/TAppEncoder_tqtrace (the reference built with
                                      its own DEBUG_TRANSFORM_AND_QUANTISE switch, TComTrQuant.cpp:1496-1662) while encoding
                                      the 192x128 fixture frame at QP 32 with --RDOQ=0 --RDOQTS=0 --SignHideFlag=0 (the flat
                                      quantiser xQuant, :1126-1249): residual, transform output, levels and, where the
                                      reference ran the inverse path, dequantised coefficients and reconstructed residual.
"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hm_util  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def transform_vectors():
    ref = C.CDLL(os.path.join(REFDIR, "libtqref.so"))
    ref.tqref_init()
    ip = np.ctypeslib.ndpointer(np.int32, flags="C")
    ref.tqref_forward.argtypes = [ip, ip, C.c_int, C.c_int]
    ref.tqref_inverse.argtypes = [ip, ip, C.c_int, C.c_int]
    rng = np.random.default_rng(20261017)
    sizes, dsts, resi, coeff, icoeff, iresi = [], [], [], [], [], []
    for N, cnt in ((4, 24), (8, 20), (16, 12), (32, 8)):
        for dst in ((0, 1) if N == 4 else (0,)):
            for t in range(cnt):
                kind = t % 4
                if kind == 0:
                    blk = rng.integers(-255, 256, (N, N))
                elif kind == 1:
                    blk = rng.integers(-12, 13, (N, N))
                elif kind == 2:
                    blk = np.full((N, N), 255 if t % 8 < 4 else -255)
                else:
                    blk = rng.integers(0, 2, (N, N)) * 510 - 255
                b = np.ascontiguousarray(blk, np.int32)
                c = np.zeros((N, N), np.int32)
                ref.tqref_forward(b.copy(), c, N, dst)
                ci = np.ascontiguousarray(rng.integers(-32768, 32768, (N, N)) if kind != 1 else rng.integers(-400, 400, (N, N)), np.int32)
                r = np.zeros((N, N), np.int32)
                ref.tqref_inverse(ci.copy(), r, N, dst)
                assert np.abs(r).max() <= 32768
                sizes.append(N); dsts.append(dst)
                resi.append(b.astype(np.int16).ravel()); coeff.append(c.ravel())
                icoeff.append(ci.ravel()); iresi.append(r.astype(np.int16).ravel())
    off = np.concatenate([[0], np.cumsum([s * s for s in sizes])]).astype(np.int64)
    np.savez_compressed(os.path.join(GOLD, "tq_transform_ref.npz"), sizes=np.array(sizes, np.uint8), dst=np.array(dsts, np.uint8), off=off,
                        resi=np.concatenate(resi), coeff=np.concatenate(coeff), icoeff=np.concatenate(icoeff), iresi=np.concatenate(iresi))
    print("tq_transform_ref.npz:", len(sizes), "blocks")


HDR = re.compile(r"^\d+: (\d+)x(\d+) channel (\d) TU (at input to transform|between transform and quantiser|at output of quantiser|"
                 r"at input to dequantiser|between dequantiser and inverse-transform|at output of inverse-transform)$")
STAGE = {"at input to transform": "resi", "between transform and quantiser": "coeff", "at output of quantiser": "level",
         "at input to dequantiser": "level2", "between dequantiser and inverse-transform": "deq", "at output of inverse-transform": "rec"}


def parse_trace(path):
    recs, cur = [], None
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        m = HDR.match(lines[i])
        if not m:
            i += 1
            continue
        n, ch, st = int(m.group(1)), int(m.group(3)), STAGE[m.group(4)]
        blk = np.array([[int(v) for v in lines[i + 1 + r].split()] for r in range(n)], np.int64)
        assert blk.shape == (n, n), (i, blk.shape)
        i += 1 + n
        if st == "resi":
            cur = {"n": n, "ch": ch, "resi": blk}
            recs.append(cur)
        else:
            assert cur is not None and cur["n"] == n and cur["ch"] == ch, (i, st)
            cur[st] = blk
    return recs


def trace_vectors(per_class=14):
    g = np.load(os.path.join(GOLD, "rmd_trace_192x128_qp32.npz"))
    Y, U, V = g["Y"], g["U"], g["V"]
    H, W = Y.shape
    qp = int(g["qp"])
    with tempfile.TemporaryDirectory() as td:
        hm_util.write_yuv(os.path.join(td, "in.yuv"), [(Y, U, V)])
        hm_util.write_pred(os.path.join(td, "pred"), 0, g["labels"])
        cmd = [os.path.join(REFDIR, "TAppEncoder_tqtrace"), "-c", hm_util.CFG, "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
               "-f", "1", "-q", str(qp), "-b", "t.bin", "--InputBitDepth=8", "--InputChromaFormat=420", "--Level=6.2", "--RDOQ=0",
               "--RDOQTS=0", "--SignHideFlag=0"]
        with open(os.path.join(td, "trace.txt"), "w") as f:
            subprocess.check_call(cmd, cwd=td, stdout=f, stderr=subprocess.DEVNULL)
        recs = parse_trace(os.path.join(td, "trace.txt"))
    print("trace:", len(recs), "TUs")
    # chroma QP of the reference for luma QP 32, ChromaQpOffset 0, 4:2:0 (TComRom.cpp g_aucChromaScale): 31
    qpc = {32: 31}[qp]
    rng = np.random.default_rng(7)
    pick = []
    classes = {}
    for k, r in enumerate(recs):
        if "level" not in r:
            continue
        tskip = r["n"] == 4 and (r["coeff"] == r["resi"] * 32).all() and r["resi"].any()
        key = (r["n"], r["ch"], "rec" in r, bool(tskip))
        classes.setdefault(key, []).append(k)
    for key, idx in sorted(classes.items()):
        sel = rng.permutation(idx)[:per_class if key[2] else 4]
        pick += [(int(k), key) for k in sel]
        print(key, len(idx), "->", len(sel))
    sizes, chans, qps, flags, has_inv = [], [], [], [], []
    resi, coeff, level, deq, rec = [], [], [], [], []
    for k, key in pick:
        r = recs[k]
        n, ch = r["n"], r["ch"]
        sizes.append(n); chans.append(ch); qps.append(qp if ch == 0 else qpc)
        flags.append((2 if key[3] else (1 if (n == 4 and ch == 0) else 0)))      # oracle/tq_oracle.c TQ_FLAG_*: DST for 4x4 luma, TSKIP
        has_inv.append(1 if "rec" in r else 0)
        resi.append(r["resi"].astype(np.int16).ravel()); coeff.append(r["coeff"].astype(np.int32).ravel())
        level.append(r["level"].astype(np.int32).ravel())
        if "rec" in r:
            assert (r["level2"] == r["level"]).all()
            deq.append(r["deq"].astype(np.int32).ravel()); rec.append(r["rec"].astype(np.int16).ravel())
        else:
            deq.append(np.zeros(n * n, np.int32)); rec.append(np.zeros(n * n, np.int16))
    off = np.concatenate([[0], np.cumsum([s * s for s in sizes])]).astype(np.int64)
    np.savez_compressed(os.path.join(GOLD, "tq_trace_192x128_qp32.npz"), sizes=np.array(sizes, np.uint8), chan=np.array(chans, np.uint8),
                        qp=np.array(qps, np.uint8), flags=np.array(flags, np.uint8), has_inv=np.array(has_inv, np.uint8), off=off,
                        resi=np.concatenate(resi), coeff=np.concatenate(coeff), level=np.concatenate(level), deq=np.concatenate(deq),
                        rec=np.concatenate(rec))
    print("tq_trace_192x128_qp32.npz:", len(sizes), "TUs,", os.path.getsize(os.path.join(GOLD, "tq_trace_192x128_qp32.npz")), "bytes")


def rdoq_vectors(per_class=7):
    """tests/golden/tq_rdoq_192x128_qp32.npz: inputs and outputs of calls of the reference's xRateDistOptQuant, dumped by
    oracle/_ref/TAppEncoder_rdoqtrace (oracle/rdoq_dump.h) while encoding the fixture frame at the reference's operating
    point (RDOQ, RDOQTS, sign-bit hiding on): a sample of every (size, component, scan type, transform skip) class."""
    import struct
    g = np.load(os.path.join(GOLD, "rmd_trace_192x128_qp32.npz"))
    Y, U, V = g["Y"], g["U"], g["V"]
    H, W = Y.shape
    qp = int(g["qp"])
    with tempfile.TemporaryDirectory() as td:
        hm_util.write_yuv(os.path.join(td, "in.yuv"), [(Y, U, V)])
        hm_util.write_pred(os.path.join(td, "pred"), 0, g["labels"])
        cmd = [os.path.join(REFDIR, "TAppEncoder_rdoqtrace"), "-c", hm_util.CFG, "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
               "-f", "1", "-q", str(qp), "-b", "t.bin", "--InputBitDepth=8", "--InputChromaFormat=420", "--Level=6.2"]
        env = dict(os.environ, HEVCDL_RDOQ_DUMP=os.path.join(td, "rdoq.bin"))
        subprocess.check_call(cmd, cwd=td, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        raw = open(os.path.join(td, "rdoq.bin"), "rb").read()
    off, recs = 0, []
    while off < len(raw):
        hdr = np.frombuffer(raw, np.int32, 16, off).copy(); off += 64
        assert hdr[0] == 0x52444F51
        lam, es = struct.unpack_from("dd", raw, off); off += 16
        est = np.frombuffer(raw, np.int32, hdr[12] // 4, off).copy(); off += int(hdr[12])
        n2 = int(hdr[1]) ** 2
        src = np.frombuffer(raw, np.int32, n2, off).copy(); off += 4 * n2
        dst = np.frombuffer(raw, np.int32, n2, off).copy(); off += 4 * n2
        a = int(np.frombuffer(raw, np.int32, 1, off)[0]); off += 4
        recs.append((hdr, lam, es, est, src, dst, a))
    print("rdoq trace:", len(recs), "calls")
    rng = np.random.default_rng(9)
    classes = {}
    for k, r in enumerate(recs):
        classes.setdefault((int(r[0][1]), int(r[0][2]), int(r[0][6]), int(r[0][7]), r[6] > 0), []).append(k)
    pick = []
    for key, idx in sorted(classes.items()):
        pick += [int(k) for k in rng.permutation(idx)[:per_class if key[4] else 2]]
    rs = [recs[k] for k in pick]
    off = np.concatenate([[0], np.cumsum([int(r[0][1]) ** 2 for r in rs])]).astype(np.int64)
    np.savez_compressed(os.path.join(GOLD, "tq_rdoq_192x128_qp32.npz"), hdr=np.stack([r[0] for r in rs]), lam=np.array([r[1] for r in rs]),
                        err_scale=np.array([r[2] for r in rs]), est=np.stack([r[3] for r in rs]), off=off, src=np.concatenate([r[4] for r in rs]),
                        dst=np.concatenate([r[5] for r in rs]), abs_sum=np.array([r[6] for r in rs], np.int64))
    print("tq_rdoq_192x128_qp32.npz:", len(rs), "calls,", os.path.getsize(os.path.join(GOLD, "tq_rdoq_192x128_qp32.npz")), "bytes")


def dbf_vectors():
    """tests/golden/dbf_pictures.npz: reconstructed pictures right before and right after the reference's own deblocking filter
    (oracle/_ref/TAppEncoder_dbftrace, oracle/dbf_dump.h) with the TU-size and QP maps it reads: the 192x128 fixture frame at
    QP 22 and 37 and the 416x240 frame (partial CTUs on both edges) at QP 32."""
    cases = []
    g = np.load(os.path.join(GOLD, "rmd_trace_192x128_qp32.npz"))
    cases += [(g["Y"], g["U"], g["V"], g["labels"], 22), (g["Y"], g["U"], g["V"], g["labels"], 37)]
    c = np.load(os.path.join(GOLD, "cnn_labels_416x240.npz"))
    cases.append((c["Y"], c["U"], c["V"], c["labels"], 32))
    out = {}
    for k, (Y, U, V, labels, qp) in enumerate(cases):
        H, W = Y.shape
        with tempfile.TemporaryDirectory() as td:
            hm_util.write_yuv(os.path.join(td, "in.yuv"), [(Y, U, V)])
            hm_util.write_pred(os.path.join(td, "pred"), 0, labels)
            cmd = [os.path.join(REFDIR, "TAppEncoder_dbftrace"), "-c", hm_util.CFG, "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
                   "-f", "1", "-q", str(qp), "-b", "t.bin", "--InputBitDepth=8", "--InputChromaFormat=420", "--Level=6.2"]
            subprocess.check_call(cmd, cwd=td, env=dict(os.environ, HEVCDL_DBF_DUMP=os.path.join(td, "dbf.bin")), stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
            raw = open(os.path.join(td, "dbf.bin"), "rb").read()
        off = 0
        for phase in (0, 1):
            hdr = np.frombuffer(raw, np.int32, 12, off).copy(); off += 48
            assert hdr[0] == 0x44424630 and hdr[1] == phase and hdr[2] == W and hdr[3] == H and hdr[9] == 1
            for name, n in (("Y", W * H), ("U", W * H // 4), ("V", W * H // 4)):
                out["%s%d_%d" % (name, phase, k)] = np.frombuffer(raw, np.int16, n, off).astype(np.uint8); off += 2 * n
            if phase == 0:
                out["hdr_%d" % k] = hdr
                out["tu_%d" % k] = np.frombuffer(raw, np.uint8, (W // 4) * (H // 4), off).copy(); off += (W // 4) * (H // 4)
                out["qp_%d" % k] = np.frombuffer(raw, np.int8, (W // 4) * (H // 4), off).copy(); off += (W // 4) * (H // 4)
        print("case", k, W, H, qp, "samples changed by the reference's filter:", int((out["Y0_%d" % k] != out["Y1_%d" % k]).sum()))
    out["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, "dbf_pictures.npz"), **out)
    print("dbf_pictures.npz:", os.path.getsize(os.path.join(GOLD, "dbf_pictures.npz")), "bytes")


def sao_vectors():
    """tests/golden/sao_stats.npz: inputs (original and deblocked pictures) and output (per CTU / component / SAO type class
    statistics) of the reference's own TEncSampleAdaptiveOffset::getStatistics (oracle/_ref/TAppEncoder_saotrace,
    oracle/sao_dump.h): the 192x128 fixture frame and the 416x240 frame (interior CTUs and partial CTUs on both edges), QP 32."""
    g = np.load(os.path.join(GOLD, "rmd_trace_192x128_qp32.npz"))
    c = np.load(os.path.join(GOLD, "cnn_labels_416x240.npz"))
    out = {}
    for k, (Y, U, V, labels) in enumerate(((g["Y"], g["U"], g["V"], g["labels"]), (c["Y"], c["U"], c["V"], c["labels"]))):
        H, W = Y.shape
        with tempfile.TemporaryDirectory() as td:
            hm_util.write_yuv(os.path.join(td, "in.yuv"), [(Y, U, V)])
            hm_util.write_pred(os.path.join(td, "pred"), 0, labels)
            cmd = [os.path.join(REFDIR, "TAppEncoder_saotrace"), "-c", hm_util.CFG, "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
                   "-f", "1", "-q", "32", "-b", "t.bin", "--InputBitDepth=8", "--InputChromaFormat=420", "--Level=6.2"]
            subprocess.check_call(cmd, cwd=td, env=dict(os.environ, HEVCDL_SAO_DUMP=os.path.join(td, "sao.bin")), stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
            raw = open(os.path.join(td, "sao.bin"), "rb").read()
        hdr = np.frombuffer(raw, np.int32, 8, 0)
        assert hdr[0] == 0x53414F30 and hdr[1] == W and hdr[2] == H and hdr[4] == 0
        n, off = int(hdr[3]), 32
        for kind in ("org", "src"):
            for name, cnt in (("Y", W * H), ("U", W * H // 4), ("V", W * H // 4)):
                out["%s%s_%d" % (kind, name, k)] = np.frombuffer(raw, np.int16, cnt, off).astype(np.uint8); off += 2 * cnt
        st = np.frombuffer(raw, np.int64, n * 3 * 5 * 64, off); off += 8 * n * 3 * 5 * 64
        assert off == len(raw) and np.abs(st).max() < 2 ** 31
        out["stats_%d" % k] = st.astype(np.int32).reshape(n, 3, 5, 2, 32)
        out["dims_%d" % k] = np.array([W, H])
        print("case", k, W, H, n, "CTUs, samples counted:", int(out["stats_%d" % k][:, :, :, 1].sum()))
    out["ncases"] = np.array(2)
    np.savez_compressed(os.path.join(GOLD, "sao_stats.npz"), **out)
    print("sao_stats.npz:", os.path.getsize(os.path.join(GOLD, "sao_stats.npz")), "bytes")


def saoapply_vectors():
    """tests/golden/sao_apply.npz: the deblocked picture, the resolved per-CTU SAO parameters and the picture after the reference's
    own TComSampleAdaptiveOffset::offsetCTU ran over every CTU (oracle/_ref/TAppEncoder_saoapplytrace, oracle/saoapply_dump.h):
    the 192x128 fixture frame at QP 37 and the 416x240 frame (partial CTUs on both edges) at QP 32 and 22 (the latter uses all five SAO types)."""
    g = np.load(os.path.join(GOLD, "rmd_trace_192x128_qp32.npz"))
    c = np.load(os.path.join(GOLD, "cnn_labels_416x240.npz"))
    out, k = {}, 0
    for (Y, U, V, labels, qp) in ((g["Y"], g["U"], g["V"], g["labels"], 37), (c["Y"], c["U"], c["V"], c["labels"], 32), (c["Y"], c["U"], c["V"], c["labels"], 22)):
        H, W = Y.shape
        with tempfile.TemporaryDirectory() as td:
            hm_util.write_yuv(os.path.join(td, "in.yuv"), [(Y, U, V)])
            hm_util.write_pred(os.path.join(td, "pred"), 0, labels)
            cmd = [os.path.join(REFDIR, "TAppEncoder_saoapplytrace"), "-c", hm_util.CFG, "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
                   "-f", "1", "-q", str(qp), "-b", "t.bin", "--InputBitDepth=8", "--InputChromaFormat=420", "--Level=6.2"]
            subprocess.check_call(cmd, cwd=td, env=dict(os.environ, HEVCDL_SAOAPPLY_DUMP=os.path.join(td, "sao.bin")), stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
            raw = open(os.path.join(td, "sao.bin"), "rb").read()
        hdr = np.frombuffer(raw, np.int32, 8, 0)
        assert hdr[0] == 0x53414F41 and hdr[1] == W and hdr[2] == H
        n, off = int(hdr[3]), 32
        for name, cnt in (("Y", W * H), ("U", W * H // 4), ("V", W * H // 4)):
            out["src%s_%d" % (name, k)] = np.frombuffer(raw, np.int16, cnt, off).astype(np.uint8); off += 2 * cnt
        out["type_%d" % k] = np.frombuffer(raw, np.int8, n * 3, off).reshape(n, 3); off += n * 3
        out["offset_%d" % k] = np.frombuffer(raw, np.int8, n * 3 * 32, off).reshape(n, 3, 32); off += n * 3 * 32
        for name, cnt in (("Y", W * H), ("U", W * H // 4), ("V", W * H // 4)):
            out["res%s_%d" % (name, k)] = np.frombuffer(raw, np.int16, cnt, off).astype(np.uint8); off += 2 * cnt
        assert off == len(raw)
        out["dims_%d" % k] = np.array([W, H, qp])
        ch = sum(int((out["src%s_%d" % (nm, k)] != out["res%s_%d" % (nm, k)]).sum()) for nm in "YUV")
        print("case", k, W, H, "qp", qp, "types used:", sorted(set(out["type_%d" % k].ravel().tolist())), "samples changed:", ch)
        k += 1
    out["ncases"] = np.array(k)
    np.savez_compressed(os.path.join(GOLD, "sao_apply.npz"), **out)
    print("sao_apply.npz:", os.path.getsize(os.path.join(GOLD, "sao_apply.npz")), "bytes")


def pred_vectors():
    """tests/golden/pred_trace_192x128_qp32.npz: reference samples and output of the reference's own TComPrediction::predIntraAng
    (oracle/_ref/TAppEncoder_predtrace, oracle/pred_dump.h) during an encode of the 192x128 fixture frame: up to 6 calls per
    (luma / chroma, block size, mode) -- first pass and RD pass, filtered and unfiltered references, Cb and Cr."""
    g = np.load(os.path.join(GOLD, "rmd_trace_192x128_qp32.npz"))
    Y, U, V, labels = g["Y"], g["U"], g["V"], g["labels"]
    H, W = Y.shape
    with tempfile.TemporaryDirectory() as td:
        hm_util.write_yuv(os.path.join(td, "in.yuv"), [(Y, U, V)])
        hm_util.write_pred(os.path.join(td, "pred"), 0, labels)
        cmd = [os.path.join(REFDIR, "TAppEncoder_predtrace"), "-c", hm_util.CFG, "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
               "-f", "1", "-q", "32", "-b", "t.bin", "--InputBitDepth=8", "--InputChromaFormat=420", "--Level=6.2"]
        subprocess.check_call(cmd, cwd=td, env=dict(os.environ, HEVCDL_PRED_DUMP=os.path.join(td, "pred.bin")), stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        raw = open(os.path.join(td, "pred.bin"), "rb").read()
    hdrs, lines, preds, off = [], [], [], 0
    while off < len(raw):
        h = np.frombuffer(raw, np.int32, 8, off); off += 32
        assert h[0] == 0x50524430
        n = int(h[3])
        lines.append(np.frombuffer(raw, np.int16, 4 * n + 1, off)); off += 2 * (4 * n + 1)
        preds.append(np.frombuffer(raw, np.int16, n * n, off)); off += 2 * n * n
        hdrs.append(h[1:6])
    hdrs = np.array(hdrs, np.int32)
    out = {"hdr": hdrs, "line": np.concatenate(lines), "pred": np.concatenate(preds),
           "line_off": np.concatenate([[0], np.cumsum([len(x) for x in lines])]), "pred_off": np.concatenate([[0], np.cumsum([len(x) for x in preds])])}
    np.savez_compressed(os.path.join(GOLD, "pred_trace_192x128_qp32.npz"), **out)
    print("pred_trace: %d calls (luma %d, chroma %d), sizes %s, %d bytes" % (len(hdrs), (hdrs[:, 0] == 0).sum(), (hdrs[:, 0] != 0).sum(),
          sorted(set(hdrs[:, 2].tolist())), os.path.getsize(os.path.join(GOLD, "pred_trace_192x128_qp32.npz"))))


if __name__ == "__main__":
    transform_vectors()
    trace_vectors()
    rdoq_vectors()
    dbf_vectors()
    sao_vectors()
    pred_vectors()
    saoapply_vectors()
