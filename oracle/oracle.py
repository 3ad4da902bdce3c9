"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/liboracle.so (cnn_oracle.c + rmd_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build():
    """Compile the C restatement (gcc, seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, s) for s in ("cnn_oracle.c", "rmd_oracle.c", "tq_oracle.c", "rdoq_oracle.c", "dbf_oracle.c", "sao_oracle.c")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", so] + srcs + ["-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_hdlw_nfloats.restype = C.c_int
        L.oracle_stage_ctu_rgb.argtypes = [u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.oracle_convnet2_forward.argtypes = [f32p, u8p, u8p, f32p]
        L.oracle_ctu_labels.argtypes = [f32p, u8p, C.c_void_p]
        L.oracle_frame_labels.argtypes = [f32p, u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p,
                                          C.c_void_p, C.c_void_p]
        L.oracle_build_ref_line.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i16p]
        L.oracle_filter_ref_line.argtypes = [i16p, C.c_int, i16p]
        L.oracle_mode_uses_filter.argtypes = [C.c_int, C.c_int]
        L.oracle_mode_uses_filter.restype = C.c_int
        L.oracle_predict.argtypes = [i16p, C.c_int, C.c_int, i16p]
        L.oracle_predict_ex.argtypes = [i16p, C.c_int, C.c_int, C.c_int, i16p]
        L.oracle_satd.argtypes = [u8p, C.c_int, i16p, C.c_int]
        L.oracle_satd.restype = C.c_uint32
        L.oracle_pu_satd35.argtypes = [C.c_void_p, C.c_int, i16p, C.c_int, u32p]
        L.oracle_cand_list.argtypes = [u32p, u32p, C.c_double, C.c_int, i32p, C.c_int, u8p, C.c_void_p]
        L.oracle_cand_list.restype = C.c_int
        L.oracle_mpm.argtypes = [C.c_int, C.c_int, i32p]
        L.oracle_mpm.restype = C.c_int
        L.oracle_enum_ctu_pus.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
        L.oracle_enum_ctu_pus.restype = C.c_int
        L.oracle_frame_rmd.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int, i32p, u32p]
        L.oracle_frame_rmd.restype = C.c_int
        L.oracle_tq_matrix.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_tq_matrix.restype = C.c_int
        L.oracle_tq_forward.argtypes = [i16p, C.c_int, C.c_int, i32p]
        L.oracle_tq_inverse.argtypes = [i32p, C.c_int, C.c_int, i16p]
        L.oracle_tq_tu.argtypes = [i16p, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, i16p]
        L.oracle_tq_tu.restype = C.c_uint32
        L.oracle_rdoq.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
        L.oracle_rdoq.restype = C.c_uint32
        L.oracle_deblock_frame.argtypes = [i16p, C.c_int, i16p, i16p, C.c_int, C.c_int, C.c_int, u8p, np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS"),
                                           C.c_int, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def set_threads(n):
    """Threads of the OpenMP loops (0 = all cores)."""
    lib().oracle_set_threads.argtypes = [C.c_int]
    lib().oracle_set_threads(int(n))


def load_weights(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"HDLW0001", "bad weight blob"
    w = np.frombuffer(raw, dtype="<f4", offset=8).astype(np.float32).copy()
    assert w.size == lib().oracle_hdlw_nfloats()
    return w


def stage_ctu_rgb(Y, U, V, ctu_x, ctu_y):
    H, W = Y.shape
    out = np.zeros((3, 64, 64), np.uint8)
    lib().oracle_stage_ctu_rgb(np.ascontiguousarray(Y), np.ascontiguousarray(U), np.ascontiguousarray(V),
                               W, H, ctu_x, ctu_y, out)
    return out


def convnet2_forward(w, x32, x64):
    out = np.zeros(16, np.float32)
    lib().oracle_convnet2_forward(w, np.ascontiguousarray(x32), np.ascontiguousarray(x64), out)
    return out


def ctu_labels(logits4):
    lab = np.zeros(16, np.uint8)
    mar = np.zeros(16, np.float32)
    lib().oracle_ctu_labels(np.ascontiguousarray(logits4, np.float32), lab, mar.ctypes.data)
    return lab, mar


def frame_labels(w, Y, U, V, ctu_begin=0, ctu_end=None, want_logits=False):
    H, W = Y.shape
    nctu = ((W + 63) // 64) * ((H + 63) // 64)
    if ctu_end is None:
        ctu_end = nctu
    labels = np.zeros((nctu, 16), np.uint8)
    logits = np.zeros((nctu, 4, 16), np.float32)
    margins = np.zeros((nctu, 16), np.float32)
    lib().oracle_frame_labels(w, np.ascontiguousarray(Y), np.ascontiguousarray(U), np.ascontiguousarray(V),
                              W, H, ctu_begin, ctu_end, labels, logits.ctypes.data, margins.ctypes.data)
    if want_logits:
        return labels, logits, margins
    return labels


def build_ref_line(pic, x0, y0, n):
    H, W = pic.shape
    line = np.zeros(4 * n + 1, np.int16)
    lib().oracle_build_ref_line(np.ascontiguousarray(pic), W, W, H, x0, y0, n, line)
    return line


def filter_ref_line(line, n):
    out = np.zeros_like(line)
    lib().oracle_filter_ref_line(np.ascontiguousarray(line), n, out)
    return out


def predict(line, n, mode, edge=True):
    """One intra-predicted block from a reference line (4n+1 samples: left column bottom-up, corner, top row).  edge: luma
    edge filters (False = chroma)."""
    out = np.zeros((n, n), np.int16)
    lib().oracle_predict_ex(np.ascontiguousarray(line, np.int16), n, mode, 1 if edge else 0, out)
    return out


def pu_satd35(org_pic, x0, y0, n, line):
    """org_pic: full uint8 picture; PU at (x0,y0)."""
    H, W = org_pic.shape
    org_pic = np.ascontiguousarray(org_pic)
    out = np.zeros(35, np.uint32)
    lib().oracle_pu_satd35(C.c_void_p(int(org_pic.ctypes.data) + int(y0) * int(W) + int(x0)), W, np.ascontiguousarray(line), n, out)
    return out


def block_satd35(org_block, line):
    n = org_block.shape[0]
    blk = np.ascontiguousarray(org_block, np.uint8)
    out = np.zeros(35, np.uint32)
    lib().oracle_pu_satd35(C.c_void_p(int(blk.ctypes.data)), n, np.ascontiguousarray(line), n, out)
    return out


def cand_list(satd, bits, sqrt_lambda, n, mpm, n_mpm_add):
    modes = np.zeros(10, np.uint8)
    k = lib().oracle_cand_list(np.ascontiguousarray(satd, np.uint32), np.ascontiguousarray(bits, np.uint32),
                               float(sqrt_lambda), n, np.ascontiguousarray(mpm, np.int32), n_mpm_add, modes, None)
    return modes[:k].copy()


def mpm(left, above):
    m = np.zeros(3, np.int32)
    k = lib().oracle_mpm(left, above, m)
    return m, k


def enum_ctu_pus(label16, ctu_x, ctu_y, W, H):
    pu = np.zeros((340, 4), np.int32)
    k = lib().oracle_enum_ctu_pus(np.ascontiguousarray(label16, np.uint8), ctu_x, ctu_y, W, H, pu)
    return pu[:k].copy()


def frame_rmd(pic, labels, ctu_begin=0, ctu_end=None):
    H, W = pic.shape
    nctu = ((W + 63) // 64) * ((H + 63) // 64)
    if ctu_end is None:
        ctu_end = nctu
    cap = (ctu_end - ctu_begin) * 320
    pu = np.zeros((cap, 4), np.int32)
    satd = np.zeros((cap, 35), np.uint32)
    k = lib().oracle_frame_rmd(np.ascontiguousarray(pic), W, H, np.ascontiguousarray(labels, np.uint8),
                               ctu_begin, ctu_end, pu, satd)
    return pu[:k].copy(), satd[:k].copy()


def label_parity(labels, ref_labels, margins, eps, logits=None, ref_logits=None):
    """Parity statistics of device labels against the oracle's (checker side of tests/ and of bench.py's `parity` key).
    margins [nctu,16]: the oracle's argmax margins (top logit minus runner-up, per 16x16 block).  The reference's fix-up
    rules (use_model.py:102-119) couple the 16 labels of a CTU, so a CTU must match whenever ALL its margins exceed eps.
    With both logit arrays the argmax decisions themselves are compared: `argmax_flips` and the largest oracle margin
    among the flipped ones (`max_flipped_margin`) -- the evidence eps is set from."""
    labels, ref_labels = np.asarray(labels), np.asarray(ref_labels)
    n = len(labels)
    mar = np.asarray(margins).reshape(n, -1)
    diff_ctu = (labels != ref_labels).any(axis=1)
    safe = mar.min(axis=1) > eps
    out = {"ctus": int(n), "labels": int(labels.size), "labels_differing": int((labels != ref_labels).sum()),
           "ctus_differing": int(diff_ctu.sum()), "eps": float(eps), "ctus_all_margins_above_eps": int(safe.sum()),
           "ctus_differing_above_eps": int((diff_ctu & safe).sum()),
           "max_min_margin_of_differing_ctu": float(mar.min(axis=1)[diff_ctu].max()) if diff_ctu.any() else 0.0}
    if logits is not None and ref_logits is not None:
        a = np.asarray(logits).reshape(n, 16, 4).argmax(axis=2)          # [ctu][quadrant*4 + group] -> digit (first maximum)
        b = np.asarray(ref_logits).reshape(n, 16, 4).argmax(axis=2)
        flip = a != b
        # margins are stored per label position (4x4 raster); group g of quadrant q sits at SCATTER[q][g]
        scatter = np.array([0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15])
        mg = mar[:, scatter]
        out["argmax_flips"] = int(flip.sum())
        out["max_flipped_margin"] = float(mg[flip].max()) if flip.any() else 0.0
        out["max_abs_dlogit"] = float(np.abs(np.asarray(logits) - np.asarray(ref_logits)).max())
    return out


# ---- transform / quantisation core (oracle/tq_oracle.c) -------------------------------------------------------------
TQ_FLAG_DST, TQ_FLAG_TSKIP, TQ_FLAG_INTER = 1, 2, 4


def tq_matrix(n):
    """The n-point HEVC core transform matrix [k][x]."""
    return np.array([[lib().oracle_tq_matrix(n, k, x) for x in range(n)] for k in range(n)], np.int32)


def tq_forward(resi, dst=False):
    n = resi.shape[0]
    out = np.zeros((n, n), np.int32)
    lib().oracle_tq_forward(np.ascontiguousarray(resi, np.int16), n, int(dst), out)
    return out


def tq_inverse(coeff, dst=False):
    n = coeff.shape[0]
    out = np.zeros((n, n), np.int16)
    lib().oracle_tq_inverse(np.ascontiguousarray(coeff, np.int32), n, int(dst), out)
    return out


def tq_tu(resi, qp, flags=0):
    """One TU through transform, flat quantiser, dequantiser and inverse transform: (coeff, level, deq, rec, abs_sum)."""
    n = resi.shape[0]
    c, q, d = (np.zeros((n, n), np.int32) for _ in range(3))
    r = np.zeros((n, n), np.int16)
    s = lib().oracle_tq_tu(np.ascontiguousarray(resi, np.int16), int(n).bit_length() - 1, int(qp), int(flags), c, q, d, r)
    return c, q, d, r, int(s)


def rdoq(coeff, ch, scan_type, qp, tskip, lam, est, ctx_cbf, is_intra=1, tr_idx_zero=0, sdh=1):
    """Rate-distortion optimised quantisation of one TU (oracle/rdoq_oracle.c): coeff (n,n) int32 transform output, ch 0 luma /
    1 chroma, scan_type 0 diag / 1 hor / 2 ver, lam = the encoder's lambda for the component, est = the 224 int32 of the
    reference's estBitsSbacStruct.  Returns (levels (n,n) int32, abs_sum)."""
    n = coeff.shape[0]
    out = np.zeros((n, n), np.int32)
    s = lib().oracle_rdoq(np.ascontiguousarray(coeff, np.int32), int(n).bit_length() - 1, int(ch), int(scan_type), int(qp), int(tskip),
                          float(lam), np.ascontiguousarray(est, np.int32), int(ctx_cbf), int(is_intra), int(tr_idx_zero), int(sdh), out)
    return out, int(s)


def deblock_frame(Y, U, V, tu_log2, qp, beta_off_div2=0, tc_off_div2=0, cb_qp_off=0, cr_qp_off=0):
    """Deblocking filter of an all-intra picture (oracle/dbf_oracle.c).  Y, U, V: 8-bit planes (any integer dtype); tu_log2, qp: one
    entry per 4x4 luma unit ((H/4, W/4) or flat).  Returns the filtered planes as uint8."""
    H, W = Y.shape
    y, u, v = (np.ascontiguousarray(p, np.int16).copy() for p in (Y, U, V))
    lib().oracle_deblock_frame(y, W, u, v, W // 2, W, H, np.ascontiguousarray(tu_log2, np.uint8).ravel(), np.ascontiguousarray(qp, np.int8).ravel(),
                               int(beta_off_div2), int(tc_off_div2), int(cb_qp_off), int(cr_qp_off))
    return y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)


def sao_stats(org, src):
    """SAO statistics of a picture (oracle/sao_oracle.c).  org, src: (Y, U, V) 8-bit planes (original, deblocked).  Returns
    int64 [nctu, 3 components, 5 types (EO 0 / 90 / 135 / 45, BO), 2 (diff, count), 32 classes]."""
    H, W = org[0].shape
    n = ((W + 63) // 64) * ((H + 63) // 64)
    o = [np.ascontiguousarray(p, np.int16) for p in org]
    s = [np.ascontiguousarray(p, np.int16) for p in src]
    P = C.POINTER(C.c_int16) * 3
    out = np.zeros((n, 3, 5, 2, 32), np.int64)
    lib().oracle_sao_stats(P(*[p.ctypes.data_as(C.POINTER(C.c_int16)) for p in o]), P(*[p.ctypes.data_as(C.POINTER(C.c_int16)) for p in s]),
                           W, H, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def sao_apply(src, types, offsets):
    """SAO application over a picture (oracle/sao_oracle.c: oracle_sao_apply).  src: (Y, U, V) deblocked 8-bit planes; types:
    int8 [nctu, 3] (-1 off, 0..3 edge offset 0 / 90 / 135 / 45, 4 band offset); offsets: int8 [nctu, 3, 32].  Returns the
    (Y, U, V) planes with the offsets applied, as uint8."""
    H, W = src[0].shape
    s = [np.ascontiguousarray(p, np.int16) for p in src]
    r = [np.zeros_like(p) for p in s]
    t = np.ascontiguousarray(types, np.int8)
    o = np.ascontiguousarray(offsets, np.int8)
    n = ((W + 63) // 64) * ((H + 63) // 64)
    assert t.shape == (n, 3) and o.shape == (n, 3, 32)
    P = C.POINTER(C.c_int16) * 3
    lib().oracle_sao_apply(P(*[p.ctypes.data_as(C.POINTER(C.c_int16)) for p in s]), P(*[p.ctypes.data_as(C.POINTER(C.c_int16)) for p in r]),
                           W, H, t.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    return [p.astype(np.uint8) for p in r]
