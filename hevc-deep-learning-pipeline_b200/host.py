"""Host-side mirror of the reference sidecar over the C-ABI (include/hevcdl.h).

The reference's depth-prediction interface is two scripts coupled to the encoder by files:
gen_frames.py (frame dump) and use_model.py (per-CTU labels written to ./pred/<frame>/ctu<i>.txt,
use_model.py:73-127).  `DepthPredictor` exposes the same operations -- predict the 16 labels of
every CTU of a frame, optionally write the reference's text files -- backed by libhevcdl.so.
There is no CPU path: constructing a DepthPredictor without the CUDA library or a B200 raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEVCDL_LIB") or os.path.join(_HERE, "csrc", "libhevcdl.so")   # HEVCDL_LIB: tuning builds only
DEFAULT_WEIGHTS = os.path.join(os.path.dirname(_HERE), "weights", "hevc_encoder_model.hdlw")

PREC_FP32, PREC_BF16_TC = 0, 1
OUT_LOGITS, OUT_SATD = 1, 2            # hevcdl_output_flags
ABI_VERSION = 3

# include/hevcdl.h: the drop-in boundary
EXPORTS = [
    "hevcdl_create", "hevcdl_destroy", "hevcdl_last_error", "hevcdl_status_str",
    "hevcdl_submit_frame_u8", "hevcdl_submit_frame_pel16", "hevcdl_wait_frame", "hevcdl_ctu_labels",
    "hevcdl_frame_labels", "hevcdl_frame_pu_count", "hevcdl_frame_pus", "hevcdl_ctu_pu_range",
    "hevcdl_frame_view_get", "hevcdl_release_frame", "hevcdl_rmd_exact", "hevcdl_tu_code", "hevcdl_tu_code_rdoq", "hevcdl_deblock_frame", "hevcdl_sao_stats", "hevcdl_sao_apply", "hevcdl_inloop_frame", "hevcdl_intra_pred", "hevcdl_get_stats", "hevcdl_numa_bind_thread", "hevcdl_host_alloc", "hevcdl_host_free", "hevcdl_host_register", "hevcdl_host_unregister",
]
# include/hevcdl_internal.h: measurement and test hooks
EXPORTS_INTERNAL = ["hevcdl_bench_resident", "hevcdl_bench_e2e", "hevcdl_debug_copy", "hevcdl_debug_rerun_rmd", "hevcdl_stream",
                    "hevcdl_last_aux_ms"]


class Cfg(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("slots", C.c_int32), ("precision", C.c_int32), ("rmd", C.c_int32), ("boundary_fix", C.c_int32),
                ("batch", C.c_int32), ("outputs", C.c_int32), ("pinned_input", C.c_int32), ("numa_bind", C.c_int32),
                ("weights_path", C.c_char_p)]


class Stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("ctus", C.c_uint64), ("pus", C.c_uint64), ("ms_cnn", C.c_double),
                ("ms_rmd", C.c_double), ("kernel_launches", C.c_uint64)]


class FrameView(C.Structure):
    _fields_ = [("labels", C.c_void_p), ("logits", C.c_void_p), ("ctu_off", C.c_void_p), ("pus", C.c_void_p),
                ("satd", C.c_void_p), ("cand", C.c_void_p), ("nctu", C.c_int32), ("npu", C.c_int32)]


TU_DTYPE = np.dtype([("log2_size", "u1"), ("qp", "u1"), ("flags", "u1"), ("reserved", "u1"), ("offset", "<u4")])
TU_DST, TU_TSKIP, TU_INTER, TU_RDOQ, TU_COEFF_IN = 1, 2, 4, 8, 16
TU_RDOQ_DTYPE = np.dtype([("lambda", "<f8"), ("est_index", "<u4"), ("channel", "u1"), ("scan_type", "u1"), ("ctx_cbf", "u1"), ("flags", "u1")])
EST_INTS = 224
PRED_REQ_DTYPE = np.dtype([("log2_size", "u1"), ("mode", "u1"), ("flags", "u1"), ("reserved", "u1"), ("line_offset", "<u4"), ("pred_offset", "<u4")])
PRED_EDGE = 1
SAO_PARAM_DTYPE = np.dtype([("type", "i1"), ("reserved", "i1", 3), ("offset", "i1", 32)])
PU_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("size", "u1"), ("part", "u1"), ("ctu", "<u2")])


class HevcdlError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libhevcdl.so (built in-tree by __graft_entry__.build()); fail loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HevcdlError("libhevcdl.so not built (%s): run __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.c_int
    L.hevcdl_create.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
    L.hevcdl_destroy.argtypes = [vp]
    L.hevcdl_destroy.restype = None
    L.hevcdl_last_error.argtypes = [vp]
    L.hevcdl_last_error.restype = C.c_char_p
    L.hevcdl_status_str.argtypes = [ip]
    L.hevcdl_status_str.restype = C.c_char_p
    L.hevcdl_submit_frame_u8.argtypes = [vp, ip, vp, ip, vp, vp, ip]
    L.hevcdl_submit_frame_pel16.argtypes = [vp, ip, vp, ip, vp, vp, ip]
    L.hevcdl_wait_frame.argtypes = [vp, ip]
    L.hevcdl_ctu_labels.argtypes = [vp, ip, ip, vp]
    L.hevcdl_frame_labels.argtypes = [vp, ip, vp, vp]
    L.hevcdl_frame_pu_count.argtypes = [vp, ip, C.POINTER(ip)]
    L.hevcdl_frame_pus.argtypes = [vp, ip, vp, vp, vp]
    L.hevcdl_ctu_pu_range.argtypes = [vp, ip, ip, C.POINTER(ip), C.POINTER(ip)]
    L.hevcdl_frame_view_get.argtypes = [vp, ip, ip, C.POINTER(FrameView)]
    L.hevcdl_release_frame.argtypes = [vp, ip]
    L.hevcdl_rmd_exact.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, C.c_double, vp, vp, vp]
    L.hevcdl_bench_resident.argtypes = [vp, vp, ip, ip, C.POINTER(C.c_float), C.POINTER(ip)]
    L.hevcdl_bench_e2e.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, ip, ip, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64)]
    L.hevcdl_debug_rerun_rmd.argtypes = [vp, ip, vp]
    L.hevcdl_debug_copy.argtypes = [vp, ip, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hevcdl_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.hevcdl_last_aux_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.hevcdl_stream.argtypes = [vp]
    L.hevcdl_stream.restype = vp
    L.hevcdl_numa_bind_thread.argtypes = [ip]
    L.hevcdl_host_alloc.argtypes = [C.c_size_t, ip]
    L.hevcdl_host_alloc.restype = vp
    L.hevcdl_host_free.argtypes = [vp]
    L.hevcdl_host_free.restype = None
    L.hevcdl_sao_stats.argtypes = [vp, vp, vp, vp, ip, ip, vp, vp, vp, ip, ip, ip, ip, vp]
    L.hevcdl_host_register.argtypes = [vp, C.c_size_t]
    L.hevcdl_host_unregister.argtypes = [vp]
    L.hevcdl_inloop_frame.argtypes = [vp, vp, ip, vp, vp, ip, ip, ip, vp, vp, ip, ip, ip, ip, vp, vp, vp, ip, ip, vp]
    L.hevcdl_sao_apply.argtypes = [vp, vp, vp, vp, ip, ip, vp, vp, vp, ip, ip, ip, ip, vp]
    L.hevcdl_intra_pred.argtypes = [vp, ip, vp, vp, C.c_size_t, vp, C.c_size_t]
    L.hevcdl_deblock_frame.argtypes = [vp, vp, ip, vp, vp, ip, ip, ip, vp, vp, ip, ip, ip, ip]
    L.hevcdl_tu_code.argtypes = [vp, ip, vp, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
    L.hevcdl_tu_code_rdoq.argtypes = [vp, ip, vp, vp, vp, ip, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class DepthPredictor:
    """One context per GPU (device = LOCAL_RANK).  Frames are independent (all-intra,
    encoder_intra_main.cfg:20-22), so ranks shard frames f -> rank f % world with no exchange."""

    def __init__(self, width, height, device=0, slots=2, precision=PREC_FP32, rmd=True, boundary_fix=False,
                 weights=DEFAULT_WEIGHTS, batch=1, outputs=OUT_LOGITS | OUT_SATD, pinned_input=False, numa_bind=False):
        """outputs: hevcdl_output_flags -- this mirror defaults to everything (logits for margin reports, SATD tables for
        the parity tests); the C default, and what bench.py's end-to-end leg uses, is 0: labels + PU list + candidate modes.
        pinned_input: planes handed to submit() are page-locked and stay untouched until the frame is waited for."""
        self.lib = load_library()
        self.outputs = int(outputs)
        self.width, self.height = int(width), int(height)
        self.ctu_w, self.ctu_h = (self.width + 63) // 64, (self.height + 63) // 64
        self.nctu = self.ctu_w * self.ctu_h
        self.rmd = bool(rmd)
        cfg = Cfg(ABI_VERSION, device, self.width, self.height, slots, precision, int(rmd), int(boundary_fix), int(batch),
                  int(outputs), int(pinned_input), int(numa_bind), os.fsencode(weights))
        h = C.c_void_p()
        rc = self.lib.hevcdl_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise HevcdlError("hevcdl_create: %s (%s)" % (self.lib.hevcdl_status_str(rc).decode(),
                                                          self.lib.hevcdl_last_error(None).decode()))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.hevcdl_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise HevcdlError("%s: %s (%s)" % (what, self.lib.hevcdl_status_str(rc).decode(),
                                               self.lib.hevcdl_last_error(self.h).decode()))

    # -- frame pipeline ---------------------------------------------------------------------
    def submit(self, frame, Y, U, V):
        """Queue one 8-bit 4:2:0 picture (uint8 planes, or int16 'Pel' planes as HM holds them)."""
        assert Y.shape == (self.height, self.width) and U.shape == (self.height // 2, self.width // 2)
        if Y.dtype == np.int16:
            fn, esz = self.lib.hevcdl_submit_frame_pel16, 2
        else:
            assert Y.dtype == np.uint8
            fn, esz = self.lib.hevcdl_submit_frame_u8, 1
        assert Y.strides[1] == esz and U.strides[1] == esz and V.strides == U.strides
        self._ck(fn(self.h, frame, _ptr(Y), Y.strides[0] // esz, _ptr(U), _ptr(V), U.strides[0] // esz), "submit_frame")

    def wait(self, frame):
        self._ck(self.lib.hevcdl_wait_frame(self.h, frame), "wait_frame")

    def labels(self, frame, want_logits=False):
        lab = np.empty((self.nctu, 16), np.uint8)
        lg = np.empty((self.nctu, 4, 16), np.float32) if want_logits else None
        self._ck(self.lib.hevcdl_frame_labels(self.h, frame, _ptr(lab), _ptr(lg)), "frame_labels")
        return (lab, lg) if want_logits else lab

    def ctu_labels(self, frame, addr):
        out = np.empty(16, np.uint8)
        self._ck(self.lib.hevcdl_ctu_labels(self.h, frame, addr, _ptr(out)), "ctu_labels")
        return out

    def pus(self, frame):
        n = C.c_int()
        self._ck(self.lib.hevcdl_frame_pu_count(self.h, frame, C.byref(n)), "frame_pu_count")
        pus = np.empty(n.value, PU_DTYPE)
        satd = np.empty((n.value, 35), np.uint32) if self.outputs & OUT_SATD else None
        cand = np.empty((n.value, 8), np.uint8)
        self._ck(self.lib.hevcdl_frame_pus(self.h, frame, _ptr(pus), _ptr(satd), _ptr(cand)), "frame_pus")
        return pus, satd, cand

    def view(self, frame, want_pus=True):
        """Zero-copy results of one frame: numpy arrays over the context's pinned host buffers, valid until
        release(frame).  Returns dict(labels [nctu,16], logits [nctu,4,16], ctu_off, pus, satd [npu,35], cand [npu,8])."""
        v = FrameView()
        self._ck(self.lib.hevcdl_frame_view_get(self.h, frame, int(want_pus and self.rmd), C.byref(v)), "frame_view_get")

        cache = self.__dict__.setdefault("_views", {})     # slot buffers are stable: wrap each (pointer, length) once

        def arr(ptr, ctype, n, shape, dtype=None):
            if not ptr or n == 0:
                return np.empty(shape, dtype or np.dtype(ctype))
            key = (ptr, n, ctype, shape, dtype)     # pus and cand may alias by address across frames with different PU counts
            a = cache.get(key)
            if a is None:
                if len(cache) > 4096:
                    cache.clear()
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))
                a = (a.view(dtype) if dtype is not None else a).reshape(shape)
                cache[key] = a
            return a
        n, m = v.nctu, v.npu
        return {"labels": arr(v.labels, C.c_uint8, n * 16, (n, 16)),
                "logits": arr(v.logits, C.c_float, n * 64 if v.logits else 0, (n if v.logits else 0, 4, 16)),
                "ctu_off": arr(v.ctu_off, C.c_int32, n + 1 if v.ctu_off else 0, (n + 1 if v.ctu_off else 0,)),
                "pus": arr(v.pus, C.c_uint8, m * 8, (m,), PU_DTYPE),
                "satd": arr(v.satd, C.c_uint32, m * 35 if v.satd else 0, (m if v.satd else 0, 35)),
                "cand": arr(v.cand, C.c_uint8, m * 8, (m, 8))}

    def ctu_pu_range(self, frame, addr):
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.hevcdl_ctu_pu_range(self.h, frame, addr, C.byref(a), C.byref(b)), "ctu_pu_range")
        return a.value, b.value

    def release(self, frame):
        self._ck(self.lib.hevcdl_release_frame(self.h, frame), "release_frame")

    def predict_frame(self, Y, U, V, frame=0, want_logits=False):
        """use_model.py's per-frame loop (:74-127) in one call: labels [nctu,16] in CTU raster order."""
        self.submit(frame, Y, U, V)
        out = self.labels(frame, want_logits)
        self.release(frame)
        return out

    def write_pred_files(self, labels, pred_dir, frame):
        """Emit the reference's handshake files (use_model.py:121-125: '<d> ' x16, no newline) so an
        UNMODIFIED TAppEncoder can consume labels computed here."""
        d = os.path.join(pred_dir, str(frame))
        os.makedirs(d, exist_ok=True)
        for i, l in enumerate(labels):
            with open(os.path.join(d, "ctu.txt"), "w") as f:
                f.write("".join("%d " % v for v in l))
            os.rename(os.path.join(d, "ctu.txt"), os.path.join(d, "ctu%d.txt" % i))

    # -- exact RMD ---------------------------------------------------------------------------
    def rmd_exact(self, sizes, org_blocks, lines, bits=None, mpm=None, mpm_add=None, sqrt_lambda=0.0):
        """sizes [n] u8; org_blocks: list of (s,s) u8; lines: list of (4s+1,) i16; bits [n,35] u32;
        mpm [n,3] i8; mpm_add [n] u8.  Returns satd [n,35], cand [n,10], ncand [n]."""
        n = len(sizes)
        sizes = np.ascontiguousarray(sizes, np.uint8)
        org = np.ascontiguousarray(np.concatenate([np.asarray(b, np.uint8).ravel() for b in org_blocks]))
        ln = np.ascontiguousarray(np.concatenate([np.asarray(l, np.int16).ravel() for l in lines]))
        bits = None if bits is None else np.ascontiguousarray(bits, np.uint32)
        mpm = None if mpm is None else np.ascontiguousarray(mpm, np.int8)
        mpm_add = None if mpm_add is None else np.ascontiguousarray(mpm_add, np.uint8)
        satd = np.empty((n, 35), np.uint32)
        cand = np.empty((n, 10), np.uint8)
        ncand = np.empty(n, np.uint8)
        self._ck(self.lib.hevcdl_rmd_exact(self.h, n, _ptr(sizes), _ptr(org), _ptr(ln), _ptr(bits), _ptr(mpm),
                                           _ptr(mpm_add), float(sqrt_lambda), _ptr(satd), _ptr(cand), _ptr(ncand)),
                 "rmd_exact")
        return satd, cand, ncand

    # -- transform-unit coding core --------------------------------------------------------------
    def tu_code(self, blocks, qps, flags=None, want_coeff=True, want_deq=True, rdoq=None, est=None):
        """blocks: list of (N,N) int16 residual blocks (N = 4, 8, 16, 32); qps, flags: per block.  Forward transform, flat
        quantiser, dequantiser, inverse transform (hevcdl_tu_code).  rdoq (TU_RDOQ_DTYPE array, one per block) + est
        ([n_est, 224] int32 bit-estimate tables): blocks flagged TU_RDOQ go through the rate-distortion optimised quantiser
        (hevcdl_tu_code_rdoq).  Returns dict of per-block lists coeff / level / deq / rec plus arrays abs_sum, ssd."""
        n = len(blocks)
        tus = np.zeros(n, TU_DTYPE)
        sizes = np.array([b.shape[0] for b in blocks], np.int64)
        off = np.concatenate([[0], np.cumsum(sizes * sizes)])
        tus["log2_size"] = [int(s).bit_length() - 1 for s in sizes]
        tus["qp"] = qps
        tus["flags"] = 0 if flags is None else flags
        tus["offset"] = off[:-1]
        nelem = int(off[-1])
        resi = np.concatenate([np.asarray(b, np.int16).ravel() for b in blocks]) if n else np.zeros(0, np.int16)
        coeff = np.zeros(nelem, np.int32) if want_coeff else None
        level = np.zeros(nelem, np.int16)
        deq = np.zeros(nelem, np.int32) if want_deq else None
        rec = np.zeros(nelem, np.int16)
        asum = np.zeros(n, np.uint32)
        ssd = np.zeros(n, np.uint64)
        if rdoq is None:
            self._ck(self.lib.hevcdl_tu_code(self.h, n, _ptr(tus), _ptr(resi), nelem, _ptr(coeff), _ptr(level), _ptr(deq), _ptr(rec),
                                             _ptr(asum), _ptr(ssd)), "tu_code")
        else:
            rq = np.ascontiguousarray(rdoq, TU_RDOQ_DTYPE)
            et = np.ascontiguousarray(est, np.int32).reshape(-1, EST_INTS)
            assert len(rq) == n
            self._ck(self.lib.hevcdl_tu_code_rdoq(self.h, n, _ptr(tus), _ptr(rq), _ptr(et), len(et), _ptr(resi), nelem, _ptr(coeff),
                                                  _ptr(level), _ptr(deq), _ptr(rec), _ptr(asum), _ptr(ssd)), "tu_code_rdoq")

        def split(a):
            return None if a is None else [a[off[i]:off[i + 1]].reshape(sizes[i], sizes[i]) for i in range(n)]
        return {"coeff": split(coeff), "level": split(level), "deq": split(deq), "rec": split(rec), "abs_sum": asum, "ssd": ssd}

    # -- intra prediction of single blocks ------------------------------------------------------------
    def intra_pred(self, lines, modes, edge):
        """lines: list of reference lines (4N+1 int16 each: left column bottom-up, corner, top row); modes, edge: per block the
        intra mode 0..34 and whether the luma edge filters apply (hevcdl_intra_pred).  Returns the list of (N, N) int16 blocks."""
        n = len(lines)
        sizes = np.array([(len(l) - 1) // 4 for l in lines], np.int64)
        loff = np.concatenate([[0], np.cumsum([len(l) for l in lines])])
        poff = np.concatenate([[0], np.cumsum(sizes * sizes)])
        rq = np.zeros(n, PRED_REQ_DTYPE)
        rq["log2_size"] = [int(s).bit_length() - 1 for s in sizes]
        rq["mode"] = modes
        rq["flags"] = [PRED_EDGE if e else 0 for e in edge]
        rq["line_offset"] = loff[:-1]
        rq["pred_offset"] = poff[:-1]
        ln = np.concatenate([np.asarray(l, np.int16) for l in lines]) if n else np.zeros(0, np.int16)
        out = np.zeros(int(poff[-1]), np.int16)
        self._ck(self.lib.hevcdl_intra_pred(self.h, n, _ptr(rq), _ptr(ln), ln.size, _ptr(out), out.size), "intra_pred")
        return [out[poff[i]:poff[i + 1]].reshape(sizes[i], sizes[i]) for i in range(n)]

    # -- in-loop deblocking filter -----------------------------------------------------------------
    def deblock_frame(self, Y, U, V, tu_log2, qp, beta_off_div2=0, tc_off_div2=0, cb_qp_off=0, cr_qp_off=0):
        """Deblocking filter of an all-intra reconstructed picture (hevcdl_deblock_frame).  Y, U, V: 8-bit planes of any integer
        dtype; tu_log2, qp: one entry per 4x4 luma unit.  Returns the filtered planes as uint8."""
        H, W = Y.shape
        y, u, v = (np.ascontiguousarray(p, np.int16).copy() for p in (Y, U, V))
        tu = np.ascontiguousarray(tu_log2, np.uint8).ravel()
        q = np.ascontiguousarray(qp, np.int8).ravel()
        assert tu.size == q.size == (W // 4) * (H // 4)
        self._ck(self.lib.hevcdl_deblock_frame(self.h, _ptr(y), W, _ptr(u), _ptr(v), W // 2, W, H, _ptr(tu), _ptr(q), int(beta_off_div2),
                                               int(tc_off_div2), int(cb_qp_off), int(cr_qp_off)), "deblock_frame")
        return y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)

    def sao_stats(self, org, rec):
        """SAO statistics of a deblocked picture (hevcdl_sao_stats).  org, rec: (Y, U, V) 8-bit planes.  Returns int64
        [nctu, 3 components, 5 types (EO 0 / 90 / 135 / 45, BO), 2 (diff, count), 32 classes]."""
        H, W = org[0].shape
        n = ((W + 63) // 64) * ((H + 63) // 64)
        o = [np.ascontiguousarray(p, np.int16) for p in org]
        r = [np.ascontiguousarray(p, np.int16) for p in rec]
        out = np.zeros((n, 3, 5, 2, 32), np.int64)
        self._ck(self.lib.hevcdl_sao_stats(self.h, _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), W, W // 2, _ptr(r[0]), _ptr(r[1]), _ptr(r[2]), W, W // 2,
                                           W, H, _ptr(out)), "sao_stats")
        return out

    def inloop_frame(self, Y, U, V, tu_log2, qp, org, beta_off_div2=0, tc_off_div2=0, cb_qp_off=0, cr_qp_off=0):
        """Deblocking + SAO statistics in one round trip (hevcdl_inloop_frame); the deblocked picture stays resident for a following
        sao_apply(None, ...).  Returns ((Y, U, V) deblocked uint8, stats int64 [nctu, 3, 5, 2, 32])."""
        H, W = Y.shape
        y, u, v = (np.ascontiguousarray(p, np.int16).copy() for p in (Y, U, V))
        o = [np.ascontiguousarray(p, np.int16) for p in org]
        tu = np.ascontiguousarray(tu_log2, np.uint8).ravel()
        q = np.ascontiguousarray(qp, np.int8).ravel()
        n = ((W + 63) // 64) * ((H + 63) // 64)
        st = np.zeros((n, 3, 5, 2, 32), np.int64)
        self._ck(self.lib.hevcdl_inloop_frame(self.h, _ptr(y), W, _ptr(u), _ptr(v), W // 2, W, H, _ptr(tu), _ptr(q), int(beta_off_div2),
                                              int(tc_off_div2), int(cb_qp_off), int(cr_qp_off), _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), W, W // 2,
                                              _ptr(st)), "inloop_frame")
        return [p.astype(np.uint8) for p in (y, u, v)], st

    def sao_apply(self, src, types, offsets, shape=None):
        """SAO application over a deblocked picture (hevcdl_sao_apply).  src: (Y, U, V) 8-bit planes, or None = the picture the
        preceding inloop_frame left on the device (then shape = (H, W)); types: int8 [nctu, 3] (-1 off, 0..3 edge offset, 4 band
        offset); offsets: int8 [nctu, 3, 32].  Returns the (Y, U, V) planes as uint8."""
        H, W = src[0].shape if src is not None else shape
        n = ((W + 63) // 64) * ((H + 63) // 64)
        s = [np.ascontiguousarray(p, np.int16) for p in src] if src is not None else [None, None, None]
        r = [np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)]
        prm = np.zeros((n, 3), SAO_PARAM_DTYPE)
        prm["type"] = np.asarray(types, np.int8).reshape(n, 3)
        prm["offset"] = np.asarray(offsets, np.int8).reshape(n, 3, 32)
        self._ck(self.lib.hevcdl_sao_apply(self.h, _ptr(s[0]), _ptr(s[1]), _ptr(s[2]), W, W // 2, _ptr(r[0]), _ptr(r[1]), _ptr(r[2]), W, W // 2,
                                           W, H, _ptr(prm)), "sao_apply")
        return [p.astype(np.uint8) for p in r]

    # -- measurement -------------------------------------------------------------------------
    def last_aux_ms(self):
        """Device time of the kernels of the most recent tu_code / deblock_frame / sao_stats call (hevcdl_last_aux_ms)."""
        ms = C.c_float()
        self._ck(self.lib.hevcdl_last_aux_ms(self.h, C.byref(ms)), "last_aux_ms")
        return ms.value

    def bench_resident(self, frames, iters):
        fr = np.ascontiguousarray(frames, np.int32)
        ms = (C.c_float * 3)()
        nl = C.c_int()
        self._ck(self.lib.hevcdl_bench_resident(self.h, _ptr(fr), len(fr), iters, ms, C.byref(nl)), "bench_resident")
        return list(ms), nl.value

    def bench_e2e(self, first_id, iters, depth, frames):
        """frames: list of (Y, U, V) uint8 host planes (same strides).  Returns (seconds, d2h_bytes, checksum)."""
        n = len(frames)
        ys = (C.c_void_p * n)(*[f[0].ctypes.data for f in frames])
        us = (C.c_void_p * n)(*[f[1].ctypes.data for f in frames])
        vs = (C.c_void_p * n)(*[f[2].ctypes.data for f in frames])
        sec, nb, chk = C.c_double(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.hevcdl_bench_e2e(self.h, first_id, iters, depth, n, ys, us, vs, frames[0][0].strides[0],
                                           frames[0][1].strides[0], C.byref(sec), C.byref(nb), C.byref(chk)), "bench_e2e")
        return sec.value, nb.value, chk.value

    def rerun_rmd(self, frame, labels):
        """Test hook: K6 of a finished frame again with the given labels [nctu,16] (e.g. labels read from the reference's files)."""
        lab = np.ascontiguousarray(labels, np.uint8)
        assert lab.shape == (self.nctu, 16)
        self._ck(self.lib.hevcdl_debug_rerun_rmd(self.h, frame, _ptr(lab)), "debug_rerun_rmd")

    def debug_copy(self, which):
        """Tensor-core path intermediates of the last frame (0 cat, 1 a2, 2 features) as raw bf16 bits."""
        n = C.c_size_t()
        self._ck(self.lib.hevcdl_debug_copy(self.h, which, None, 0, C.byref(n)), "debug_copy")
        out = np.empty(n.value // 2, np.uint16)
        self._ck(self.lib.hevcdl_debug_copy(self.h, which, _ptr(out), n.value, C.byref(n)), "debug_copy")
        return out

    def stats(self):
        s = Stats()
        self._ck(self.lib.hevcdl_get_stats(self.h, C.byref(s)), "get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}


class PinnedBuffer:
    """Page-locked host bytes from hevcdl_host_alloc as a numpy uint8 array (`.a`); freed with the object."""

    def __init__(self, nbytes, write_combined=False):
        self.lib = load_library()
        self.p = self.lib.hevcdl_host_alloc(int(nbytes), int(write_combined))
        if not self.p:
            raise HevcdlError("hevcdl_host_alloc(%d) failed" % nbytes)
        self.a = np.ctypeslib.as_array(C.cast(self.p, C.POINTER(C.c_uint8)), shape=(int(nbytes),))

    def close(self):
        if getattr(self, "p", None):
            self.a = None
            self.lib.hevcdl_host_free(self.p)
            self.p = None

    __del__ = close


def numa_bind_thread(device):
    """Pin the calling thread to the CPUs of the device's NUMA node (hevcdl_numa_bind_thread); returns the node."""
    return load_library().hevcdl_numa_bind_thread(int(device))


def frame_to_rank(frame, world_size):
    """All-intra frames shard one-frame-per-GPU: frame f -> rank f mod G."""
    return frame % world_size


def rank_frames(n_frames, rank, world_size):
    return [f for f in range(n_frames) if frame_to_rank(f, world_size) == rank]
