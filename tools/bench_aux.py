#!/usr/bin/env python
"""Device time, end-to-end call time and CPU-port time of the entry points beside the hot path:
hevcdl_tu_code / hevcdl_tu_code_rdoq, hevcdl_deblock_frame, hevcdl_sao_stats -- one 1920x1080 (and 3840x2160) picture's worth
of work per call.  Device time = CUDA events around the kernels of the call (hevcdl_last_aux_ms), median of `--reps` calls;
call time = host clock around the C-ABI call (pack + H2D + kernels + D2H + unpack).  Prints one JSON object; the oracle legs
(single-threaded C ports, oracle/) are the checker AND the CPU baseline here, exactly as in bench.py's cpu_baseline.
    python tools/bench_aux.py --out gpurun_out/r02n_aux.json"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 0)) or 6536.0, "MEASURED_PEAKS.json"
    return 6536.0, "fallback (B200_PROFILING.md)"


def tu_map(rng, W, H):
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
    return tu


def picture(rng, W, H):
    base = np.kron(rng.integers(30, 226, ((H + 7) // 8, (W + 7) // 8)), np.ones((8, 8)))[:H, :W]
    Y = np.clip(base + rng.integers(-3, 4, (H, W)), 0, 255).astype(np.uint8)
    U = np.clip(np.kron(rng.integers(60, 196, ((H + 15) // 16, (W + 15) // 16)), np.ones((8, 8)))[:H // 2, :W // 2] + rng.integers(-2, 3, (H // 2, W // 2)), 0, 255).astype(np.uint8)
    V = np.clip(U.astype(np.int32)[::-1, ::-1] + 7, 0, 255).astype(np.uint8)
    return Y, U, V


def med(v):
    return float(np.median(v))


def timed(dp, fn, reps):
    wall, dev = [], []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        wall.append((time.perf_counter() - t) * 1e3)
        dev.append(dp.last_aux_ms())
    return r, med(wall[1:]), med(dev[1:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--sizes", default="1920x1080,3840x2160")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
    from oracle import oracle
    hbm, src = peaks()
    rep = {"hbm_peak_gbs": hbm, "peak_source": src, "reps": a.reps, "pictures": {}}
    dp = host.DepthPredictor(64, 64, precision=host.PREC_FP32, rmd=False, outputs=0)
    rng = np.random.default_rng(5)
    g = np.load(os.path.join(ROOT, "tests", "golden", "tq_rdoq_192x128_qp32.npz"))
    for sz in a.sizes.split(","):
        W, H = (int(v) for v in sz.split("x"))
        H8 = H // 8 * 8
        out = rep["pictures"][sz] = {}
        # ---- deblocking ------------------------------------------------------------------------------
        Y, U, V = picture(rng, W, H8)
        tu = tu_map(rng, W, H8)
        qp = np.full(tu.shape, 32, np.int8)
        got, wall, dev = timed(dp, lambda: dp.deblock_frame(Y, U, V, tu, qp), a.reps)
        byts = 12 * W * H8           # int16 samples, luma + chroma read and written once per pass, two passes
        d = out["deblock"] = {"device_ms": dev, "call_ms": wall, "algorithmic_mb": byts / 1e6, "achieved_gbs": byts / dev / 1e6,
                              "hbm_frac": byts / dev / 1e6 / hbm, "samples_changed": int((got[0] != Y).sum())}
        if not a.no_cpu:
            t = time.perf_counter()
            want = oracle.deblock_frame(Y, U, V, tu, qp)
            d["cpu_port_ms"] = (time.perf_counter() - t) * 1e3
            d["identical"] = bool(all((x == y).all() for x, y in zip(got, want)))
        # ---- SAO statistics ---------------------------------------------------------------------------
        org = picture(rng, W, H8)
        src_ = [np.clip(p.astype(np.int32) + rng.integers(-4, 5, p.shape), 0, 255).astype(np.uint8) for p in org]
        got, wall, dev = timed(dp, lambda: dp.sao_stats(org, src_), a.reps)
        byts = 6 * W * H8            # two int16 pictures read once
        d = out["sao_stats"] = {"device_ms": dev, "call_ms": wall, "algorithmic_mb": byts / 1e6, "achieved_gbs": byts / dev / 1e6,
                                "hbm_frac": byts / dev / 1e6 / hbm}
        if not a.no_cpu:
            t = time.perf_counter()
            want = oracle.sao_stats(org, src_)
            d["cpu_port_ms"] = (time.perf_counter() - t) * 1e3
            d["identical"] = bool((got == want).all())
        # ---- SAO application ---------------------------------------------------------------------------
        nctu = ((W + 63) // 64) * ((H8 + 63) // 64)
        ty = rng.integers(-1, 5, (nctu, 3)).astype(np.int8)
        of = rng.integers(-7, 8, (nctu, 3, 32)).astype(np.int8)
        got, wall, dev = timed(dp, lambda: dp.sao_apply(src_, ty, of), a.reps)
        byts = 6 * W * H8            # one int16 picture read, one written
        d = out["sao_apply"] = {"device_ms": dev, "call_ms": wall, "algorithmic_mb": byts / 1e6, "achieved_gbs": byts / dev / 1e6,
                                "hbm_frac": byts / dev / 1e6 / hbm}
        if not a.no_cpu:
            t = time.perf_counter()
            want = oracle.sao_apply(src_, ty, of)
            d["cpu_port_ms"] = (time.perf_counter() - t) * 1e3
            d["identical"] = bool(all((x == y).all() for x, y in zip(got, want)))
        # ---- the three in-loop passes through the C-ABI proper (prepared int16 planes, no numpy conversions in the timed
        #      region): three separate calls against hevcdl_inloop_frame + hevcdl_sao_apply on the resident picture
        lib, hctx = dp.lib, dp.h
        P16 = lambda planes: [np.ascontiguousarray(p, np.int16) for p in planes]
        rec0, orgp = P16((Y, U, V)), P16(org)
        tuf, qpf = np.ascontiguousarray(tu, np.uint8).ravel(), np.ascontiguousarray(qp, np.int8).ravel()
        prm = np.zeros((nctu, 3), host.SAO_PARAM_DTYPE)
        prm["type"] = ty; prm["offset"] = of
        st_buf = np.zeros((nctu, 3, 5, 2, 32), np.int64)
        res = [np.zeros_like(p) for p in rec0]
        vp = lambda a_: C.c_void_p(a_.ctypes.data)

        def separate():
            r = [p.copy() for p in rec0]
            t0 = time.perf_counter()
            lib.hevcdl_deblock_frame(hctx, vp(r[0]), W, vp(r[1]), vp(r[2]), W // 2, W, H8, vp(tuf), vp(qpf), 0, 0, 0, 0)
            lib.hevcdl_sao_stats(hctx, vp(orgp[0]), vp(orgp[1]), vp(orgp[2]), W, W // 2, vp(r[0]), vp(r[1]), vp(r[2]), W, W // 2, W, H8, vp(st_buf))
            lib.hevcdl_sao_apply(hctx, vp(r[0]), vp(r[1]), vp(r[2]), W, W // 2, vp(res[0]), vp(res[1]), vp(res[2]), W, W // 2, W, H8, vp(prm))
            return (time.perf_counter() - t0) * 1e3, [p.copy() for p in res], st_buf.copy()

        def fused():
            r = [p.copy() for p in rec0]
            t0 = time.perf_counter()
            lib.hevcdl_inloop_frame(hctx, vp(r[0]), W, vp(r[1]), vp(r[2]), W // 2, W, H8, vp(tuf), vp(qpf), 0, 0, 0, 0,
                                    vp(orgp[0]), vp(orgp[1]), vp(orgp[2]), W, W // 2, vp(st_buf))
            lib.hevcdl_sao_apply(hctx, None, None, None, W, W // 2, vp(res[0]), vp(res[1]), vp(res[2]), W, W // 2, W, H8, vp(prm))
            return (time.perf_counter() - t0) * 1e3, [p.copy() for p in res], st_buf.copy()
        # the same fused pair with every plane in page-locked memory (hevcdl_host_register, what hm_plugin does with HM's picture
        # buffers): planes go straight between the caller's rows and the device, no packing pass
        rp = [p.copy() for p in rec0]
        resp = [np.zeros_like(p) for p in rec0]
        pinned_ok = all(lib.hevcdl_host_register(vp(x), x.nbytes) == 0 for x in rp + resp + orgp)

        def fused_pinned():
            for dst, src0 in zip(rp, rec0):
                np.copyto(dst, src0)
            t0 = time.perf_counter()
            lib.hevcdl_inloop_frame(hctx, vp(rp[0]), W, vp(rp[1]), vp(rp[2]), W // 2, W, H8, vp(tuf), vp(qpf), 0, 0, 0, 0,
                                    vp(orgp[0]), vp(orgp[1]), vp(orgp[2]), W, W // 2, vp(st_buf))
            lib.hevcdl_sao_apply(hctx, None, None, None, W, W // 2, vp(resp[0]), vp(resp[1]), vp(resp[2]), W, W // 2, W, H8, vp(prm))
            return (time.perf_counter() - t0) * 1e3, [p.copy() for p in resp], st_buf.copy()
        ts, tf, tp = [], [], []
        for _ in range(a.reps):
            ms, res_s, st_s = separate(); ts.append(ms)
            ms, res_f, st_f = fused(); tf.append(ms)
            ms, res_p, st_p = fused_pinned(); tp.append(ms)
        for x in rp + resp + orgp:
            lib.hevcdl_host_unregister(vp(x))
        out["inloop_three_passes_c_abi"] = {"separate_calls_ms": med(ts[1:]), "fused_resident_ms": med(tf[1:]),
                                            "fused_resident_page_locked_planes_ms": med(tp[1:]) if pinned_ok else None,
                                            "identical": bool(all((x == y).all() for x, y in zip(res_s, res_f)) and (st_s == st_f).all() and
                                                              all((x == y).all() for x, y in zip(res_s, res_p)) and (st_s == st_p).all()),
                                            "pcie_mb_separate": 37.4 * W * H8 / (1920 * 1080), "pcie_mb_fused": 25.0 * W * H8 / (1920 * 1080)}
        # ---- intra predictor: every 16x16 block of the luma picture x 35 modes -----------------------------
        nblk = (W // 16) * (H8 // 16)
        lines = [rng.integers(0, 256, 65).astype(np.int16) for _ in range(64)]
        ll = [lines[i % 64] for i in range(nblk * 35)]
        mm = [i % 35 for i in range(nblk * 35)]
        got, wall, dev = timed(dp, lambda: dp.intra_pred(ll, mm, [True] * len(ll)), 3)
        out["intra_pred_16x16_x35"] = {"blocks": len(ll), "device_ms": dev, "call_ms_incl_python_marshalling": wall,
                                       "mblocks_per_s": len(ll) / dev / 1e3, "gsamples_per_s": len(ll) * 256 / dev / 1e6}
        # ---- transform-unit core: the luma area of one picture, a quarter of it per TU size -------------
        blocks, qps, flags, rq, ests = [], [], [], [], []
        hdr = g["hdr"]
        luma = [i for i in range(len(hdr)) if hdr[i][2] == 0 and hdr[i][7] == 0 and hdr[i][6] == 0]
        for n in (4, 8, 16, 32):
            cnt = W * H8 // 4 // (n * n)
            lap = np.round(rng.laplace(0, 6, (cnt, n, n))).clip(-255, 255).astype(np.int16)
            cand = [i for i in luma if hdr[i][1] == n]
            for k in range(cnt):
                i = cand[k % len(cand)]
                blocks.append(lap[k]); qps.append(32); flags.append(host.TU_DST if n == 4 else 0)
                rq.append((float(g["lam"][i]), len(ests) % 64, 0, 0, int(hdr[i][8]), 1 | 2))
                if len(ests) < 64:
                    ests.append(g["est"][i])
        nel = sum(b.size for b in blocks)
        _, wall, dev = timed(dp, lambda: dp.tu_code(blocks, qps, flags), max(3, a.reps // 3))
        out["tu_code_flat"] = {"tus": len(blocks), "samples": nel, "device_ms": dev, "call_ms_incl_python_marshalling": wall,
                               "mtu_per_s": len(blocks) / dev / 1e3, "gsamples_per_s": nel / dev / 1e6,
                               "algorithmic_mb": 14 * nel / 1e6, "achieved_gbs": 14 * nel / dev / 1e6, "hbm_frac": 14 * nel / dev / 1e6 / hbm}
        fl = [f | host.TU_RDOQ for f in flags]
        _, wall, dev = timed(dp, lambda: dp.tu_code(blocks, qps, fl, rdoq=np.array(rq, host.TU_RDOQ_DTYPE), est=np.stack(ests)), max(3, a.reps // 3))
        out["tu_code_rdoq"] = {"tus": len(blocks), "samples": nel, "device_ms": dev, "call_ms_incl_python_marshalling": wall,
                               "mtu_per_s": len(blocks) / dev / 1e3, "gsamples_per_s": nel / dev / 1e6}
        if not a.no_cpu:
            cpu = {}
            for n in (4, 8, 16, 32):
                sub = [i for i, b in enumerate(blocks) if b.shape[0] == n][:300]
                t = time.perf_counter()
                for i in sub:
                    oracle.tq_tu(blocks[i], 32, flags[i])
                tf = (time.perf_counter() - t) / len(sub)
                t = time.perf_counter()
                for i in sub:
                    c = oracle.tq_tu(blocks[i], 32, flags[i])[0]
                    oracle.rdoq(c, 0, 0, 32, 0, rq[i][0], ests[rq[i][1]], rq[i][4], 1, 0, 1)
                tr = (time.perf_counter() - t) / len(sub)
                cpu[str(n)] = {"flat_us_per_tu": tf * 1e6, "with_rdoq_us_per_tu": tr * 1e6}
            cnts = {n: W * H8 // 4 // (n * n) for n in (4, 8, 16, 32)}
            out["tu_code_flat"]["cpu_port_ms_extrapolated_incl_ctypes_call_overhead"] = sum(cnts[n] * cpu[str(n)]["flat_us_per_tu"] for n in cnts) / 1e3
            out["tu_code_rdoq"]["cpu_port_ms_extrapolated_incl_ctypes_call_overhead"] = sum(cnts[n] * cpu[str(n)]["with_rdoq_us_per_tu"] for n in cnts) / 1e3
            out["tu_cpu_port_per_size"] = cpu
    dp.close()
    s = json.dumps(rep, indent=1)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
