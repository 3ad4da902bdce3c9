#!/usr/bin/env python
"""bench.py -- intra CTUs/s of the CNN-gated partition hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp32|bf16]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; frames shard across ranks)

A step = one 1920x1080 frame (510 CTUs, BASELINE configs[1]) through K0 -> CNN -> labels -> PU /
work-item plan -> K6 35-mode SATD + ranking.  The K steps are repeated back to back inside ONE timed region until at
least 200 frames have been timed (a 20-frame region lasts 2.5 ms: too short to be stable); `frames_timed` says how many.
`value`: planes resident in HBM, CUDA events on the context's stream, rotating over a pool of distinct frames larger
than L2.  `e2e`: the same step through the C-ABI with pinned HOST buffers: H2D of the frame, kernels, D2H of labels +
PU list + candidate modes (hevcdl_cfg.outputs = 0: what an encoder consumes), read on the host through zero-copy views.
`parity`: the benchmarked precision (and its fp32 sibling) against a full-frame pass of the CPU oracle, in this run.
`--impl reference`: the reference's CPU path (torch port of use_model.py's batch-1 forwards incl. crop and per-CTU
file write on all host threads + single-threaded C port of the RMD pass, as HM runs it; the reference files themselves
cannot travel to the GPU box) and, where oracle/_ref travelled, the reference's own encoder binaries.
"""
import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "hevc-deep-learning-pipeline_b200"

FLOP_PER_CTU = 99.49e6        # CNN MACs*2 with conv64 evaluated once per CTU
WORKLOAD = "1 frame %dx%d all-intra QP32 per step (%d CTUs), CNN labels + 35-mode SATD (RMD)"   # BASELINE.json configs[1]
INTOP_PER_CTU = 1.72e6        # ~420 integer ops per luma pixel for the 35-mode RMD pass (NxN trials not counted)
BYTES_PER_CTU = 6144 + 400    # 64x64 Y + 2x32x32 C in, labels + candidate lists out
MIN_FRAMES_TIMED = 200


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1650.0, "tensor_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "10"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def make_pool(synth, w, h, n, rank, kind="mixed", nbase=4):
    """n distinct frames: a few seeded base frames plus cyclic shifts (content differs per frame)."""
    base = [synth.synth_frame(w, h, rank * 8 + i, kind) for i in range(min(n, nbase))]
    pool = []
    for i in range(n):
        Y, U, V = base[i % len(base)]
        s = 2 * (i // len(base)) * 37
        pool.append((np.roll(Y, (s, 2 * s), (0, 1)), np.roll(U, (s // 2, s), (0, 1)), np.roll(V, (s // 2, s), (0, 1))))
    return pool


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's path on the host cores (oracle/ is the checker / baseline, never the product)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_sample(w, h, seconds, ctus_per_step=16):
    """Bounded sample of the reference's CPU hot path on one 1080p frame, as the reference runs it:
    (a) sidecar: crop + ToTensor + 4 batch-1 forwards per CTU + label rules + per-CTU file write (use_model.py:86-127) on
        all host threads (torch intra-op, as the reference's torch would); (b) the RMD pass of the same CTUs in ONE thread
        (it lives inside HM's single encoder thread).  value = CTUs / (t_a + t_b)."""
    import torch
    from oracle import oracle
    from oracle.torch_ref import TorchConvNet2
    pkg = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    cores = len(os.sched_getaffinity(0)) or 1
    torch.set_num_threads(cores)
    m = TorchConvNet2(host.DEFAULT_WEIGHTS)
    nctu = ((w + 63) // 64) * ((h + 63) // 64)
    Y, U, V = pkg.synth.synth_frame(w, h, 0)
    td = tempfile.mkdtemp(prefix="hevcdl_cpu_")
    try:
        m.sidecar_ctus(Y, U, V, 0, 4, td)                                   # warm-up
        oracle.set_threads(1)
        t_cnn = t_rmd = 0.0
        done = 0
        a = 0
        t_start = time.perf_counter()
        while time.perf_counter() - t_start < seconds:
            b = min(nctu, a + ctus_per_step)
            t0 = time.perf_counter()
            lab = m.sidecar_ctus(Y, U, V, a, b, td)
            t1 = time.perf_counter()
            full = np.zeros((nctu, 16), np.uint8)
            full[a:b] = lab
            oracle.frame_rmd(Y, full, a, b)
            t2 = time.perf_counter()
            t_cnn += t1 - t0; t_rmd += t2 - t1; done += b - a
            a = b % nctu
        oracle.set_threads(0)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    return {"value": done / (t_cnn + t_rmd), "unit": "CTU/s", "cores": cores, "kind": "port",
            "sidecar_ctus_s": done / t_cnn, "rmd_1thread_ctus_s": done / t_rmd,
            "sample": "%d CTUs of one %dx%d frame in %.1f s: torch-functional port of use_model.py incl. crop + per-CTU file write "
                      "(4 batch-1 forwards/CTU, train-mode BN, %d threads) + C port of the RMD pass in 1 thread (as inside HM)"
                      % (done, w, h, t_cnn + t_rmd, cores)}


def reference_binaries_sample(w, h, labels, frame, qp=32):
    """The reference's own encoder binaries (compiled from /root/reference by oracle/Makefile into oracle/_ref, which
    travels to the GPU box) on ONE frame: HM_dl with its labels already on disk, and the full-search anchor.  HM's own
    `Total Time` (encmain.cpp:113: clock(), one core).  Returns {} when the binaries did not travel."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hm_util
    if not hm_util.have("ref", "anchor"):
        return {}
    nctu = ((w + 63) // 64) * ((h + 63) // 64)
    out = {}
    td = tempfile.mkdtemp(prefix="hevcdl_hm_")
    try:
        hm_util.write_yuv(os.path.join(td, "in.yuv"), [frame])
        hm_util.write_pred(os.path.join(td, "pred"), 0, labels)

        def run(kind, res):
            r = hm_util.encode(kind, td, "in.yuv", w, h, 1, qp, out=kind + ".bin")
            if r["rc"] == 0 and r.get("seconds"):
                res[kind] = r
        res = {}
        th = [threading.Thread(target=run, args=(k, res)) for k in ("ref", "anchor")]     # one core each, side by side
        for t in th:
            t.start()
        for t in th:
            t.join()
        if "ref" in res:
            out["hm_dl_ctus_s"] = nctu / res["ref"]["seconds"]
            out["hm_dl_seconds"] = res["ref"]["seconds"]
        if "anchor" in res:
            out["anchor_ctus_s"] = nctu / res["anchor"]["seconds"]
            out["anchor_seconds"] = res["anchor"]["seconds"]
        out["encoder_sample"] = "oracle/_ref/TAppEncoder_ref (UNMODIFIED reference, labels on disk) and TAppEncoder_anchor (full search), 1 frame %dx%d QP%d, HM Total Time, 1 core each" % (w, h, qp)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle
    pkg = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    w, h = args.width, args.height
    nctu = ((w + 63) // 64) * ((h + 63) // 64)
    n_ctus = args.ref_ctus
    # K steps of n_ctus CTUs each, W warm-up steps of the same size
    import torch
    from oracle.torch_ref import TorchConvNet2
    cores = len(os.sched_getaffinity(0)) or 1
    torch.set_num_threads(cores)
    m = TorchConvNet2(host.DEFAULT_WEIGHTS)
    Y, U, V = pkg.synth.synth_frame(w, h, 0)
    td = tempfile.mkdtemp(prefix="hevcdl_ref_")
    oracle.set_threads(1)

    def step(i):
        a = (i * n_ctus) % max(1, nctu - n_ctus)
        lab = m.sidecar_ctus(Y, U, V, a, a + n_ctus, td)
        full = np.zeros((nctu, 16), np.uint8)
        full[a:a + n_ctus] = lab
        oracle.frame_rmd(Y, full, a, a + n_ctus)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    shutil.rmtree(td, ignore_errors=True)
    oracle.set_threads(0)
    v = args.steps * n_ctus / dt
    sample = ("%d CTUs per step of a %dx%d frame: torch-functional port of use_model.py incl. crop + per-CTU file write (4 batch-1 "
              "forwards/CTU, train-mode BN, %d threads) + C port of the RMD pass in 1 thread" % (n_ctus, w, h, cores))
    cb = {"value": v, "unit": "CTU/s", "cores": cores, "kind": "port", "sample": sample}
    if not args.no_ref_binaries:
        labels = oracle.frame_labels(oracle.load_weights(host.DEFAULT_WEIGHTS), Y, U, V)
        cb.update(reference_binaries_sample(w, h, labels, (Y, U, V)))
    print(json.dumps({
        "impl": "reference", "metric": "intra CTUs/sec", "value": v, "unit": "CTU/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (w, h, nctu), "arm": "the reference's CPU path on the host cores: %d CTUs of the frame per step" % n_ctus},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def measure(host, torch, dist, args, w, h, pool, local_rank, world, frames_timed, batch, prec, batch_e2e=0):
    """Resident and end-to-end throughput of one picture size on this rank's GPU; every rank calls it with its own frames."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    pool_n = len(pool)
    frame_bytes = w * h * 3 // 2
    # outputs = 0: what an encoder consumes (labels, PU list, candidate modes); pinned_input: the e2e planes below are
    # page-locked and untouched while in flight; numa_bind: context buffers on the GPU's own NUMA node
    dp = host.DepthPredictor(w, h, device=local_rank, slots=pool_n, precision=prec, rmd=True, batch=batch, outputs=0,
                             pinned_input=True, numa_bind=True)
    nctu = dp.nctu
    # ---- resident: upload the pool once ---------------------------------------------------------
    for i, (Y, U, V) in enumerate(pool):
        dp.submit(i, Y, U, V)
    npu_total = 0
    for i in range(pool_n):
        dp.wait(i)
        npu_total += len(dp.pus(i)[0])
    frames = list(range(pool_n))
    dp.bench_resident(frames, max(3, args.warmup) * batch)
    barrier()
    ms, launches = dp.bench_resident(frames, frames_timed)
    barrier()
    ms_total = maxr(ms[0])
    for i in frames:
        dp.release(i)
    # ---- e2e: pinned host planes -> labels + PU lists + candidates back on the host, pipelined ------------
    # Frames per launch is the caller's choice (hevcdl_cfg.batch; results do not depend on it).  Resident throughput likes
    # long launches (8); the end-to-end pipeline overlaps copies and launches better with shorter ones, which shows when 8
    # ranks share the host's copy bandwidth (8 GPUs end to end: 31.0 M CTU/s with 4 frames per launch, 30.3 M with 8).
    be = batch_e2e if batch_e2e > 0 else batch
    if be != batch:
        dp.close()
        dp = host.DepthPredictor(w, h, device=local_rank, slots=pool_n, precision=prec, rmd=True, batch=be, outputs=0,
                                 pinned_input=True, numa_bind=True)
    pinned = []
    for (Y, U, V) in pool[:min(pool_n, 8)]:
        buf = host.PinnedBuffer(frame_bytes, write_combined=args.wc)      # page-locked by the library's own allocator
        a = buf.a
        a[:w * h] = Y.ravel(); a[w * h:w * h * 5 // 4] = U.ravel(); a[w * h * 5 // 4:] = V.ravel()
        pinned.append((buf, a[:w * h].reshape(h, w), a[w * h:w * h * 5 // 4].reshape(h // 2, w // 2),
                       a[w * h * 5 // 4:].reshape(h // 2, w // 2)))
    depth = min(args.depth if args.depth > 0 else 3 * be, pool_n)
    d2h = [0]

    def consume(f):
        v = dp.view(f)
        d2h[0] += sum(v[k].nbytes for k in ("labels", "logits", "ctu_off", "pus", "satd", "cand"))
        dp.release(f)

    def e2e_steps(n, first_id):
        inflight = []
        for i in range(n):
            _, Y, U, V = pinned[i % len(pinned)]
            dp.submit(first_id + i, Y, U, V)
            inflight.append(first_id + i)
            if len(inflight) >= depth:
                consume(inflight.pop(0))
        for f in inflight:
            consume(f)
    # (a) driven from Python through host.DepthPredictor (ctypes); (b) the same C-ABI calls driven from C
    # (hevcdl_bench_e2e: submit_frame_u8 / frame_view_get / release_frame, host steady clock).  The headline e2e is (b).
    e2e_steps(max(3, args.warmup) + 2 * depth, 1000)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(frames_timed, 100000)
    torch.cuda.synchronize()
    dt_py = maxr(time.perf_counter() - t0)
    barrier()
    planes = [(Y, U, V) for (_, Y, U, V) in pinned]
    dp.bench_e2e(200000, max(3, args.warmup) + 2 * depth, depth, planes)
    barrier()
    sec, nb, _ = dp.bench_e2e(300000, frames_timed, depth, planes)
    dt = maxr(sec)
    barrier()
    st = dp.stats()
    dp.close()
    del pinned
    return {"nctu": nctu, "ms_total": ms_total, "ms_cnn": ms[1], "ms_rmd": ms[2], "launches": launches,
            "value": world * frames_timed * nctu / (ms_total / 1000.0),
            "e2e_value": world * frames_timed * nctu / dt, "e2e_py_value": world * frames_timed * nctu / dt_py,
            "h2d_per_frame": frame_bytes, "d2h_per_frame": nb // frames_timed, "depth": depth, "batch_e2e": be,
            "pus_per_frame": npu_total / pool_n, "stats": st}


def copy_probe(host, torch, dist, world, frame_bytes, reps=64):
    """All ranks copy at once (the situation of the e2e leg): per-rank H2D and D2H GB/s from / to pinned host memory
    (h2d_wc: from write-combined pinned memory)."""
    src = torch.empty(frame_bytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(frame_bytes, dtype=torch.uint8, device="cuda")
    wcb = host.PinnedBuffer(frame_bytes, write_combined=True)
    wcb.a[:] = 7
    wc = torch.from_numpy(wcb.a)
    out = {}
    for name, (a, b) in (("h2d", (dst, src)), ("d2h", (src, dst)), ("h2d_wc", (dst, wc))):
        for _ in range(4):
            a.copy_(b, non_blocking=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a.copy_(b, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = reps * frame_bytes / (e0.elapsed_time(e1) / 1e3) / 1e9
        if world > 1:
            t = torch.tensor([gbs, -gbs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            out[name + "_gbs_per_rank_min"], out[name + "_gbs_per_rank_max"] = float(t[0]), float(-t[1])
        else:
            out[name + "_gbs_per_rank_min"] = out[name + "_gbs_per_rank_max"] = gbs
    del wc
    wcb.close()
    return out


def parity_block(host, oracle, Y, U, V, local_rank):
    """The benchmarked precision and its fp32 sibling against a FULL-FRAME pass of the CPU oracle (510 CTUs at 1080p)."""
    w_ = oracle.load_weights(host.DEFAULT_WEIGHTS)
    olab, olg, mar = oracle.frame_labels(w_, Y, U, V, want_logits=True)
    h, w = Y.shape
    out = {"frame": "%dx%d synth_frame(0): every CTU" % (w, h)}
    labels = {}
    for name, prec, eps in (("bf16", host.PREC_BF16_TC, 0.1), ("fp32", host.PREC_FP32, 1e-3)):
        dp = host.DepthPredictor(w, h, device=local_rank, precision=prec, rmd=True, outputs=host.OUT_LOGITS | host.OUT_SATD)
        dp.submit(0, Y, U, V)
        lab, lg = dp.labels(0, want_logits=True)
        pus, satd, cand = dp.pus(0)
        dp.release(0)
        dp.close()
        rep = oracle.label_parity(lab, olab, mar, eps, lg, olg)
        opu, osatd = oracle.frame_rmd(Y, lab)                              # K6 against the oracle for the labels used
        rep["rmd_pus"] = int(len(pus))
        rep["rmd_satd_mismatches"] = int((osatd != satd).sum()) if len(opu) == len(pus) else -1
        rep["rmd_cand0_is_argmin"] = bool((cand[:, 0] == osatd.argmin(axis=1)).all()) if len(opu) == len(pus) else False
        out[name] = rep
        labels[name] = lab
    return out, labels


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    torch.cuda.set_device(local_rank)
    numa_node = host.numa_bind_thread(local_rank)     # before any pinned allocation of this process (torch's included)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h = args.width, args.height
    prec = host.PREC_BF16_TC if args.precision == "bf16" else host.PREC_FP32
    frames_timed = -(-max(args.steps, MIN_FRAMES_TIMED) // args.steps) * args.steps
    pool = make_pool(pkg.synth, w, h, args.pool, rank, args.content)
    sampler = ClockSampler(local_rank)
    sampler.start()
    r = measure(host, torch, dist, args, w, h, pool, local_rank, world, frames_timed, args.batch, prec, args.batch_e2e)
    clocks = sampler.stop()                          # sampled every 10 ms over both timed regions (resident and e2e)
    nctu = r["nctu"]
    frame_bytes = w * h * 3 // 2

    pk = peaks()
    ms_cnn = r["ms_cnn"] / frames_timed              # CNN stage per frame (K1-K4: the tensor-core kernels)
    ms_rmd = r["ms_rmd"] / frames_timed
    ms_step = r["ms_total"] / frames_timed
    ach_cnn = FLOP_PER_CTU * nctu / (ms_cnn / 1000.0) / 1e12
    ach_fused = FLOP_PER_CTU * nctu / (ms_step / 1000.0) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get(args.precision + "_by_frames_per_launch", {}).get(str(args.batch), tj.get(args.precision) if args.batch == 4 else None)
    out = {
        "metric": "intra CTUs/sec", "value": r["value"], "unit": "CTU/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "frames_timed": frames_timed, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (w, h, nctu), "arm": "one B200 per rank",
                   "precision": args.precision, "content": args.content, "frames_per_cnn_launch": args.batch,
                   "frames_sharded": "frame f -> rank f mod N, no data-path collective",
                   "timed_region": "the %d steps repeated back to back until %d frames are timed in one region" % (args.steps, frames_timed),
                   "l2": "inputs rotate over %d resident frames per rank (%.0f MB planes + outputs > 126 MB L2)" % (args.pool, args.pool * frame_bytes / 1e6),
                   "pus_per_frame": r["pus_per_frame"], "numa_node": numa_node},
        "clocks": clocks,
        "e2e": {"value": r["e2e_value"], "unit": "CTU/s", "h2d_bytes_per_step": r["h2d_per_frame"],
                "d2h_bytes_per_step": r["d2h_per_frame"], "pipeline_depth": r["depth"], "frames_per_cnn_launch": r["batch_e2e"], "frames_timed": frames_timed,
                "python_value": r["e2e_py_value"], "host_planes": "page-locked (hevcdl_host_alloc%s), hevcdl_cfg.pinned_input = 1" % (", write-combined" if args.wc else ""), "outputs": "labels + PU list + ranked candidate modes (hevcdl_cfg.outputs = 0)"},
        "gpu_launches": r["launches"],
        "roofline": {"bound": "tensor", "achieved": ach_cnn, "peak": pk["tensor_burst"], "unit": "TFLOP/s",
                     "frac": ach_cnn / pk["tensor_burst"], "frac_sustained": ach_cnn / pk["tensor_sustained"],
                     "peak_sustained": pk["tensor_sustained"], "traffic": traffic,
                     "traffic_over_algorithmic": (traffic / (BYTES_PER_CTU * nctu)) if traffic else None,
                     "traffic_note": "DRAM bytes of the CNN kernels per frame in the pipeline's natural cache state (ncu --cache-control none, profiles/traffic.json): the bf16 intermediates of a multi-frame launch exceed L2; null if not measured for this --batch",
                     "peak_source": pk["src"] + " bf16: burst (a %.0f ms region at full clocks); frac_sustained is against the long-run figure" % r["ms_total"],
                     "kernel": "CNN stage = k_tc_l1 + k_tc_conv2 + k_tc_conv3 + k_tc_fc (tcgen05)" if prec else "k_cnn_fp32", "kernel_ms": ms_cnn,
                     "fused_path": {"achieved": ach_fused, "frac": ach_fused / pk["tensor_burst"], "frac_sustained": ach_fused / pk["tensor_sustained"],
                                    "what": "CNN FLOPs over the whole step (CNN + plan + K6), the north-star's fused CNN+SATD path"},
                     "hbm_achieved_gbs": BYTES_PER_CTU * nctu / (ms_step / 1000.0) / 1e9, "hbm_peak_gbs": pk["hbm"],
                     "stage_ms": {"cnn": ms_cnn, "rmd": ms_rmd},
                     # K6 is not a contraction: algorithmic integer ops against the CUDA-core issue peak (SMs x 128 lanes x clock)
                     "rmd_alu": {"achieved_tiops": INTOP_PER_CTU * nctu / (ms_rmd / 1000.0) / 1e12,
                                 "peak_tiops": 148 * 128 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12}},
    }
    bp = os.path.join(ROOT, "profiles", "r02_bdrate_1080p_100f.json")
    if os.path.exists(bp):        # the other half of BASELINE.json's metric: NOT measured in this run (hours of encoder time), read from the committed sweep
        bd = json.load(open(bp))
        pick = {"hm_dl_vs_anchor": "reference_hm_dl_vs_anchor", "dropin_bf16_gpu_rmd_fix0_vs_hm_dl": "dropin_bf16_labels_gpu_rmd_vs_reference_hm_dl",
                "dropin_bf16_gpu_rmd_fix0_vs_anchor": "dropin_bf16_labels_gpu_rmd_vs_anchor", "hm_dl_labels_bf16_fix0_vs_hm_dl": "bf16_labels_in_reference_encoder_vs_reference_hm_dl",
                "dropin_bf16_gpu_rmd_fix1_vs_anchor": "dropin_boundary_fix_vs_anchor"}
        out["bd_rate"] = {"source": "profiles/r02_bdrate_1080p_100f.json (tools/bdrate_100f.py): BASELINE configs[2], %dx%d, %d frames, QP %s; committed sweep, not re-run by bench.py" % (bd["width"], bd["height"], bd["frames"], bd["qps"]),
                          "tolerance": "north_star: within +-1 % BD-rate / +-0.05 dB BD-PSNR of the reference HM_dl"}
        for k, name in pick.items():
            if k in bd.get("bd", {}):
                out["bd_rate"][name] = {"bd_rate_y_pct": round(bd["bd"][k]["bd_rate_y_pct"], 3), "bd_psnr_y_db": round(bd["bd"][k]["bd_psnr_y_db"], 4)}
    bp8 = os.path.join(ROOT, "profiles", "r02_bdrate_8k_8f.json")
    if os.path.exists(bp8) and "bd_rate" in out:     # BASELINE configs[4], same procedure on 8 frames 7680x4320
        bd8 = json.load(open(bp8))
        out["bd_rate"]["config_8k"] = {"source": "profiles/r02_bdrate_8k_8f.json: %dx%d, %d frames, QP %s" % (bd8["width"], bd8["height"], bd8["frames"], bd8["qps"])}
        for k, name in (("hm_dl_vs_anchor", "reference_hm_dl_vs_anchor"), ("dropin_bf16_gpu_rmd_fix0_vs_hm_dl", "dropin_bf16_labels_gpu_rmd_vs_reference_hm_dl")):
            if k in bd8.get("bd", {}):
                out["bd_rate"]["config_8k"][name] = {"bd_rate_y_pct": round(bd8["bd"][k]["bd_rate_y_pct"], 3), "bd_psnr_y_db": round(bd8["bd"][k]["bd_psnr_y_db"], 4)}
    if world > 1 or args.probe:
        out["copy_probe"] = copy_probe(host, torch, dist, world, frame_bytes)
        out["copy_probe"]["what"] = "all %d ranks copying one frame's planes at once from / to pinned host memory, per-rank GB/s" % world
    if (world > 1 or args.config4k) and not args.no_config4k:
        # BASELINE configs[3]: 3840x2160, 240 frames, frame f -> rank f mod N
        w4, h4, total = 3840, 2160, 240
        per_rank = -(-total // world)
        pool4 = make_pool(pkg.synth, w4, h4, 12, rank, args.content, nbase=2)
        r4 = measure(host, torch, dist, args, w4, h4, pool4, local_rank, world, per_rank, 2, prec)
        out["config_4k"] = {"workload": "3840x2160 all-intra, %d frames sharded f mod %d (%d per rank), %d CTUs per frame" % (per_rank * world, world, per_rank, r4["nctu"]),
                            "value": r4["value"], "e2e": r4["e2e_value"], "unit": "CTU/s", "ms_per_frame": r4["ms_total"] / per_rank,
                            "h2d_bytes_per_frame": r4["h2d_per_frame"], "d2h_bytes_per_frame": r4["d2h_per_frame"],
                            "frames_per_cnn_launch": 2, "pus_per_frame": r4["pus_per_frame"]}
    if rank == 0 and world == 1 and not args.no_parity:
        from oracle import oracle
        out["parity"], labels = parity_block(host, oracle, *pool[0], local_rank)
        # fp32 sibling of the headline: the same step with the CUDA-core fp32 CNN (the bit-tight precision)
        dpf = host.DepthPredictor(w, h, device=local_rank, slots=8, precision=host.PREC_FP32, rmd=True, outputs=0)
        for i in range(8):
            dpf.submit(i, *pool[i])
        for i in range(8):
            dpf.wait(i)
        dpf.bench_resident(list(range(8)), 3)
        msf, _ = dpf.bench_resident(list(range(8)), 40)
        dpf.close()
        out["fp32_sibling"] = {"value": 40 * nctu / (msf[0] / 1000.0), "unit": "CTU/s", "ms_per_step": msf[0] / 40, "frames_timed": 40}
        if not args.no_cpu_baseline:
            cb = cpu_port_sample(w, h, args.cpu_seconds)
            if not args.no_ref_binaries:
                cb.update(reference_binaries_sample(w, h, labels["fp32"], pool[0]))
            out["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HEVCDL_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--pool", type=int, default=48)
    ap.add_argument("--content", default="mixed", choices=["mixed", "noise", "flat"], help="synthetic content (SURVEY.md 8(d)); noise / flat are the stress cases")
    ap.add_argument("--depth", type=int, default=0, help="frames in flight in the e2e measurement (0: three launch batches)")
    ap.add_argument("--batch", type=int, default=8, help="frames per CNN launch (hevcdl_cfg.batch); results do not depend on it")
    ap.add_argument("--batch-e2e", type=int, default=4, help="frames per CNN launch in the end-to-end leg (0: same as --batch)")
    ap.add_argument("--ref-ctus", type=int, default=24)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-binaries", action="store_true", help="skip timing oracle/_ref/TAppEncoder_{ref,anchor} (about 15 s)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--wc", action="store_true", help="write-combined pinned frame buffers in the e2e leg")
    ap.add_argument("--probe", action="store_true", help="add the host<->device copy probe at N=1 too")
    ap.add_argument("--config4k", action="store_true", help="add the 3840x2160 x 240 frames block at N=1 too (always on for N>1)")
    ap.add_argument("--no-config4k", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
