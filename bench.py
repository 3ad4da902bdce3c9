#!/usr/bin/env python
"""bench.py -- intra CTUs/s of the CNN-gated partition hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp32|bf16]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; frames shard across ranks)

A step = one 1920x1080 frame (510 CTUs, BASELINE configs[1]) through K0 -> CNN -> labels -> PU /
work-item plan -> K6 35-mode SATD + ranking.  `value`: planes resident in HBM, CUDA events on the context's
stream, rotating over a pool of distinct frames larger than L2.  `e2e`: the same step through the
C-ABI with pinned HOST buffers: H2D of the frame, kernels, D2H of labels + logits + PU SATD lists,
read on the host through zero-copy views (hevcdl_frame_view_get).
`--impl reference`: the reference's CPU path (torch port of use_model.py's batch-1 forwards + C port
of the RMD pass; the reference files themselves cannot travel to the GPU box) on all host cores.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_pool(synth, w, h, n, rank, kind="mixed"):
    """n distinct frames: a few seeded base frames plus cyclic shifts (content differs per frame)."""
    base = [synth.synth_frame(w, h, rank * 8 + i, kind) for i in range(min(n, 4))]
    pool = []
    for i in range(n):
        Y, U, V = base[i % len(base)]
        s = 2 * (i // len(base)) * 37
        pool.append((np.roll(Y, (s, 2 * s), (0, 1)), np.roll(U, (s // 2, s), (0, 1)), np.roll(V, (s // 2, s), (0, 1))))
    return pool


def cpu_reference_step(torch_model, oracle, pool, nctu, n_ctus, step):
    """The reference's CPU path on n_ctus CTUs of one frame: labels (torch port of use_model.py,
    4 batch-1 forwards per CTU) then the RMD pass for those CTUs (C port)."""
    Y, U, V = pool[step % len(pool)]
    a = (step * n_ctus) % max(1, nctu - n_ctus)
    lab = torch_model.frame_labels(Y, U, V, a, a + n_ctus)
    full = np.zeros((nctu, 16), np.uint8)
    full[a:a + n_ctus] = lab
    oracle.frame_rmd(Y, full, a, a + n_ctus)
    return n_ctus


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    from oracle import oracle
    from oracle.torch_ref import TorchConvNet2
    pkg = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = TorchConvNet2(host.DEFAULT_WEIGHTS)
    w, h = args.width, args.height
    nctu = ((w + 63) // 64) * ((h + 63) // 64)
    pool = make_pool(pkg.synth, w, h, 2, 0)
    n_ctus = args.ref_ctus
    for i in range(args.warmup):
        cpu_reference_step(m, oracle, pool, nctu, n_ctus, i)
    t0 = time.perf_counter()
    done = 0
    for i in range(args.steps):
        done += cpu_reference_step(m, oracle, pool, nctu, n_ctus, i)
    dt = time.perf_counter() - t0
    v = done / dt
    sample = "%d CTUs per step of a %dx%d frame: torch-functional port of use_model.py (4 batch-1 forwards/CTU, train-mode BN) + C port of the RMD pass" % (n_ctus, w, h)
    print(json.dumps({
        "impl": "reference", "metric": "intra CTUs/sec", "value": v, "unit": "CTU/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (w, h, nctu), "arm": "the reference's CPU path on the host cores: %d CTUs of the frame per step" % n_ctus},
        "cpu_baseline": {"value": v, "unit": "CTU/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h = args.width, args.height
    prec = host.PREC_BF16_TC if args.precision == "bf16" else host.PREC_FP32
    pool_n = args.pool
    pool = make_pool(pkg.synth, w, h, pool_n, rank, args.content)
    dp = host.DepthPredictor(w, h, device=local_rank, slots=pool_n, precision=prec, rmd=True, batch=args.batch)
    nctu = dp.nctu
    frame_bytes = w * h * 3 // 2

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

    # ---- resident: upload the pool once ---------------------------------------------------------
    for i, (Y, U, V) in enumerate(pool):
        dp.submit(i, Y, U, V)
    npu_total = 0
    for i in range(pool_n):
        dp.wait(i)
        npu_total += len(dp.pus(i)[0])
    frames = list(range(pool_n))
    dp.bench_resident(frames, max(3, args.warmup))
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms, launches = dp.bench_resident(frames, args.steps)
    barrier()
    ms_total = maxr(ms[0])
    value = world * args.steps * nctu / (ms_total / 1000.0)
    ms_cnn = ms[1] / args.steps                      # dominant kernel: one CNN launch per step
    for i in frames:
        dp.release(i)

    # ---- e2e: pinned host planes -> labels + PU SATD lists back on the host, pipelined ------------
    pinned = []
    for (Y, U, V) in pool[:min(pool_n, 8)]:
        buf = torch.empty(frame_bytes, dtype=torch.uint8).pin_memory()
        a = buf.numpy()
        a[:w * h] = Y.ravel(); a[w * h:w * h * 5 // 4] = U.ravel(); a[w * h * 5 // 4:] = V.ravel()
        pinned.append((buf, a[:w * h].reshape(h, w), a[w * h:w * h * 5 // 4].reshape(h // 2, w // 2),
                       a[w * h * 5 // 4:].reshape(h // 2, w // 2)))
    depth = min(args.depth if args.depth > 0 else 3 * args.batch, pool_n)
    d2h_bytes = [0]

    chk = [0]

    def consume(f):
        # the step's results, read on the host: zero-copy views over the context's pinned buffers (the D2H copies
        # themselves were queued by the library behind the kernels)
        v = dp.view(f)
        d2h_bytes[0] += sum(v[k].nbytes for k in ("labels", "logits", "ctu_off", "pus", "satd", "cand"))
        chk[0] += int(v["labels"][-1, -1]) + (int(v["cand"][-1, 0]) if len(v["cand"]) else 0)
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
    # (hevcdl_bench_e2e: submit_frame_u8 / frame_view_get / release_frame, host steady clock).  The headline e2e is (b);
    # (a) is reported as e2e.python_value.
    e2e_iters = max(args.steps, 200)                 # a 40-frame loop lasts ~6 ms: time at least 200 frames and scale
    e2e_steps(max(3, args.warmup) + 2 * depth, 1000)
    d2h_bytes[0] = 0
    barrier()
    t0 = time.perf_counter()
    e2e_steps(e2e_iters, 2000)
    torch.cuda.synchronize()
    dt_py = maxr(time.perf_counter() - t0) * args.steps / e2e_iters
    barrier()
    planes = [(Y, U, V) for (_, Y, U, V) in pinned]
    dp.bench_e2e(3000, max(3, args.warmup) + 2 * depth, depth, planes)
    barrier()
    sec, nb, _ = dp.bench_e2e(4000, e2e_iters, depth, planes)
    dt = maxr(sec) * args.steps / e2e_iters
    d2h_bytes[0] = nb * args.steps // e2e_iters
    barrier()
    clocks = sampler.stop()                          # sampled every 10 ms over both timed regions (resident and e2e)
    e2e_py_value = world * args.steps * nctu / dt_py
    e2e_value = world * args.steps * nctu / dt
    st = dp.stats()
    dp.close()

    pk = peaks()
    ach_tflops = FLOP_PER_CTU * nctu / (ms_cnn / 1000.0) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.precision)
    out = {
        "metric": "intra CTUs/sec", "value": value, "unit": "CTU/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (w, h, nctu), "arm": "one B200 per rank",
                   "precision": args.precision, "content": args.content, "frames_per_cnn_launch": args.batch, "frames_sharded": "frame f -> rank f mod N, no data-path collective",
                   "l2": "inputs rotate over %d resident frames per rank (%.0f MB planes + outputs > 126 MB L2)" % (pool_n, pool_n * frame_bytes / 1e6),
                   "pus_per_frame": npu_total / pool_n},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "CTU/s", "h2d_bytes_per_step": frame_bytes,
                "d2h_bytes_per_step": d2h_bytes[0] // args.steps, "pipeline_depth": depth, "frames_timed": e2e_iters,
                "python_value": e2e_py_value},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": ach_tflops, "peak": pk["tensor"], "unit": "TFLOP/s",
                     "frac": ach_tflops / pk["tensor"], "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained",
                     "kernel": "CNN stage = k_tc_l1 + k_tc_conv2 + k_tc_conv3 + k_tc_fc (tcgen05)" if prec else "k_cnn_fp32", "kernel_ms": ms_cnn,
                     "hbm_achieved_gbs": BYTES_PER_CTU * nctu / (ms_cnn / 1000.0) / 1e9, "hbm_peak_gbs": pk["hbm"],
                     "stage_ms": {"cnn": ms_cnn, "rmd": ms[2] / args.steps},
                     # K6 is not a contraction: algorithmic integer ops against the CUDA-core issue peak (SMs x 128 lanes x clock)
                     "rmd_alu": {"achieved_tiops": INTOP_PER_CTU * nctu / (ms[2] / args.steps / 1000.0) / 1e12,
                                 "peak_tiops": 148 * 128 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, pkg, host)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, pkg, host):
    import torch
    from oracle import oracle
    from oracle.torch_ref import TorchConvNet2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = TorchConvNet2(host.DEFAULT_WEIGHTS)
    w, h = args.width, args.height
    nctu = ((w + 63) // 64) * ((h + 63) // 64)
    pool = make_pool(pkg.synth, w, h, 1, 0)
    cpu_reference_step(m, oracle, pool, nctu, 4, 0)           # warm-up
    t0 = time.perf_counter()
    done, i = 0, 0
    while time.perf_counter() - t0 < args.cpu_seconds:
        done += cpu_reference_step(m, oracle, pool, nctu, 16, i)
        i += 1
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "CTU/s", "cores": cores, "kind": "port",
            "sample": "%d CTUs of one %dx%d frame in %.1f s: torch-functional port of use_model.py (4 batch-1 forwards/CTU, train-mode BN, %d threads) + C port of the RMD pass" % (done, w, h, dt, cores)}


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
    ap.add_argument("--batch", type=int, default=4, help="frames per CNN launch (hevcdl_cfg.batch); results do not depend on it")
    ap.add_argument("--ref-ctus", type=int, default=24)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
