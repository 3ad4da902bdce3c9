#!/usr/bin/env python
"""BD-rate on BASELINE.json configs[2]: 1920x1080 all-intra, 100 frames, QP {22,27,32,37} (README.md:23 procedure of the
reference: encode with the anchor and with the pruned encoder at four QPs, feed rate/PSNR to calc_BDBR).

The sweep is split by what each part needs, every stage appending to one JSON report:

  --stage cpu    (no GPU)  synthetic sequence + fp32 labels of the oracle (= what the reference's sidecar would write),
                           then TAppEncoder_anchor (stock HM decision) and TAppEncoder_ref (UNMODIFIED reference, file
                           handshake) at the four QPs, `--jobs` encodes at a time, one core each
  --stage labels (B200)    labels of the same sequence from libhevcdl.so, fp32 and bf16, boundary_fix 0 and 1 -> .npz
  --stage files  (no GPU)  TAppEncoder_ref fed the label files made from that .npz (bf16 labels through the reference binary)
  --stage dropin (B200)    hm_plugin/_build/TAppEncoder_hevcdl: bf16 labels + batched first-pass SATDs from the device
                           (HEVCDL_RMD=1), boundary_fix 0 and 1
  --stage report           BD numbers (hevc-deep-learning-pipeline_b200/bdrate.py, pinned on JCTVC-B055) from whatever runs exist

usage: python tools/bdrate_100f.py --stage cpu --work /tmp/bd100 --out profiles/r02_bdrate_1080p_100f.json
"""
import argparse
import concurrent.futures as cf
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hm_util  # noqa: E402

pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
bd = importlib.import_module("hevc-deep-learning-pipeline_b200.bdrate")


def load(path):
    return json.load(open(path)) if os.path.exists(path) else {"runs": {}}


def save(rep, path):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    json.dump(rep, open(tmp, "w"), indent=1)
    os.replace(tmp, path)


def sequence(a):
    """The synthetic sequence as one I420 file under --work (made once)."""
    os.makedirs(a.work, exist_ok=True)
    yuv = os.path.join(a.work, "in.yuv")
    fbytes = a.width * a.height * 3 // 2
    if not (os.path.exists(yuv) and os.path.getsize(yuv) == fbytes * a.frames):
        with open(yuv + ".tmp", "wb") as f:
            for i in range(a.frames):
                f.write(pkg.synth.to_i420_bytes(*pkg.synth.synth_frame(a.width, a.height, i)))
        os.replace(yuv + ".tmp", yuv)
    return yuv


def read_frame(yuv, w, h, i):
    fb = w * h * 3 // 2
    m = np.memmap(yuv, np.uint8, "r", offset=i * fb, shape=(fb,))
    return (np.array(m[:w * h]).reshape(h, w), np.array(m[w * h:w * h * 5 // 4]).reshape(h // 2, w // 2),
            np.array(m[w * h * 5 // 4:]).reshape(h // 2, w // 2))


def run_set(a, rep, name, kind, cwd, env=None):
    """One encoder at every QP, --jobs at a time; rows land in rep['runs'][name]."""
    if name in rep["runs"] and len(rep["runs"][name]) == len(a.qp_list):
        print(name, "already done")
        return

    def one(qp):
        t0 = time.time()
        r = hm_util.encode(kind, cwd, os.path.join(a.work, "in.yuv"), a.width, a.height, a.frames, qp,
                           out="%s_%d.bin" % (name, qp), env=env)
        if r["rc"] != 0 or "kbps" not in r:
            raise SystemExit("%s qp %d failed: %s" % (name, qp, (r["stderr"] or r["stdout"])[-800:]))
        os.remove(os.path.join(cwd, "%s_%d.bin" % (name, qp)))
        return {"qp": qp, "kbps": r["kbps"], "psnr_y": r["psnr_y"], "psnr_u": r["psnr_u"], "psnr_v": r["psnr_v"],
                "seconds": r.get("seconds"), "wall_seconds": time.time() - t0, "sha1": r["sha1"], "bytes": r["bytes"]}
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        rows = list(ex.map(one, a.qp_list))
    rep["runs"][name] = rows
    for r in rows:
        print(name, r, flush=True)
    save(rep, a.out)


def pred_dir(a, tag, labels):
    d = os.path.join(a.work, "cwd_" + tag)
    if not os.path.exists(os.path.join(d, "pred", str(a.frames - 1), "ctu0.txt")):
        for f in range(a.frames):
            hm_util.write_pred(os.path.join(d, "pred"), f, labels[f])
    return d


def stage_cpu(a, rep):
    from oracle import oracle
    host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
    yuv = sequence(a)
    lp = os.path.join(a.work, "labels_oracle.npy")
    if os.path.exists(lp):
        labels = np.load(lp)
    elif a.labels_key:
        # large pictures: the fp32 device path's labels (stage `labels`), which equal the oracle's label for label
        # (0 of 816 000 differ over the 100-frame 1080p sequence) -- the oracle needs ~25 ms of CPU per CTU
        labels = np.load(a.labels_npz)[a.labels_key]
        rep["reference_labels"] = "fp32 device labels (%s of %s)" % (a.labels_key, os.path.basename(a.labels_npz))
        np.save(lp, labels)
    else:
        w = oracle.load_weights(host.DEFAULT_WEIGHTS)
        labels = np.stack([oracle.frame_labels(w, *read_frame(yuv, a.width, a.height, i)) for i in range(a.frames)])
        np.save(lp, labels)
    rep.update({"width": a.width, "height": a.height, "frames": a.frames, "qps": a.qp_list, "content": "synth_frame(seed=frame)",
                "label_hist_oracle": np.bincount(labels.ravel(), minlength=4).tolist()})
    d = pred_dir(a, "oracle", labels)
    run_set(a, rep, "hm_dl", "ref", d)
    run_set(a, rep, "anchor", "anchor", d)    # the anchor build still polls ./pred (its labels are ignored: every depth is tried)


def stage_labels(a, rep):
    host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
    yuv = sequence(a)
    out = {}
    for prec, pn in ((host.PREC_FP32, "fp32"), (host.PREC_BF16_TC, "bf16")):
        for fix in (0, 1):
            dp = host.DepthPredictor(a.width, a.height, precision=prec, rmd=False, boundary_fix=bool(fix), slots=2)
            out["%s_fix%d" % (pn, fix)] = np.stack([dp.predict_frame(*read_frame(yuv, a.width, a.height, i), frame=i) for i in range(a.frames)])
            dp.close()
    np.savez_compressed(a.labels_npz, **out)
    print("labels written:", a.labels_npz, {k: v.shape for k, v in out.items()})


def stage_files(a, rep):
    sequence(a)
    z = np.load(a.labels_npz)
    ora = np.load(os.path.join(a.work, "labels_oracle.npy")) if os.path.exists(os.path.join(a.work, "labels_oracle.npy")) else None
    rep["labels_vs_oracle"] = {}
    for k in z.files:
        if ora is not None and not k.endswith("fix1"):
            rep["labels_vs_oracle"][k] = {"labels_differing": int((z[k] != ora).sum()), "ctus_differing": int((z[k] != ora).any(axis=2).sum()),
                                          "labels": int(ora.size), "ctus": int(ora.shape[0] * ora.shape[1])}
    save(rep, a.out)
    for k in ("bf16_fix0", "bf16_fix1", "fp32_fix1"):
        run_set(a, rep, "hm_dl_labels_" + k, "ref", pred_dir(a, k, z[k]))


def stage_dropin(a, rep):
    sequence(a)
    for fix in (0, 1):
        d = os.path.join(a.work, "cwd_dropin%d" % fix)
        os.makedirs(d, exist_ok=True)
        run_set(a, rep, "dropin_bf16_gpu_rmd_fix%d" % fix, "hevcdl", d,
                env={"HEVCDL_PRECISION": "bf16", "HEVCDL_RMD": "1", "HEVCDL_BOUNDARY_FIX": str(fix)})


def stage_report(a, rep):
    def curve(n):
        return np.array([x["kbps"] for x in rep["runs"][n]]), np.array([x["psnr_y"] for x in rep["runs"][n]])

    def cmp(test, anchor):
        ra, pa = curve(anchor)
        rt, pt = curve(test)
        return {"bd_rate_y_pct": bd.bd_rate(ra, pa, rt, pt), "bd_psnr_y_db": bd.bd_psnr(ra, pa, rt, pt),
                "encoder_time_ratio": float(np.mean([x["seconds"] for x in rep["runs"][anchor]]) / max(1e-9, np.mean([x["seconds"] for x in rep["runs"][test]])))}
    rep["bd"] = {}
    for t in rep["runs"]:
        for an in ("anchor", "hm_dl"):
            if t != an and an in rep["runs"] and not (t == "anchor"):
                rep["bd"]["%s_vs_%s" % (t, an)] = cmp(t, an)
    save(rep, a.out)
    print(json.dumps(rep["bd"], indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True, choices=["cpu", "labels", "files", "dropin", "report"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--qps", default="22,27,32,37")
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--work", default="/tmp/bd100")
    ap.add_argument("--labels-npz", default=os.path.join(ROOT, "gpurun_out", "bd100_labels.npz"))
    ap.add_argument("--labels-key", default="", help="stage cpu: take the reference's labels from this key of --labels-npz instead of running the CPU oracle")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_bdrate_1080p_100f.json"))
    a = ap.parse_args()
    a.qp_list = [int(q) for q in a.qps.split(",")]
    rep = load(a.out)
    {"cpu": stage_cpu, "labels": stage_labels, "files": stage_files, "dropin": stage_dropin, "report": stage_report}[a.stage](a, rep)


if __name__ == "__main__":
    main()
