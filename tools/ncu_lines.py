#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on.
usage: tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N] [launch-skip]
Aggregates the SASS rows of `ncu --page source --print-source cuda,sass --csv` onto the CUDA line
they belong to: warp instructions executed, stall samples and the dominant stall reasons."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    fname, hdr = None, None
    agg = defaultdict(lambda: {"inst": 0, "samp": 0, "stall": defaultdict(int), "src": "", "shexc": 0})
    cur = None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Name":
            fname = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr is None or len(row) < 3:
            continue
        if row[0]:                      # a CUDA source line
            cur = (fname, int(row[0]))
            agg[cur]["src"] = row[1].strip()
        if len(row) > 8 and row[2].startswith("0x") and cur:
            d = dict(zip(hdr[4:], row[4:]))
            a = agg[cur]
            a["inst"] += int(d.get("Instructions Executed") or 0)
            a["samp"] += int(d.get("# Samples") or 0)
            a["shexc"] += int(d.get("L1 Wavefronts Shared Excessive") or 0)
            for k, v in d.items():
                if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
                    a["stall"][k[6:]] += int(v)
    ti = sum(a["inst"] for a in agg.values()) or 1
    ts = sum(a["samp"] for a in agg.values()) or 1
    # the source page lists an instruction under every line of its inline stack, so this total is inclusive (it exceeds
    # smsp__inst_executed.sum); the per-line percentages are relative to it
    print("kernel %s: %d warp instructions (inclusive over inlined lines), %d samples" % (kern, ti, ts))
    tot = defaultdict(int)
    for a in agg.values():
        for k, v in a["stall"].items():
            tot[k] += v
    print("stall samples: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / ts) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
    print("| file:line | inst %% | samples %% | smem excess wavefronts | top stalls | source |\n|---|---|---|---|---|---|")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
        st = ", ".join("%s %d" % kv for kv in sorted(a["stall"].items(), key=lambda kv: -kv[1])[:3])
        print("| %s:%d | %.1f | %.1f | %d | %s | `%s` |" % (f, l, 100 * a["inst"] / ti, 100 * a["samp"] / ts, a["shexc"], st, a["src"][:90]))


if __name__ == "__main__":
    main()
