#!/usr/bin/env python
"""Reads the CTA log of a -DHEVCDL_TIMELINE build (HEVCDL_TIMELINE_OUT=file) and prints, for the last complete launch
batches, when each kernel's CTAs entered, passed their predecessor wait and exited (us, relative to the batch's K1).
usage: tools/timeline.py <log.bin> [batches to show]"""
import sys

import numpy as np

NAMES = {1: "k_tc_l1", 2: "k_tc_conv2", 3: "k_tc_conv3", 4: "k_tc_fc", 5: "k_rmd_plan", 6: "k_rmd_items"}


def main():
    rec = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
    show = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    kid = (rec[:, 0] >> np.uint64(32)).astype(int)
    ent, wai, ext = (rec[:, i].astype(np.int64) for i in (1, 2, 3))
    launches = []                                 # (kernel, first entry, first waited, median waited, last waited, first exit, median exit, last exit, CTAs)
    for k in sorted(NAMES):
        m = kid == k
        if not m.any():
            continue
        order = np.argsort(ent[m])
        e, w, x = ent[m][order], wai[m][order], ext[m][order]
        cuts = np.nonzero(np.diff(e) > 20000)[0] + 1          # launches of one kernel are > 20 us apart
        for a, b in zip(np.r_[0, cuts], np.r_[cuts, len(e)]):
            launches.append((k, e[a:b].min(), w[a:b].min(), np.median(w[a:b]), w[a:b].max(), x[a:b].min(), np.median(x[a:b]), x[a:b].max(), b - a))
    launches.sort(key=lambda r: r[1])
    k1 = [i for i, r in enumerate(launches) if r[0] == 1]
    for n in range(max(0, len(k1) - show - 1), len(k1) - 1):   # skip the very last batch (may be a partial one)
        t0, t1 = launches[k1[n]][1], launches[k1[n + 1]][1]
        print("window of the K1 launch at %.1f us (next K1 enters at +%.1f us):" % ((t0 - launches[k1[0]][1]) / 1e3, (t1 - t0) / 1e3))
        print("  %-12s %5s | entry first | wait passed first/median/last | exit first/median/last | last exit - first wait" % ("kernel", "CTAs"))
        f = lambda v: "%8.1f" % ((v - t0) / 1e3)
        for r in launches:
            if t0 <= r[1] < t1:
                print("  %-12s %5d | %s    | %s %s %s   | %s %s %s | %7.1f" % (NAMES[r[0]], r[8], f(r[1]), f(r[2]), f(r[3]), f(r[4]), f(r[5]), f(r[6]), f(r[7]), (r[7] - r[2]) / 1e3))


if __name__ == "__main__":
    main()
