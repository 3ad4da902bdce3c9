"""Host-side tools that read GPU logs (no GPU needed)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_timeline_reader_groups_launches(tmp_path):
    """tools/timeline.py: per-CTA (kernel, block, entry, waited, exit) records of a -DHEVCDL_TIMELINE build are grouped
    into launches and printed per K1 window."""
    rec = []
    t = 1_000_000
    for batch in range(3):                       # three batches: K1, K2, K6 plan + items, K3, K4 back to back
        for kid, ctas, dur in ((1, 148, 90_000), (2, 148, 70_000), (5, 255, 20_000), (6, 1184, 240_000), (3, 148, 55_000), (4, 128, 25_000)):
            for b in range(ctas):
                rec.append(((kid << 32) | b, t + b, t + 500 + b, t + dur - (b % 7) * 100))
            t += dur + 2_000
    log = tmp_path / "tl.bin"
    np.array(rec, dtype=np.uint64).tofile(log)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "timeline.py"), str(log), "2"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert sum(l.startswith("window of the K1 launch") for l in lines) == 2
    assert sum("k_rmd_items" in l and " 1184 " in l for l in lines) == 2
    assert sum("k_tc_fc" in l and "  128 " in l for l in lines) == 2
