#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (code-size hot spots; no GPU needed).
usage: tools/sass_lines.py <lib.so> <kernel substring> [top N]"""
import re
import subprocess
import sys
import tempfile
from collections import Counter

so, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
d = tempfile.mkdtemp()
import os
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
import glob
cub = glob.glob(d + "/*.cubin")[0]
out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur, fn, cnt, inl = None, None, Counter(), Counter()
for l in out.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    if fn is None or kern not in fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l) and cur:
        cnt[cur] += 1
print("total SASS instructions:", sum(cnt.values()), "=", sum(cnt.values()) * 16, "bytes")
for (f, ln), c in cnt.most_common(top):
    print("%s:%d  %d" % (f, ln, c))
