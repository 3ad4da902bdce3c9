"""The deblocking oracle (oracle/dbf_oracle.c) against reconstructed pictures dumped by the reference encoder itself right
before and right after its own TComLoopFilter::loopFilterPic (tools/gen_golden_tq.py, oracle/_ref/TAppEncoder_dbftrace)."""
import os

import numpy as np

from conftest import GOLDEN


def cases():
    g = np.load(os.path.join(GOLDEN, "dbf_pictures.npz"))
    for k in range(int(g["ncases"])):
        h = g["hdr_%d" % k]
        W, H = int(h[2]), int(h[3])
        pre = [g["%s0_%d" % (n, k)].reshape(s) for n, s in (("Y", (H, W)), ("U", (H // 2, W // 2)), ("V", (H // 2, W // 2)))]
        post = [g["%s1_%d" % (n, k)].reshape(s) for n, s in (("Y", (H, W)), ("U", (H // 2, W // 2)), ("V", (H // 2, W // 2)))]
        yield k, h, pre, post, g["tu_%d" % k], g["qp_%d" % k]


def test_deblocking_equals_the_references_own_filter(oracle):
    n = 0
    for k, h, pre, post, tu, qp in cases():
        out = oracle.deblock_frame(*pre, tu, qp, int(h[4]), int(h[5]), int(h[6]), int(h[7]))
        for a, b, name in zip(out, post, "YUV"):
            assert (a == b).all(), (k, name, int((a != b).sum()))
        assert (pre[0] != post[0]).sum() > 1000 and (pre[1] != post[1]).sum() > 100      # the filter did something
        n += 1
    assert n == 3
