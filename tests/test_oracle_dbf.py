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


def sao_cases():
    g = np.load(os.path.join(GOLDEN, "sao_stats.npz"))
    for k in range(int(g["ncases"])):
        W, H = (int(v) for v in g["dims_%d" % k])
        shp = ((H, W), (H // 2, W // 2), (H // 2, W // 2))
        org = [g["org%s_%d" % (n, k)].reshape(s) for n, s in zip("YUV", shp)]
        src = [g["src%s_%d" % (n, k)].reshape(s) for n, s in zip("YUV", shp)]
        yield k, org, src, g["stats_%d" % k]


def test_sao_statistics_equal_the_references_own(oracle):
    """oracle/sao_oracle.c against inputs / output of the reference's own TEncSampleAdaptiveOffset::getStatistics."""
    n = 0
    for k, org, src, want in sao_cases():
        got = oracle.sao_stats(org, src)
        assert got.shape == want.shape and (got == want).all(), (k, int((got != want).sum()))
        assert want[:, :, :, 1].sum() > 10000
        n += 1
    assert n == 2


def sao_apply_cases():
    g = np.load(os.path.join(GOLDEN, "sao_apply.npz"))
    for k in range(int(g["ncases"])):
        W, H, qp = (int(v) for v in g["dims_%d" % k])
        shp = ((H, W), (H // 2, W // 2), (H // 2, W // 2))
        src = [g["src%s_%d" % (n, k)].reshape(s) for n, s in zip("YUV", shp)]
        res = [g["res%s_%d" % (n, k)].reshape(s) for n, s in zip("YUV", shp)]
        yield k, src, g["type_%d" % k], g["offset_%d" % k], res


def test_sao_application_equals_the_references_own(oracle):
    """oracle/sao_oracle.c: oracle_sao_apply against deblocked picture / per-CTU parameters / output picture of the reference's own
    TComSampleAdaptiveOffset::offsetCTU over three encodes that between them use all five SAO types."""
    n, types = 0, set()
    for k, src, t, o, res in sao_apply_cases():
        got = oracle.sao_apply(src, t, o)
        for a, b, name in zip(got, res, "YUV"):
            assert (a == b).all(), (k, name, int((a != b).sum()))
        assert sum(int((a != b).sum()) for a, b in zip(src, res)) > 5000
        types |= set(t.ravel().tolist())
        n += 1
    assert n == 3 and types == {-1, 0, 1, 2, 3, 4}
