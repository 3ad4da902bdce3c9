"""Transform-unit coding core on the device (hevcdl_tu_code, csrc/tq.cuh) through the C-ABI: bit-exact against the
reference's own functions / encoder dumps (tests/golden/tq_*.npz, tools/gen_golden_tq.py) and against the oracle on random
TUs of every size, QP and flag."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp(built, host):
    d = host.DepthPredictor(64, 64, precision=host.PREC_FP32, rmd=False, outputs=0)
    yield d
    d.close()


def test_tu_core_vs_reference_encoder_dump(dp, host):
    """Every TU of the fixture (sizes 4..32, luma + chroma QP, DST, transform skip) equals what the reference encoder printed."""
    g = np.load(os.path.join(GOLDEN, "tq_trace_192x128_qp32.npz"))
    n = len(g["sizes"])
    blocks = [g["resi"][g["off"][i]:g["off"][i + 1]].reshape(int(g["sizes"][i]), int(g["sizes"][i])) for i in range(n)]
    out = dp.tu_code(blocks, g["qp"], g["flags"])
    for i in range(n):
        a, b = int(g["off"][i]), int(g["off"][i + 1])
        assert (out["coeff"][i].ravel() == g["coeff"][a:b]).all(), ("coeff", i, g["sizes"][i], g["flags"][i])
        assert (out["level"][i].ravel() == g["level"][a:b]).all(), ("level", i)
        assert out["abs_sum"][i] == np.abs(g["level"][a:b]).sum()
        if g["has_inv"][i]:
            assert (out["deq"][i].ravel() == g["deq"][a:b]).all(), ("deq", i)
            assert (out["rec"][i].ravel() == g["rec"][a:b]).all(), ("rec", i)
            assert out["ssd"][i] == ((blocks[i].astype(np.int64) - g["rec"][a:b].reshape(blocks[i].shape)) ** 2).sum()


def test_forward_transform_vs_the_references_own_function(dp):
    """Transform output on the reference's xTrMxN vectors (random, small, +-255 checkerboards), DCT 4..32 and DST."""
    g = np.load(os.path.join(GOLDEN, "tq_transform_ref.npz"))
    n = len(g["sizes"])
    blocks = [g["resi"][g["off"][i]:g["off"][i + 1]].reshape(int(g["sizes"][i]), int(g["sizes"][i])) for i in range(n)]
    out = dp.tu_code(blocks, np.full(n, 32), g["dst"].astype(np.uint8))
    for i in range(n):
        assert (out["coeff"][i].ravel() == g["coeff"][g["off"][i]:g["off"][i + 1]]).all(), (i, g["sizes"][i], g["dst"][i])


def test_tu_core_vs_oracle_random(dp, oracle, host):
    """2000 random TUs: every size, QP 0..51, DST / transform-skip / inter rounding, residuals up to +-255 and sparse ones;
    coeff, level, deq, rec, abs_sum and ssd all equal the oracle's."""
    rng = np.random.default_rng(11)
    blocks, qps, flags = [], [], []
    for k in range(2000):
        n = int(rng.choice([4, 8, 16, 32], p=[0.4, 0.3, 0.2, 0.1]))
        kind = k % 5
        if kind == 0:
            b = rng.integers(-255, 256, (n, n))
        elif kind == 1:
            b = rng.integers(-6, 7, (n, n))
        elif kind == 2:
            b = np.zeros((n, n), np.int64); b[rng.integers(0, n), rng.integers(0, n)] = rng.integers(-255, 256)
        elif kind == 3:
            b = (rng.integers(0, 2, (n, n)) * 2 - 1) * 255
        else:
            b = np.add.outer(np.arange(n), np.arange(n)) * rng.integers(-7, 8) + rng.integers(-40, 41)
            b = np.clip(b, -255, 255)
        f = 0
        if n == 4:
            f = int(rng.choice([0, host.TU_DST, host.TU_TSKIP]))
        if rng.random() < 0.2:
            f |= host.TU_INTER
        blocks.append(b.astype(np.int16)); qps.append(int(rng.integers(0, 52))); flags.append(f)
    out = dp.tu_code(blocks, qps, flags)
    for i, b in enumerate(blocks):
        c, q, d, r, s = oracle.tq_tu(b, qps[i], flags[i])
        assert (out["coeff"][i] == c).all(), ("coeff", i, b.shape, qps[i], flags[i])
        assert (out["level"][i] == q).all(), ("level", i, b.shape, qps[i], flags[i])
        assert (out["deq"][i] == d).all(), ("deq", i, b.shape, qps[i], flags[i])
        assert (out["rec"][i] == r).all(), ("rec", i, b.shape, qps[i], flags[i])
        assert out["abs_sum"][i] == s and out["ssd"][i] == ((b.astype(np.int64) - r) ** 2).sum()


def test_tu_core_rejects_bad_descriptors(dp, host):
    with pytest.raises(host.HevcdlError):
        dp.tu_code([np.zeros((8, 8), np.int16)], [32], [host.TU_TSKIP])      # transform skip is 4x4 only
    with pytest.raises(host.HevcdlError):
        dp.tu_code([np.zeros((4, 4), np.int16)], [52])                       # QP out of range
    assert dp.tu_code([], [])["abs_sum"].size == 0


def _rdoq_params(host, g, idx):
    rq = np.zeros(len(idx), host.TU_RDOQ_DTYPE)
    for k, i in enumerate(idx):
        h = g["hdr"][i]
        rq[k] = (float(g["lam"][i]), k, 0 if h[2] == 0 else 1, int(h[6]), int(h[8]), int(h[10]) | (int(h[9]) << 1) | (int(h[13]) << 2))
    return rq, np.stack([g["est"][i] for i in idx])


def test_rdoq_vs_reference_encoder_calls(dp, host):
    """The device RDOQ on the inputs of 168 calls of the reference's own xRateDistOptQuant (every size, luma / chroma, the
    three scans, transform skip; real CABAC bit-estimate tables and lambdas): levels and uiAbsSum identical."""
    g = np.load(os.path.join(GOLDEN, "tq_rdoq_192x128_qp32.npz"))
    n = len(g["hdr"])
    idx = list(range(n))
    blocks = [g["src"][g["off"][i]:g["off"][i + 1]].reshape(int(g["hdr"][i][1]), -1).astype(np.int16) for i in idx]
    assert all((b == g["src"][g["off"][i]:g["off"][i + 1]].reshape(b.shape)).all() for i, b in enumerate(blocks))   # coefficients are 16-bit
    rq, est = _rdoq_params(host, g, idx)
    flags = [host.TU_RDOQ | host.TU_COEFF_IN | (host.TU_TSKIP if g["hdr"][i][7] else 0) for i in idx]
    out = dp.tu_code(blocks, [int(g["hdr"][i][3]) for i in idx], flags, rdoq=rq, est=est)
    for i in idx:
        want = g["dst"][g["off"][i]:g["off"][i + 1]]
        assert (out["level"][i].ravel() == want).all(), (i, g["hdr"][i][:8], int((out["level"][i].ravel() != want).sum()))
        assert out["abs_sum"][i] == g["abs_sum"][i]


def test_rdoq_full_path_vs_oracle_random(dp, oracle, host):
    """Residual -> transform -> RDOQ -> dequantiser -> inverse on 600 random TUs with the fixture's real bit-estimate tables
    and lambdas scaled over two decades: coefficients from the device, then levels against oracle.rdoq on those coefficients,
    and dequantised / reconstructed values against the oracle's flat pipeline fed the same levels."""
    g = np.load(os.path.join(GOLDEN, "tq_rdoq_192x128_qp32.npz"))
    rng = np.random.default_rng(21)
    hdr = g["hdr"]
    blocks, qps, flags, rq, ests, meta = [], [], [], [], [], []
    for k in range(600):
        i = int(rng.integers(0, len(hdr)))
        n, ch, scan, ts = int(hdr[i][1]), int(hdr[i][2] != 0), int(hdr[i][6]), int(hdr[i][7])
        amp = int(rng.choice([3, 12, 60, 255]))
        b = rng.integers(-amp, amp + 1, (n, n)).astype(np.int16)
        if k % 4 == 0:
            b = (np.add.outer(np.arange(n), np.arange(n)) * int(rng.integers(-6, 7)) + rng.integers(-3, 4, (n, n))).clip(-255, 255).astype(np.int16)
        qp = int(rng.integers(10, 46))
        lam = float(g["lam"][i]) * float(10 ** rng.uniform(-1, 1))
        f = host.TU_RDOQ | (host.TU_TSKIP if ts else (host.TU_DST if (n == 4 and ch == 0) else 0))
        blocks.append(b); qps.append(qp); flags.append(f)
        rq.append((lam, k, ch, scan, int(hdr[i][8]), 1 | 2)); ests.append(g["est"][i]); meta.append((ch, scan, ts, lam, i))
    out = dp.tu_code(blocks, qps, flags, rdoq=np.array(rq, host.TU_RDOQ_DTYPE), est=np.stack(ests))
    nz = 0
    for k, b in enumerate(blocks):
        ch, scan, ts, lam, i = meta[k]
        c = out["coeff"][k]
        oc, _, _, _, _ = oracle.tq_tu(b, qps[k], flags[k] & 7)
        assert (c == oc).all(), ("coeff", k)
        lev, s = oracle.rdoq(c, ch, scan, qps[k], ts, lam, ests[k], int(hdr[i][8]), 1, 0, 1)
        assert (out["level"][k] == lev).all() and out["abs_sum"][k] == s, ("level", k, b.shape, qps[k], scan, ts)
        nz += s > 0
    assert nz > 200


def test_tu_core_one_8k_picture_of_tus(dp, oracle, host):
    """BASELINE configs[4] scale: the luma area of a 7680x4320 picture as ~690 k TUs (a quarter of the area per size) in ONE
    call.  Size-independent properties on all of them -- abs_sum is the sum of |level|, ssd the squared error of the
    reconstructed residual, an all-zero residual codes to nothing, the coding of a block does not depend on its neighbours
    in the batch -- and equality with the oracle on a sample of 400."""
    rng = np.random.default_rng(3)
    blocks, qps, flags = [], [], []
    for n in (4, 8, 16, 32):
        cnt = 7680 * 4320 // 4 // (n * n)
        lap = np.round(rng.laplace(0, 5, (cnt, n, n))).clip(-255, 255).astype(np.int16)
        lap[::97] = 0
        blocks += list(lap)
        qps += list(rng.integers(12, 45, cnt))
        flags += [host.TU_DST if n == 4 else 0] * cnt
    out = dp.tu_code(blocks, qps, flags, want_coeff=False, want_deq=False)
    n = len(blocks)
    assert n > 600000
    for i in rng.integers(0, n, 400):
        c, q, d, r, s = oracle.tq_tu(blocks[i], int(qps[i]), flags[i])
        assert (out["level"][i] == q).all() and (out["rec"][i] == r).all() and out["abs_sum"][i] == s, i
    lev = np.concatenate([x.ravel() for x in out["level"]]).astype(np.int64)
    rec = np.concatenate([x.ravel() for x in out["rec"]]).astype(np.int64)
    res = np.concatenate([b.ravel() for b in blocks]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum([b.size for b in blocks])])
    assert (np.add.reduceat(np.abs(lev), off[:-1]) == out["abs_sum"]).all()
    assert (np.add.reduceat((res - rec) ** 2, off[:-1]) == out["ssd"]).all()
    zero = np.add.reduceat(np.abs(res), off[:-1]) == 0
    assert zero.sum() > 5000 and (out["abs_sum"][zero] == 0).all()
    sub = rng.integers(0, n, 2000)
    again = dp.tu_code([blocks[i] for i in sub], [qps[i] for i in sub], [flags[i] for i in sub], want_coeff=False, want_deq=False)
    for k, i in enumerate(sub):
        assert (again["level"][k] == out["level"][i]).all() and (again["rec"][k] == out["rec"][i]).all()
