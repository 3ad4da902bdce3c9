#!/usr/bin/env python
"""CPU emulation of the tensor-core CNN kernels' DATA LAYOUT AND ADDRESSING (test infrastructure).

Every tcgen05.mma of csrc/cnn_tc.cuh is replayed here by reading its A/B operands out of byte
buffers through the same shared-memory descriptors (start, LBO, SBO; K-major, no swizzle) the
kernels build, against the same packed weight blob (tools/tc_pack.py) and the same global
intermediate layouts.  Products are bf16 x bf16 accumulated in fp32/fp64 -- what the tensor core
computes up to summation order -- so comparing the emulated logits with the fp32 oracle also
measures the label impact of bf16 operands before any GPU time is spent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tc_pack  # noqa: E402


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x):
    return tc_pack.to_bf16_bits(np.ascontiguousarray(x, np.float32))


def operand(buf_u16, start, lbo, sbo, rows):
    """[rows][16] fp32 matrix addressed by a K-major no-swizzle descriptor over a bf16 byte buffer."""
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    byte = start + (r % 8) * 16 + (r // 8) * sbo + (k // 8) * lbo + (k % 8) * 2
    return bf16_to_f32(buf_u16[byte // 2]).astype(np.float64)


class Blob:
    def __init__(self, blob_bytes):
        b = np.frombuffer(blob_bytes, np.uint8)
        o = 0
        self.l1w = b[o:o + tc_pack.SZ_L1W].view(np.uint16); o += tc_pack.SZ_L1W
        self.w2 = b[o:o + tc_pack.SZ_W2].view(np.uint16); o += tc_pack.SZ_W2
        self.w3 = b[o:o + tc_pack.SZ_W3].view(np.uint16); o += tc_pack.SZ_W3
        self.fc1 = b[o:o + tc_pack.SZ_FC1].view(np.uint16); o += tc_pack.SZ_FC1
        self.fc2 = b[o:o + tc_pack.SZ_FC2].view(np.uint16); o += tc_pack.SZ_FC2
        f = b[o:].view(np.float32)
        i = 0
        def take(n):
            nonlocal i
            v = f[i:i + n]; i += n
            return v
        self.g64, self.b64, self.g1, self.b1 = take(16), take(16), take(16), take(16)
        self.g2, self.b2, self.g3, self.b3 = take(64), take(64), take(128), take(128)
        self.f1b, self.f2b, self.f3w, self.f3b = take(256), take(64), take(1024).reshape(16, 64), take(16)


def bn_relu(pooled_raw, mean, var, gamma_abs, beta, eps):
    y = (pooled_raw - mean) * (gamma_abs / np.sqrt(var + eps)) + beta
    return np.maximum(y, 0.0)


# ---- K1: stage + conv64 + conv1 -> cat ----------------------------------------------------------
P64_PITCH, P64_ROWS = 68, 68            # pixels; 4 bytes per pixel per plane (2 bf16)
P64_BYTES = P64_ROWS * P64_PITCH * 4
P1_PITCH, P1_ROWS = 36, 36
P1_BYTES = P1_ROWS * P1_PITCH * 4
CAT_PLANE = 18 * 18 * 16
CAT_BYTES = 10 * CAT_PLANE
A2_PLANE = 40 * 10 * 16
A2_BYTES = 8 * A2_PLANE


def k1_ctu(rgb, blob):
    """rgb: [3][64][64] u8 -> cat buffer (uint16[CAT_BYTES/2]) in the global layout."""
    px = rgb.astype(np.float32)
    # conv64 planes: [plane][68][68][2]
    p64 = np.zeros((2, P64_ROWS, P64_PITCH, 2), np.float32)
    p64[0, 2:66, 2:66, 0] = px[0]; p64[0, 2:66, 2:66, 1] = px[1]; p64[1, 2:66, 2:66, 0] = px[2]
    b64 = f32_to_bf16(p64).ravel()
    # conv1 planes: [quadrant][plane][36][36][2]
    p1 = np.zeros((4, 2, P1_ROWS, P1_PITCH, 2), np.float32)
    for q in range(4):
        oy, ox = (q // 2) * 32, (q % 2) * 32
        p1[q, 0, 2:34, 2:34, 0] = px[0, oy:oy + 32, ox:ox + 32]
        p1[q, 0, 2:34, 2:34, 1] = px[1, oy:oy + 32, ox:ox + 32]
        p1[q, 1, 2:34, 2:34, 0] = px[2, oy:oy + 32, ox:ox + 32]
    b1 = f32_to_bf16(p1).ravel()

    def tile(buf, plane_bytes, base, pitch_px, row0, col0, wofs):
        D = np.zeros((128, 128))
        for wr in range(6):
            for p in range(2):
                A = operand(buf, base + p * plane_bytes + ((row0 + wr) * pitch_px + col0) * 4, 16, 2 * pitch_px * 4, 128)
                B = operand(blob.l1w, (wofs + wr * 2 + p) * 4096, 128, 256, 128)
                D += A @ B.T
        # D[m = g*8+i][n = h*64 + (dy*4+dx)*8 + c8] -> out[c = 8h+c8][2g+dy][4i+dx]
        return D.reshape(16, 8, 2, 2, 4, 8).transpose(2, 5, 0, 3, 1, 4).reshape(16, 32, 32)

    conv64 = np.zeros((16, 64, 64))
    for qy in range(2):
        for qx in range(2):
            conv64[:, qy * 32:qy * 32 + 32, qx * 32:qx * 32 + 32] = tile(b64, P64_BYTES, 0, P64_PITCH, qy * 32, qx * 32, 0)
    conv1 = np.zeros((4, 16, 32, 32))
    for q in range(4):
        conv1[q] = tile(b1, P1_BYTES, q * 2 * P1_BYTES, P1_PITCH, 0, 0, 12)

    eps1 = 1e-5 * 255.0 * 255.0
    cat = np.zeros((10, 18, 18, 8), np.float32)
    m, v = conv64.mean((1, 2)), conv64.var((1, 2))
    pooled = conv64.reshape(16, 16, 4, 16, 4).max((2, 4))
    a64 = bn_relu(pooled, m[:, None, None], v[:, None, None], blob.g64[:, None, None], blob.b64[:, None, None], eps1)
    for h in range(2):
        cat[8 + h, 1:17, 1:17, :] = a64[8 * h:8 * h + 8].transpose(1, 2, 0)
    for q in range(4):
        m, v = conv1[q].mean((1, 2)), conv1[q].var((1, 2))
        pooled = conv1[q].reshape(16, 16, 2, 16, 2).max((2, 4))
        a1 = bn_relu(pooled, m[:, None, None], v[:, None, None], blob.g1[:, None, None], blob.b1[:, None, None], eps1)
        for h in range(2):
            cat[q * 2 + h, 1:17, 1:17, :] = a1[8 * h:8 * h + 8].transpose(1, 2, 0)
    return f32_to_bf16(cat).ravel(), (conv64, conv1)


# ---- K2: conv2 (cat -> a2) -----------------------------------------------------------------------
def k2_ctu(cat_u16, blob):
    a2 = np.zeros((8, 40, 10, 8), np.float32)
    for s in range(4):
        out = np.zeros((64, 16, 16))
        for xh in range(2):
            D = np.zeros((128, 64))
            for tap in range(9):
                ky, kx = tap // 3, tap % 3
                for j in range(2):
                    plane = s * 2 if j == 0 else 8
                    A = operand(cat_u16, plane * CAT_PLANE + (ky * 18 + kx + 8 * xh) * 16, CAT_PLANE, 288, 128)
                    B = operand(blob.w2, (tap * 2 + j) * 2048, 128, 256, 64)
                    D += A @ B.T
            out[:, :, 8 * xh:8 * xh + 8] = D.reshape(16, 8, 64).transpose(2, 0, 1)
        m, v = out.mean((1, 2)), out.var((1, 2))
        pooled = out.reshape(64, 8, 2, 8, 2).max((2, 4))
        y = bn_relu(pooled, m[:, None, None], v[:, None, None], blob.g2[:, None, None], blob.b2[:, None, None], 1e-5)
        for c8 in range(8):
            # plane c8: rows rho = 4*(py+1)+s, cols px+1
            a2[c8, 4 + s:36 + s:4, 1:9, :] = y[8 * c8:8 * c8 + 8].transpose(1, 2, 0)
    return f32_to_bf16(a2).ravel()


# ---- K3: conv3 (a2 -> features) ------------------------------------------------------------------
def k3_ctu(a2_u16, blob):
    D = np.zeros((128, 256))
    for j in range(4):
        for tap in range(9):
            ky, kx = tap // 3, tap % 3
            A = operand(blob.w3, (j * 9 + tap) * 4096, 128, 256, 128)
            B = operand(a2_u16, (2 * j) * A2_PLANE + (4 * ky * 10 + kx) * 16, A2_PLANE, 160, 256)
            D += A @ B.T
    # D[c][n = 32*y + 8*s + x]
    out = D.reshape(128, 8, 4, 8).transpose(2, 0, 1, 3)          # [s][c][y][x]
    feats = np.zeros((4, 2048), np.float32)
    for s in range(4):
        m, v = out[s].mean((1, 2)), out[s].var((1, 2))
        pooled = out[s].reshape(128, 4, 2, 4, 2).max((2, 4))
        y = bn_relu(pooled, m[:, None, None], v[:, None, None], blob.g3[:, None, None], blob.b3[:, None, None], 1e-5)
        feats[s] = y.reshape(2048)
    return feats


def feats_layout(feats, npad):
    """[n][2048] fp32 -> global tiled bf16 layout [k/64][n/8][(k/8)%8][n%8][k%8]."""
    n_, _ = feats.shape
    out = np.zeros(32 * (npad // 8) * 512, np.uint16)
    n = np.arange(n_)[:, None]
    k = np.arange(2048)[None, :]
    idx = ((k // 64) * (npad // 8) + n // 8) * 512 + ((k // 8) % 8) * 64 + (n % 8) * 8 + k % 8
    out[idx.ravel()] = f32_to_bf16(feats).ravel()
    return out


# ---- K4: fc1 + fc2 + fc3 --------------------------------------------------------------------------
def k4_tile(feats_u16, npad, nt, blob):
    """128 samples [nt*128, nt*128+128) -> logits [128][16] (the kernel walks the same operand layout in 32-sample tiles)."""
    D = np.zeros((256, 128))
    for kc in range(32):
        for t in range(4):
            B = operand(feats_u16, (kc * (npad // 8) + nt * 16) * 1024 + t * 256, 128, 1024, 128)
            for mh in range(2):
                A = operand(blob.fc1, kc * 32768 + mh * 16 * 1024 + t * 256, 128, 1024, 128)
                D[mh * 128:mh * 128 + 128] += A @ B.T
    h1 = np.maximum(D + blob.f1b[:, None], 0.0).astype(np.float32)          # [256][128 samples]
    # fc2 B operand in smem: [n 128][k 256] canonical, LBO 128, SBO 4096
    b2 = np.zeros(128 * 256, np.uint16)
    n = np.arange(128)[:, None]
    k = np.arange(256)[None, :]
    b2[(((n // 8) * 4096 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2) // 2).ravel()] = f32_to_bf16(h1.T).ravel()
    D2 = np.zeros((128, 128))
    for t in range(16):
        A = operand(blob.fc2, t * 256, 128, 4096, 128)
        B = operand(b2, t * 256, 128, 4096, 128)
        D2 += A @ B.T
    h2 = np.maximum(D2[:64] + blob.f2b[:, None], 0.0).astype(np.float32)    # [64][128]
    return (blob.f3w.astype(np.float64) @ h2.astype(np.float64) + blob.f3b[:, None]).T.astype(np.float32)


def frame_logits(Y, U, V, blob, oracle):
    H, W = Y.shape
    cw, ch = (W + 63) // 64, (H + 63) // 64
    nctu = cw * ch
    feats = np.zeros((nctu * 4, 2048), np.float32)
    for a in range(nctu):
        rgb = oracle.stage_ctu_rgb(Y, U, V, a % cw, a // cw)
        cat, _ = k1_ctu(rgb, blob)
        feats[4 * a:4 * a + 4] = k3_ctu(k2_ctu(cat, blob), blob)
    npad = (nctu * 4 + 127) // 128 * 128
    fl = feats_layout(feats, npad)
    logits = np.concatenate([k4_tile(fl, npad, nt, blob) for nt in range(npad // 128)])[:nctu * 4]
    return logits.reshape(nctu, 4, 16)


def main():
    import importlib
    from oracle import oracle
    pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
    host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
    w = tc_pack.load_hdlw(host.DEFAULT_WEIGHTS)
    blob = Blob(tc_pack.pack(w))
    wflat = oracle.load_weights(host.DEFAULT_WEIGHTS)
    tot = flips = 0
    for seed, (W, H) in enumerate([(416, 240), (256, 192), (320, 128)]):
        Y, U, V = pkg.synth.synth_frame(W, H, seed)
        lg = frame_logits(Y, U, V, blob, oracle)
        olab, olg, mar = oracle.frame_labels(wflat, Y, U, V, want_logits=True)
        lab = np.stack([oracle.ctu_labels(l)[0] for l in lg])
        d = np.abs(lg - olg)
        bad = lab != olab
        print("%dx%d: max|dlogit| %.4f mean %.5f; labels differ %d/%d; max oracle margin among differing %.4f" % (
            W, H, d.max(), d.mean(), bad.sum(), bad.size, mar[bad].max() if bad.any() else 0.0))
        tot += bad.size; flips += bad.sum()
    print("total label flips %d/%d" % (flips, tot))


if __name__ == "__main__":
    main()
