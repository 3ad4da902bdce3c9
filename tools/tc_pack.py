#!/usr/bin/env python
"""Pack the HDLW fp32 weights into the tensor-core operand blob (HDLT) that libhevcdl.so's tcgen05
CNN path loads: every weight matrix is stored as bf16 in the exact shared-memory byte layout the
UMMA descriptors address (K-major, no swizzle: 8-row x 16-byte core matrices, LBO = 128 B between
the two K halves of a K=16 block, SBO between 8-row groups), so the kernels bulk-copy it into place.

Conventions shared with csrc/cnn_tc.cuh (see DESIGN.md "tensor-core formulation"):
  * conv biases are dropped: train-mode BatchNorm subtracts the per-sample mean, so a per-channel
    constant cancels exactly (use_model.py:16-47 with BN in training mode);
  * the sign of each BN gamma is folded into the conv weights so the kernels pool with max only
    (bn is monotone increasing for gamma > 0): stored gamma is |gamma|;
  * layer 1 consumes integer pixels 0..255 (exact in bf16); the 1/255 of ToTensor is folded into
    the BN epsilon (eps * 255^2) instead of the weights.

Usage: python tools/tc_pack.py weights/hevc_encoder_model.hdlw weights/hevc_encoder_model.hdlt
"""
import sys

import numpy as np

NAMES = [("c1w", (16, 3, 5, 5)), ("c1b", (16,)), ("g1", (16,)), ("b1", (16,)),
         ("c64w", (16, 3, 5, 5)), ("c64b", (16,)), ("g64", (16,)), ("b64", (16,)),
         ("c2w", (64, 32, 3, 3)), ("c2b", (64,)), ("g2", (64,)), ("b2", (64,)),
         ("c3w", (128, 64, 3, 3)), ("c3b", (128,)), ("g3", (128,)), ("b3", (128,)),
         ("f1w", (256, 2048)), ("f1b", (256,)), ("f2w", (64, 256)), ("f2b", (64,)), ("f3w", (16, 64)), ("f3b", (16,))]

# section sizes in bytes (fixed; mirrored in csrc/cnn_tc.cuh)
SZ_L1W = 24 * 4096
SZ_W2 = 18 * 2048
SZ_W3 = 36 * 4096
SZ_FC1 = 32 * 32768
SZ_FC2 = 128 * 256 * 2
N_F32 = 32 + 32 + 128 + 256 + 256 + 64 + 1024 + 16
SZ_TOTAL = SZ_L1W + SZ_W2 + SZ_W3 + SZ_FC1 + SZ_FC2 + 4 * N_F32


def load_hdlw(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"HDLW0001"
    w = np.frombuffer(raw, "<f4", offset=8)
    out, o = {}, 0
    for name, shape in NAMES:
        n = int(np.prod(shape))
        out[name] = w[o:o + n].reshape(shape).astype(np.float32)
        o += n
    assert o == w.size
    return out


def to_bf16_bits(x):
    """float32 -> bf16 bit pattern (uint16), round to nearest even."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def canonical(mat, sbo_bytes=None):
    """[R][K] (K multiple of 16 handled by the caller: here K == 16) -> uint16 array in the
    no-swizzle K-major core-matrix layout with LBO = 128 B, SBO = 256 B."""
    R, K = mat.shape
    assert K == 16 and R % 8 == 0
    out = np.zeros(R * 16, np.uint16)
    bits = to_bf16_bits(mat)
    r = np.arange(R)[:, None]
    k = np.arange(16)[None, :]
    byte = (r // 8) * 256 + (k // 8) * 128 + (r % 8) * 16 + (k % 8) * 2
    out[(byte // 2).ravel()] = bits.ravel()
    return out


def pack_l1(w, gamma):
    """Layer-1 B operands: 12 matrices [128 n][16 k], kb = wr*2 + plane.
    n = (c//8)*64 + (dy*4 + dx)*8 + c%8  (output pixel (2g+dy, 4i+dx) of the 4x2 super-pixel, channel c: the 64
    accumulator columns an epilogue warp reads -- 8 positions x its 8 channels -- are contiguous in tensor memory);
    k = j*2 + ch2 (input pixel column j = 0..7 of the 8-px window, ch2 of the plane):
    plane 0 holds (R,G), plane 1 holds (B,0).  Window row wr = 0..5 <-> input row 2g + wr - 2."""
    sgn = np.where(gamma < 0, -1.0, 1.0).astype(np.float32)
    mats = []
    for wr in range(6):
        for p in range(2):
            m = np.zeros((128, 16), np.float32)
            for dy in range(2):
                ky = wr - dy
                if not 0 <= ky <= 4:
                    continue
                for dx in range(4):
                    for j in range(8):
                        kx = j - dx
                        if not 0 <= kx <= 4:
                            continue
                        for ch2 in range(2):
                            ci = p * 2 + ch2
                            if ci > 2:
                                continue
                            pos = dy * 4 + dx
                            for hh in range(2):
                                n0 = hh * 64 + pos * 8
                                m[n0:n0 + 8, j * 2 + ch2] = (w[:, ci, ky, kx] * sgn)[8 * hh:8 * hh + 8]
            mats.append(canonical(m))
    return np.concatenate(mats)


def pack_w2(w, gamma):
    """conv2 B operands: 18 matrices [64 n][16 k], kb = tap*2 + j; k <-> input channel 16*j + k."""
    sgn = np.where(gamma < 0, -1.0, 1.0).astype(np.float32)
    mats = []
    for tap in range(9):
        for j in range(2):
            m = w[:, 16 * j:16 * j + 16, tap // 3, tap % 3] * sgn[:, None]
            mats.append(canonical(np.ascontiguousarray(m)))
    return np.concatenate(mats)


def pack_w3(w, gamma):
    """conv3 A operands: 36 matrices [128 m][16 k], kb = j*9 + tap; k <-> input channel 16*j + k."""
    sgn = np.where(gamma < 0, -1.0, 1.0).astype(np.float32)
    mats = []
    for j in range(4):
        for tap in range(9):
            m = w[:, 16 * j:16 * j + 16, tap // 3, tap % 3] * sgn[:, None]
            mats.append(canonical(np.ascontiguousarray(m)))
    return np.concatenate(mats)


def pack_fc1(w):
    """fc1 A operand [256 m][2048 k] tiled [kc = k/64][m/8][(k/8)%8][m%8][k%8]: one 32 KB block per kc,
    inside it LBO = 128 B (next 8 k), SBO = 1024 B (next 8 rows)."""
    bits = to_bf16_bits(w)
    out = np.zeros(256 * 2048, np.uint16)
    m = np.arange(256)[:, None]
    k = np.arange(2048)[None, :]
    idx = ((k // 64) * 32 + m // 8) * 512 + ((k // 8) % 8) * 64 + (m % 8) * 8 + k % 8
    out[idx.ravel()] = bits.ravel()
    return out


def pack_fc2(w):
    """fc2 A operand [128 m (64 real, rest zero)][256 k]: elem (m,k) at (m/8)*4096 + (k/8)*128 + (m%8)*16 + (k%8)*2 bytes."""
    full = np.zeros((128, 256), np.float32)
    full[:64] = w
    bits = to_bf16_bits(full)
    out = np.zeros(128 * 256, np.uint16)
    m = np.arange(128)[:, None]
    k = np.arange(256)[None, :]
    byte = (m // 8) * 4096 + (k // 8) * 128 + (m % 8) * 16 + (k % 8) * 2
    out[(byte // 2).ravel()] = bits.ravel()
    return out


def pack(w):
    f32 = np.concatenate([np.abs(w["g64"]), w["b64"], np.abs(w["g1"]), w["b1"], np.abs(w["g2"]), w["b2"],
                          np.abs(w["g3"]), w["b3"], w["f1b"], w["f2b"], w["f3w"].ravel(), w["f3b"]]).astype("<f4")
    assert f32.size == N_F32
    parts = [np.concatenate([pack_l1(w["c64w"], w["g64"]), pack_l1(w["c1w"], w["g1"])]), pack_w2(w["c2w"], w["g2"]),
             pack_w3(w["c3w"], w["g3"]), pack_fc1(w["f1w"]), pack_fc2(w["f2w"])]
    blob = b"".join(p.astype("<u2").tobytes() for p in parts) + f32.tobytes()
    assert len(blob) == SZ_TOTAL, (len(blob), SZ_TOTAL)
    return blob


def main(src, dst):
    blob = pack(load_hdlw(src))
    with open(dst, "wb") as f:
        f.write(b"HDLT0001")
        f.write(blob)
    print("wrote", dst, 8 + len(blob), "bytes")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
