"""Seeded synthetic YUV 4:2:0 8-bit content (SURVEY.md 8(d) recipe) used by tests and bench.py.

The reference ships no test content; this generator makes frames whose CNN label histogram is
not degenerate (smooth gradient + noise band + flat rectangle + square-wave texture + discs).
"""
import numpy as np


def _blur(a, sigma=0.8):
    r = 3
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(a, ((r, r), (0, 0)), mode="edge")
    a = sum(k[i] * p[i:i + a.shape[0]] for i in range(2 * r + 1))
    p = np.pad(a, ((0, 0), (r, r)), mode="edge")
    return sum(k[i] * p[:, i:i + a.shape[1]] for i in range(2 * r + 1))


def synth_frame(width, height, frame=0, kind="mixed"):
    """Return (Y[h,w], U[h/2,w/2], V[h/2,w/2]) uint8."""
    rng = np.random.default_rng(1234 + frame)
    h, w = height, width
    if kind == "flat":
        return (np.full((h, w), 128, np.uint8), np.full((h // 2, w // 2), 128, np.uint8),
                np.full((h // 2, w // 2), 128, np.uint8))
    if kind == "noise":
        return (rng.integers(16, 236, (h, w), dtype=np.uint8),
                rng.integers(16, 241, (h // 2, w // 2), dtype=np.uint8),
                rng.integers(16, 241, (h // 2, w // 2), dtype=np.uint8))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    y = 128 + 50 * np.sin(xx / 97.0) * np.cos(yy / 61.0)
    x0, x1 = int(0.31 * w), int(0.52 * w)
    y[:, x0:x1] += rng.normal(0, 25, (h, x1 - x0))
    y[int(0.1 * h):int(0.3 * h), int(0.6 * w):int(0.85 * w)] = 200
    ty0, ty1, tx0, tx1 = int(0.55 * h), int(0.8 * h), int(0.05 * w), int(0.28 * w)
    sq = 40 * np.sign(np.sin(xx[ty0:ty1, tx0:tx1] * np.pi / 4) * np.sin(yy[ty0:ty1, tx0:tx1] * np.pi / 4) + 1e-9)
    y[ty0:ty1, tx0:tx1] = 128 + sq
    for _ in range(60):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        r = rng.integers(8, 81)
        lvl = rng.integers(40, 221)
        ys, ye, xs, xe = max(0, cy - r), min(h, cy + r + 1), max(0, cx - r), min(w, cx + r + 1)
        m = (xx[ys:ye, xs:xe] - cx) ** 2 + (yy[ys:ye, xs:xe] - cy) ** 2 <= r * r
        y[ys:ye, xs:xe][m] = lvl
    y = np.clip(np.rint(_blur(y)), 16, 235).astype(np.uint8)
    cyy, cxx = np.mgrid[0:h // 2, 0:w // 2].astype(np.float64)
    u = np.clip(np.rint(128 + 30 * np.sin(cxx * 2 / 150.0)), 16, 240).astype(np.uint8)
    v = np.clip(np.rint(128 + 30 * np.cos(cyy * 2 / 140.0)), 16, 240).astype(np.uint8)
    return y, u, v


def synth_sequence(width, height, frames, kind="mixed"):
    """[frames] list of (Y,U,V)."""
    return [synth_frame(width, height, f, kind) for f in range(frames)]


def to_i420_bytes(y, u, v):
    return y.tobytes() + u.tobytes() + v.tobytes()
