"""Seeded synthetic point clouds shared by the tests and the benchmark."""
import numpy as np


def clouds(B, N, seed, kind="gauss"):
    """Unit-ball normalised clouds (mirrors pc_normalize, Dataset/ModelNet.py:12-17)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, N, 3)).astype(np.float32)
    if kind == "surface":  # spheres + a plane, ~1% exact duplicates, a few near-origin points
        h = N // 2
        x[:, :h] /= np.linalg.norm(x[:, :h], axis=-1, keepdims=True)
        x[:, h:, 2] = 0.25
        ndup = max(1, N // 100)
        for b in range(B):
            src = rng.integers(0, N, ndup)
            dst = rng.integers(0, N, ndup)
            x[b, dst] = x[b, src]
    x = x - x.mean(axis=1, keepdims=True)
    x = x / np.linalg.norm(x, axis=-1).max(axis=1)[:, None, None]
    if kind == "surface":
        x[:, : min(3, N)] *= 1e-3
    return np.ascontiguousarray(x.astype(np.float32))


def jitter(x, seed, sigma=0.01, clip=0.05):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray((x + np.clip(sigma * rng.standard_normal(x.shape), -clip, clip)).astype(np.float32))


def normwise(a, b):
    """max_b  ||a_b - b_b||_inf / ||b_b||_inf  (the gradient parity metric)."""
    B = a.shape[0]
    num = np.abs(a - b).reshape(B, -1).max(1)
    den = np.maximum(np.abs(b).reshape(B, -1).max(1), 1e-30)
    return float((num / den).max())
