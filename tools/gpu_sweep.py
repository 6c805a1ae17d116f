"""Tile-shape / size sweep of the hot kernels on one GPU (development tool; results land in gpurun_out/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hitgeom import _lib  # noqa: E402
from hitgeom import functional as F  # noqa: E402


def timed(tag, fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    ms, n = _lib.prof_read(tag)
    _lib.prof_enable(False)
    return ms / max(n, 1)


def main():
    res = []
    info = _lib.device_info()
    print(info)
    for (B, N) in [(388, 1024), (4096, 1024), (64, 16384), (512, 4096)]:
        x = torch.randn(B, N, 3, device="cuda")
        x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
        y = x + 0.01 * torch.randn_like(x)
        pairs = float(B) * N * N
        for T in (8, 16):
            for RB in (32, 64, 128, 256, 512):
                F.tune_nn_bidir(T, RB)
                ms = timed("nn_bidir", lambda: F.nn_bidir(x, y))
                r = {"kernel": "nn_bidir", "B": B, "N": N, "T": T, "RB": RB, "ms": ms, "pairs_per_s": pairs / ms * 1e3}
                print(r, flush=True)
                res.append(r)
        F.tune_nn_bidir(0, 0)
        for k1 in (6, 20):
            ms = timed("knn", lambda: F.knn_self(y, k1))
            r = {"kernel": "knn3", "B": B, "N": N, "k1": k1, "ms": ms, "pairs_per_s": pairs / ms * 1e3}
            print(r, flush=True)
            res.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"device": info, "results": res}, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
