"""Whole self-kNN call (seed pre-pass + main kernel) with and without grid-seeded thresholds, per cloud size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom import _lib  # noqa: E402
from hitgeom import functional as F  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


torch.manual_seed(0)
for N in (512, 1024, 2048, 4096, 8192, 16384):
    B = max(8, (1 << 22) // N)
    x = torch.randn(B, N, 3, device="cuda")
    x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
    for k1 in (6, 20):
        res = {}
        for name, minn, near8 in (("seeded27", 1, 2), ("seeded8", 1, 1), ("unseeded", 1 << 30, 0)):
            _lib.lib().hg_knn_tune(minn, near8)
            res[name] = timed(lambda: F.knn_self(x, k1))
        print(f"B={B} N={N} k1={k1}: seeded(27 cells) {res['seeded27']:.3f} ms  seeded(8 cells) {res['seeded8']:.3f} ms  "
              f"unseeded {res['unseeded']:.3f} ms", flush=True)
_lib.lib().hg_knn_tune(0, 0)
