"""Time the 3-D kNN kernel (profiling hook HG_PROF_KNN) over a few shapes (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom import _lib  # noqa: E402
from hitgeom import functional as F  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    ms, n = _lib.prof_read("knn")
    _lib.prof_enable(False)
    return ms / max(n, 1)


torch.manual_seed(0)
for (B, N) in [(128, 16384), (388, 1024), (4096, 1024), (512, 4096)]:
    x = torch.randn(B, N, 3, device="cuda")
    x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
    for k1 in (6, 20, 32):
        ms = timed(lambda: F.knn_self(x, k1))
        print(f"knn3 B={B} N={N} k1={k1}: {ms:.3f} ms  {float(B) * N * N / ms * 1e3:.3e} pair-evals/s", flush=True)
