"""Time the Chamfer distance pass (profiling hook HG_PROF_NN_BIDIR) with and without the seeded filter kernel."""
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
    ms, n = _lib.prof_read("nn_bidir")
    _lib.prof_enable(False)
    return ms / max(n, 1)


torch.manual_seed(0)
for (B, N) in [(128, 16384), (388, 1024), (4096, 1024), (512, 4096)]:
    x = torch.randn(B, N, 3, device="cuda")
    x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
    y = x + torch.clamp(0.01 * torch.randn_like(x), -0.05, 0.05)
    z = torch.randn(B, N, 3, device="cuda")
    z = z / z.norm(dim=-1).amax(dim=1)[:, None, None]
    for name, other in (("jittered copy", y), ("unrelated cloud", z)):
        for mode in (0, -1):
            F.set_nn_filter(mode)
            ms = timed(lambda: F.nn_bidir(x, other))
            print(f"nn_bidir B={B} N={N} {name:16s} filter={'auto' if mode == 0 else 'off '}: {ms:.3f} ms  "
                  f"{float(B) * N * N / ms * 1e3:.3e} pair-evals/s", flush=True)
F.set_nn_filter(0)
