"""Run one configuration of a hot kernel a few times (to be wrapped in ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom import functional as F  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "nn"
B, N = int(sys.argv[2]) if len(sys.argv) > 2 else 256, int(sys.argv[3]) if len(sys.argv) > 3 else 4096
T, RB = int(sys.argv[4]) if len(sys.argv) > 4 else 0, int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
torch.manual_seed(0)
x = torch.randn(B, N, 3, device="cuda")
x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
y = x + 0.01 * torch.randn_like(x)
F.tune_nn_bidir(T, RB)
for _ in range(iters):
    if which == "nn":
        F.nn_bidir(x, y)
    else:
        F.knn_self(y, 6)
torch.cuda.synchronize()
