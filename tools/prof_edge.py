"""One forward+backward of the DGCNN edge-feature op at config-4 layer sizes (to be wrapped in ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom import functional as F  # noqa: E402
from hitgeom import model_seams as ms  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
x = torch.randn(32, C, 1024, device="cuda")
idx = ms.knn(x, 20)
go = torch.randn(32, 2 * C, 1024, 20, device="cuda")
xg = x.clone().requires_grad_()
for _ in range(3):
    xg.grad = None
    F.edge_feature(xg, idx).backward(go)
torch.cuda.synchronize()
