"""group_points at the PointNet++ SA-level shape (to be wrapped in ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom.pointnet2_ops import _ext  # noqa: E402

B, C, n, S, ns = 64, 64, 1024, 512, 32
pts = torch.randn(B, C, n, device="cuda")
idx = torch.randint(0, n, (B, S, ns), device="cuda", dtype=torch.int32)
for _ in range(3):
    _ext.group_points(pts, idx)
torch.cuda.synchronize()
