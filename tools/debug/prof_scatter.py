import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch
from hitgeom.pointnet2_ops import _ext
B, S, ns, n = 64, 512, 32, 1024
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
idx = torch.randint(0, n, (B, S, ns), device="cuda", dtype=torch.int32)
go = torch.randn(B, C, S, ns, device="cuda")
for _ in range(3):
    _ext.group_points_grad(go, idx, n)
torch.cuda.synchronize()
