import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch
from hitgeom import functional as F
B, N = 128, 16384
torch.manual_seed(0)
x = torch.randn(B, N, 3, device="cuda"); x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
y = x + torch.clamp(0.01 * torch.randn_like(x), -0.05, 0.05)
for _ in range(3):
    F.nn_bidir(x, y)
torch.cuda.synchronize()
