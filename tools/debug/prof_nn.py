import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch
from hitgeom import functional as F
from hitgeom import _lib
B, N = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
x = torch.randn(B, N, 3, device="cuda")
x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
y = x + 0.01 * torch.randn_like(x)
for _ in range(2):
    out = F.nn_bidir(x, y)
torch.cuda.synchronize()
ws = _lib.lib().hg_nn_bidir_workspace_bytes(B, N, N, 3)
print("workspace MB", ws / 1e6)
