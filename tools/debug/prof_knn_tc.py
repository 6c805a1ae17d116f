import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch
from hitgeom import model_seams as ms
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
x = torch.randn(32, C, 1024, device="cuda")
for _ in range(3):
    ms.knn(x, 20)
torch.cuda.synchronize()
