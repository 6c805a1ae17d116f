import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import torch
from hitgeom import _lib
from hitgeom import model_seams as ms
torch.manual_seed(0)
for (B, C, N, k) in ((2, 64, 1024, 20), (32, 64, 1024, 20), (32, 32, 300, 7), (32, 128, 1024, 20), (32, 64, 2048, 20), (40, 64, 1024, 21)):
    x = torch.randn(B, C, N, device="cuda")
    _lib.lib().hg_tune(b"knn_tc", 1)
    ref = ms.knn(x, k)
    _lib.lib().hg_tune(b"knn_tc", 2)
    for rep in range(3):
        out = ms.knn(x, k)
        torch.cuda.synchronize()
    print(B, C, N, k, "same:", torch.equal(out, ref), flush=True)
_lib.lib().hg_tune(b"knn_tc", 0)
