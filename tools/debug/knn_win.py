import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from hitgeom import _lib, functional as F
from util_inputs import clouds
from bench_ops import timeit
flush = (torch.empty(256 << 20, dtype=torch.uint8, device="cuda"), torch.zeros(64 << 20, dtype=torch.float32, device="cuda"))
for kind in ("gauss", "surface"):
    x = torch.from_numpy(clouds(388, 1024, 1234, kind)).cuda()
    for k1 in (6, 20):
        for win in (0, 24, 32, 48, 64, 96):
            _lib.lib().hg_tune(b"knn_win", win)
            t = timeit(lambda: F.knn_self(x, k1), flush=flush)
            print(f"{kind} k1={k1} win={win}: whole {t*1e3:.1f} us", flush=True)
_lib.lib().hg_tune(b"knn_win", 0)
