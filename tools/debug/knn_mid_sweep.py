"""Self-kNN at mid sizes: streaming path (grid seeds + knn3_kernel) vs small-cloud path (Z-order seeds + knn3_small_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from hitgeom import functional as F
from bench import make_clouds
from bench_ops import timeit

flush = (torch.empty(256 << 20, dtype=torch.uint8, device="cuda"), torch.zeros(64 << 20, dtype=torch.float32, device="cuda"))
for B, N, k1 in ((128, 2047, 6), (128, 2048, 6), (96, 3000, 6), (64, 4096, 6), (32, 8192, 6), (64, 4096, 20), (128, 2048, 20)):
    x = torch.from_numpy(make_clouds(B, N, 3)[1]).cuda()
    F.tune_knn_small(-1)
    ref = F.knn_self(x, k1)
    t_stream = timeit(lambda: F.knn_self(x, k1), flush=flush)
    line = f"B={B} N={N} k1={k1}: streaming {t_stream * 1e3:8.1f} us"
    F.tune_knn_small(8192)
    for gp in (0, 2, 4):
        F.force_knn_shape(0, gp)
        try:
            out = F.knn_self(x, k1)
            same = torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0])
            t = timeit(lambda: F.knn_self(x, k1), flush=flush)
            line += f"   small(gp={gp}) {t * 1e3:8.1f} us same={same}"
        except Exception as e:
            line += f"   small(gp={gp}) {type(e).__name__}"
    F.force_knn_shape(0, 0)
    F.tune_knn_small(0)
    print(line, flush=True)
