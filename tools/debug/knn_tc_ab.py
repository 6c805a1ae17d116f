"""A/B: DGCNN kNN on feature clouds, tensor-core prefilter + exact re-evaluation vs the FP32 tile + row-select path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from hitgeom import _lib
from hitgeom import model_seams as ms
from hitgeom import functional as F
from bench_ops import timeit

flush = (torch.empty(256 << 20, dtype=torch.uint8, device="cuda"), torch.zeros(64 << 20, dtype=torch.float32, device="cuda"))
torch.manual_seed(0)
import sys
SH = [(96, 3, 1024, 20), (128, 3, 1024, 20), (192, 3, 1024, 20), (388, 3, 1024, 20), (128, 3, 1024, 12), (96, 3, 2048, 20)] if 'c3' in sys.argv else None
for (B, C, N, k) in SH or ((32, 64, 1024, 20), (32, 64, 1024, 20), (16, 64, 1024, 20), (4, 64, 1024, 20), (32, 128, 1024, 20), (32, 64, 2048, 20), (8, 64, 1024, 20), (32, 64, 1024, 5), (32, 32, 1024, 20), (32, 3, 1024, 20), (64, 3, 1024, 20), (32, 3, 1024, 12), (32, 3, 2048, 20), (16, 3, 1024, 20)):
    x = torch.randn(B, C, N, device="cuda")
    res = {}
    for off in (1, 2, 5, 0):  # 1: FP32 path; 2: tensor-core path forced; 5: forced, single-sweep threshold; 0: default dispatch
        _lib.lib().hg_tune(b"knn_tc", off)
        res[off] = (timeit(lambda: ms.knn(x, k), flush=flush), ms.knn(x, k))
    _lib.lib().hg_tune(b"knn_tc", 0)
    same = all(torch.equal(res[m][1], res[1][1]) for m in (0, 2, 5))

    def ref_knn():
        inner = -2 * torch.matmul(x.transpose(2, 1), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]

    t_ref = timeit(ref_knn, flush=flush)
    print(f"DGCNN knn B={B} C={C} N={N} k={k}: FP32 path {res[1][0]*1e3:7.1f} us   tensor-core (forced) {res[2][0]*1e3:7.1f} us   single sweep {res[5][0]*1e3:7.1f} us   default {res[0][0]*1e3:7.1f} us   "
          f"torch matmul+topk {t_ref*1e3:7.1f} us   same indices: {same}", flush=True)
