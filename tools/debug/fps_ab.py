"""A/B: furthest point sampling, 128 vs 256 threads per cloud."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hit-adv_b200")]
import torch
from hitgeom._lib import lib
from hitgeom.pointnet2_ops import _ext

for B, N, M in ((64, 1024, 512), (64, 512, 128), (388, 1024, 512), (32, 2048, 512)):
    x = torch.randn(B, N, 3, device="cuda")
    outs = []
    for th in (256, 128):
        lib().hg_tune(b"fps_threads", th)
        for _ in range(3):
            o = _ext.furthest_point_sampling(x, M)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            o = _ext.furthest_point_sampling(x, M)
        e.record()
        torch.cuda.synchronize()
        outs.append(o.clone())
        print(f"B={B} N={N} M={M} threads={th}: {s.elapsed_time(e) / 10 * 1e3:.1f} us")
    assert torch.equal(outs[0], outs[1])
lib().hg_tune(b"fps_threads", 0)
