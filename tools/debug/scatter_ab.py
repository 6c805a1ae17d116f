"""A/B: gather/group gradient, staged single-buffer kernel vs bulk-copy (UBLKCP) pipeline, and the CSR build alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from hitgeom import _lib
from hitgeom.pointnet2_ops import _ext
from bench_ops import timeit

flush = (torch.empty(256 << 20, dtype=torch.uint8, device="cuda"), torch.zeros(64 << 20, dtype=torch.float32, device="cuda"))
B = 64
for (C, S, ns, n) in ((1, 512, 32, 1024), (3, 512, 32, 1024), (64, 512, 32, 1024), (131, 128, 64, 512), (256, 512, 32, 1024)):
    idx = torch.randint(0, n, (B, S, ns), device="cuda", dtype=torch.int32)
    go = torch.randn(B, C, S, ns, device="cuda")
    by = 4 * B * (C * n + S * ns + C * S * ns)
    out = []
    for mode in (1, 3, 2):
        _lib.lib().hg_tune(b"scatter", mode)
        t = timeit(lambda: _ext.group_points_grad(go, idx, n), flush=flush)
        out.append(t)
    _lib.lib().hg_tune(b"scatter", 0)
    print(f"group_points_grad C={C} n={n} S={S} ns={ns}: staged {out[0]*1e3:7.1f} us  bulk {out[1]*1e3:7.1f} us  bulk+list in smem {out[2]*1e3:7.1f} us "
          f"(alg {by/1e6:.0f} MB -> {by/out[2]/1e6:.0f} GB/s)", flush=True)
