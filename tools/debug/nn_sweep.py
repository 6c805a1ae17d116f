"""nn_bidir tile-shape sweep at config-1 size (T = columns per lane, RB = rows per CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from hitgeom import _lib, functional as F
from bench import make_clouds
from bench_ops import timeit

flush = (torch.empty(256 << 20, dtype=torch.uint8, device="cuda"), torch.zeros(64 << 20, dtype=torch.float32, device="cuda"))
for B, N in ((388, 1024), (128, 2048), (64, 4096), (256, 16384)):
    ori, adv = make_clouds(B, N, 3)
    ori, adv = torch.from_numpy(ori).cuda(), torch.from_numpy(adv).cuda()
    F.tune_nn_bidir(0, 0)
    ref = F.nn_bidir(ori, adv)
    base = timeit(lambda: F.nn_bidir(ori, adv), flush=flush)
    base2 = timeit(lambda: F.nn_bidir(ori, adv), flush=flush)
    print(f"B={B} N={N}: auto {base * 1e3:.1f} / {base2 * 1e3:.1f} us", flush=True)
    for T in (8, 16):
        for RB in (128, 256, 512, 1024):
            if RB > N:
                continue
            F.tune_nn_bidir(T, RB)
            try:
                out = F.nn_bidir(ori, adv)
                same = all(torch.equal(a, b) for a, b in zip(out, ref))
                t = timeit(lambda: F.nn_bidir(ori, adv), flush=flush)
                print(f"   T={T} RB={RB}: {t * 1e3:.1f} us same={same}", flush=True)
            except Exception as e:
                print(f"   T={T} RB={RB}: {type(e).__name__}", flush=True)
    F.tune_nn_bidir(0, 0)
