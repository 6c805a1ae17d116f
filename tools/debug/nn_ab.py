"""A/B: hg_nn_bidir_f32 with the approximate tracker (default) vs the exact tracker, whole call and tagged main kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from hitgeom import _lib
from hitgeom import functional as F
from util_inputs import clouds, jitter

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (B, N, kind) in ((388, 1024, "gauss"), (388, 1024, "surface"), (256, 4096, "gauss"), (128, 16384, "gauss"), (1024, 16384, "gauss")):
    if B * N > 4e6:
        x = torch.randn(B, N, 3, device="cuda")
        x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
        y = x + 0.01 * torch.randn_like(x).clamp(-5, 5)
    else:
        xn = clouds(B, N, 1234, kind)
        x, y = torch.from_numpy(xn).cuda(), torch.from_numpy(jitter(xn, 1)).cuda()
    res = {}
    for mode in (1, 0):
        _lib.lib().hg_tune(b"nn_exact", mode)
        for _ in range(3):
            out = F.nn_bidir(x, y)
        torch.cuda.synchronize()
        _lib.prof_enable(True)
        evs = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = F.nn_bidir(x, y); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        kms, n = _lib.prof_read("nn_bidir")
        _lib.prof_enable(False)
        res[mode] = (sorted(a.elapsed_time(b) for a, b in evs)[2], kms / n, out)
    same = all(torch.equal(p, q) for p, q in zip(res[0][2], res[1][2]))
    pairs = float(B) * N * N
    for mode, tag in ((1, "exact  "), (0, "approx ")):
        w, k, _ = res[mode]
        print(f"B={B} N={N} {kind}: {tag} whole {w*1e3:9.1f} us  main kernel {k*1e3:9.1f} us  "
              f"{pairs/k*1e3:.3e} pair-evals/s = {pairs*8/k*1e3/74.45e12*100:.1f}% of FP32 peak   same bits: {same}", flush=True)
_lib.lib().hg_tune(b"nn_exact", 1)
