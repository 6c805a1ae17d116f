"""A/B: config-5 shard step (Chamfer + kNN, fwd+bwd) serial vs. kNN on a forked stream (eager; 100 ms of kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hit-adv_b200")]
import numpy as np, torch
from hitgeom.dist_utils import ChamferDist, KNNDist
from hitgeom.overlap import side_branch

B, N = int(os.environ.get("B", 512)), 16384
g = torch.Generator(device="cuda").manual_seed(0)
ori = torch.randn(B, N, 3, device="cuda", generator=g)
ori = ori / ori.norm(dim=2, keepdim=True).amax(dim=1, keepdim=True)
adv = (ori + 0.01 * torch.randn(B, N, 3, device="cuda", generator=g)).requires_grad_()
adv.grad = torch.zeros_like(adv)
for temporal in (False, True):
    for mode in ("serial", "overlap"):
        cd, kd = ChamferDist(), KNNDist(k=5).temporal_seeds(temporal)
        def fn():
            adv.grad.zero_()
            if mode == "serial":
                loss = cd(adv, ori) * 5.0 + kd(adv) * 3.0
            else:
                with side_branch() as br:
                    lk = kd(adv)
                loss = cd(adv, ori) * 5.0 + br.join(lk) * 3.0
            loss.backward()
            return loss
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); out = fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        print(mode, "temporal" if temporal else "cold", "B", B, "median ms %.3f" % np.median(ts), "loss", float(out), flush=True)
