"""A/B: config-1 step (CD + HD + kNN, fwd+bwd) serial vs. kNN on a forked stream, both as CUDA graphs, L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hit-adv_b200")]
import numpy as np, torch
from hitgeom.dist_utils import ChamferDist, HausdorffDist, KNNDist, shared_distance_pass
from hitgeom.overlap import side_branch

B, N = 388, 1024
rng = np.random.default_rng(0)
ori = torch.from_numpy(rng.standard_normal((B, N, 3)).astype(np.float32)).cuda()
ori = ori / ori.norm(dim=2, keepdim=True).amax(dim=1, keepdim=True)
adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")

def make(mode, temporal):
    cd, hd, kd = ChamferDist(), HausdorffDist(), KNNDist(k=5).temporal_seeds(temporal)
    def fn():
        adv.grad.zero_()
        with shared_distance_pass():
            if mode == "serial":
                loss = cd(adv, ori) + hd(adv, ori) + kd(adv)
            else:
                with side_branch() as br:
                    lk = kd(adv)
                loss = cd(adv, ori) + hd(adv, ori) + br.join(lk)
        loss.backward()
        return loss
    return fn

adv.grad = torch.zeros_like(adv)
res = {}
for temporal in (False, True):
    for mode in ("serial", "overlap"):
        fn = make(mode, temporal)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            flush.zero_(); flush_rd.sum()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); g.replay(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        res[(mode, temporal)] = (np.median(ts[5:]), float(out), adv.grad.clone())
        print(mode, "temporal" if temporal else "cold", "median ms %.4f" % np.median(ts[5:]), "loss", float(out))
print("same grads serial-vs-overlap:", torch.equal(res[("serial", False)][2], res[("overlap", False)][2]),
      torch.equal(res[("serial", True)][2], res[("overlap", True)][2]))
print("same grads cold-vs-temporal: serial", torch.equal(res[("serial", False)][2], res[("serial", True)][2]),
      "overlap", torch.equal(res[("overlap", False)][2], res[("overlap", True)][2]))
d = (res[("overlap", False)][2] - res[("overlap", True)][2]).abs()
print("overlap cold-vs-temporal max abs diff", float(d.max()), "nonzero", int((d > 0).sum()))

# the two arms alone (graphs, L2 flushed): which one is the critical path
def arm(which, temporal=False):
    cd, hd, kd = ChamferDist(), HausdorffDist(), KNNDist(k=5).temporal_seeds(temporal)
    def fn():
        adv.grad.zero_()
        with shared_distance_pass():
            loss = kd(adv) if which == "knn" else cd(adv, ori) + hd(adv, ori)
        loss.backward()
        return loss
    return fn

for which, temporal in (("knn", False), ("knn", True), ("cd+hd", False)):
    fn = arm(which, temporal)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_(); flush_rd.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print("arm", which, "temporal" if temporal else "cold", "median ms %.4f" % np.median(ts[5:]))
