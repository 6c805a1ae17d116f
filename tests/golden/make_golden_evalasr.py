"""tests/golden/evalasr_ref.npz: the UNMODIFIED reference evaluation loop `eval_ASR` (util/other_utils.py:15-101) run on
this container's CPU over two small seeded batches, with an attack object that returns FIXED adversarial clouds (so
the fixture pins the metric block -- KNNDist(k=4), uniform_loss, CurvStdDist(k=4), the ASR counters -- not an attack
trajectory).  Native dependencies absent here are answered by stand-ins backed by the pinned oracle, as in
make_golden_metrics.py.  Build container only."""
import logging
import os
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _refload  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util_inputs import clouds, jitter  # noqa: E402
from util_models import TinyPointNet  # noqa: E402


def _np(t):
    return np.ascontiguousarray(t.detach().numpy())


class FixedAttack:
    """Stands in for val_attack: returns pre-computed adversarial clouds, batch by batch (numpy, as HiT_ADV does)."""

    def __init__(self, advs):
        self.advs, self.i = advs, 0

    def attack(self, data, label):
        out = self.advs[self.i]
        self.i += 1
        return out, None


def main():
    _refload.install_stubs()
    KNN = namedtuple("KNN", "dists idx knn")

    def knn_points(p1, p2, K=1, **kw):
        d, i = O.knn_points(_np(p1), _np(p2), K)
        return KNN(torch.from_numpy(d), torch.from_numpy(i), None)

    def knn_gather(x, idx):
        B, L, K = idx.shape
        return torch.gather(x, 1, idx.reshape(B, L * K, 1).expand(-1, -1, x.shape[-1])).view(B, L, K, x.shape[-1])

    sys.modules["pytorch3d.ops"].knn_points = knn_points
    sys.modules["pytorch3d.ops"].knn_gather = knn_gather
    pu = types.ModuleType("pointnet2_ops_lib.pointnet2_ops.pointnet2_utils")
    pu.furthest_point_sample = lambda xyz, m: torch.from_numpy(O.p2_fps(_np(xyz), m))
    pu.gather_operation = lambda feats, idx: torch.from_numpy(O.p2_gather(_np(feats), _np(idx)))
    pu.ball_query = lambda r, ns, xyz, new_xyz: torch.from_numpy(O.p2_ball_query(_np(new_xyz), _np(xyz), np.float32(r), ns))
    pu.grouping_operation = lambda feats, idx: torch.from_numpy(O.p2_group(_np(feats), _np(idx)))
    lib = types.ModuleType("pointnet2_ops_lib")
    ops = types.ModuleType("pointnet2_ops_lib.pointnet2_ops")
    lib.pointnet2_ops, ops.pointnet2_utils = ops, pu
    sys.modules.update({"pointnet2_ops_lib": lib, "pointnet2_ops_lib.pointnet2_ops": ops,
                        "pointnet2_ops_lib.pointnet2_ops.pointnet2_utils": pu})
    sys.modules.setdefault("scipy.io", __import__("scipy.io"))
    import importlib

    other = importlib.import_module("util.other_utils")

    B, K, nb = 3, 512, 2
    model = TinyPointNet(40, seed=3)
    batches, advs = [], []
    for i in range(nb):
        pts = clouds(B, K, 700 + i, "surface")
        rng = np.random.default_rng(40 + i)
        nrm = rng.standard_normal((B, K, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        with torch.no_grad():
            label = model(torch.from_numpy(pts).transpose(1, 2)).argmax(1)
        label[0] = (label[0] + 1) % 40  # one cloud the victim gets wrong: excluded from the ASR denominator
        batches.append((torch.from_numpy(np.concatenate([pts, nrm], -1)), label))
        adv = jitter(pts, 90 + i, sigma=0.03, clip=0.1)
        adv[1] = 3.0 * pts[1]  # one cloud blown up threefold: the (random-init) victim's prediction flips
        advs.append(adv.astype(np.float64))
    args = types.SimpleNamespace(ker_weight=1.0, hide_weight=1.0, budget=0.55, max_sigm=1.2, min_sigm=0.1, central_num=192,
                                 attack_type="HiT_ADV", k=5, model="tiny")
    records = []

    class Grab(logging.Handler):
        def emit(self, record):
            records.append(record.getMessage())

    h = Grab()
    logging.getLogger().addHandler(h)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "log"))
        os.chdir(tmp)
        try:
            asr = other.eval_ASR(model, batches, args, FixedAttack(advs))
        finally:
            os.chdir(cwd)
            for hd in list(logging.getLogger().handlers):
                logging.getLogger().removeHandler(hd)
    vals = {}
    for m in records:
        for key, tag in (("knn", "Overall KNN dist: "), ("uniform", "Overall Uniform dist: "), ("curvstd", "Overall CurvStd dist: "),
                         ("asr_logged", "Overall attack success rate: ")):
            if m.startswith(tag):
                vals[key] = float(m[len(tag):].replace("tensor(", "").replace(")", "").split(",")[0])
    out = dict(asr=float(asr), **vals, model_seed=3, k=5)
    for i, ((d, l), a) in enumerate(zip(batches, advs)):
        out[f"data{i}"], out[f"label{i}"], out[f"adv{i}"] = d.numpy(), l.numpy(), a
    np.savez_compressed(os.path.join(HERE, "evalasr_ref.npz"), **out)
    print({k: v for k, v in out.items() if not hasattr(v, "shape") or v.ndim == 0})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
