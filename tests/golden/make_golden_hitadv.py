"""tests/golden/hitadv_ref.npz: the UNMODIFIED reference attack class (ShapeAttack/HiT_ADV.py) run on this container's
CPU on a small seeded problem.  `pytorch3d.ops` (absent third party) is stubbed with the oracle's documented-semantics
knn_points / a plain gather; mayavi/open3d are empty stubs; `.cuda()` is neutralised.  Build container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _refload  # noqa: E402
from collections import namedtuple  # noqa: E402

from oracle import oracle as O  # noqa: E402
from util_inputs import clouds  # noqa: E402
from util_models import TinyPointNet  # noqa: E402

HP = dict(attack_lr=1e-2, init_weight=10.0, max_weight=80.0, binary_step=2, num_iter=6, cd_weight=1e-4, curv_weight=0,
          ker_weight=1.0, hide_weight=1.0, curv_loss_knn=8, central_num=8, total_central_num=16, max_sigm=1.2,
          min_sigm=0.1, budget=0.55, alpha=1, kappa=30.0)


def main():
    _refload.install_stubs()
    KNN = namedtuple("KNN", "dists idx knn")

    def knn_points(p1, p2, K=1, **kw):
        d, i = O.knn_points(p1.detach().numpy(), p2.detach().numpy(), K)
        return KNN(torch.from_numpy(d), torch.from_numpy(i), None)

    def knn_gather(x, idx):
        B, L, K = idx.shape
        return torch.gather(x, 1, idx.reshape(B, L * K, 1).expand(-1, -1, x.shape[-1])).view(B, L, K, x.shape[-1])

    sys.modules["pytorch3d.ops"].knn_points = knn_points
    sys.modules["pytorch3d.ops"].knn_gather = knn_gather
    import importlib

    hit = importlib.import_module("ShapeAttack.HiT_ADV")
    adv_utils = importlib.import_module("util.adv_utils")
    B, K = 4, 256
    pts = clouds(B, K, 2024, "gauss")
    rng = np.random.default_rng(7)
    nrm = rng.standard_normal((B, K, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    data = torch.from_numpy(np.concatenate([pts, nrm], axis=-1))
    model = TinyPointNet(40, seed=3)
    with torch.no_grad():
        target = model(data[:, :, :3].transpose(1, 2)).argmax(1)  # attack the predicted class (so success is possible)
    kw = {k: v for k, v in HP.items() if k != "kappa"}
    attacker = hit.HiT_ADV(model, adv_utils.UntargetedLogitsAdvLoss(kappa=HP["kappa"]), clip_func=None, **kw)
    torch.manual_seed(11)
    best, succ = attacker.attack(data, target)
    out = dict(data=data.numpy(), target=target.numpy(), best=best, success=int(succ), seed=11, model_seed=3)
    out.update({"hp_" + k: v for k, v in HP.items()})
    np.savez_compressed(os.path.join(HERE, "hitadv_ref.npz"), **out)
    print("success", int(succ), "moved", float(np.abs(best - pts).max()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
