"""tests/golden/cwknn_ref.npz: the UNMODIFIED reference CW-kNN attackers (CW/kNN.py, CW/UKNN.py) with the reference's
own ChamferkNNDist, adversarial losses and clip functions, run on this container's CPU on a small seeded problem
(`.cuda()` neutralised by _refload), plus direct input/output vectors of util/clip_utils.py and util/adv_utils.py.
Build container only."""
import contextlib
import io
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
from util_inputs import clouds  # noqa: E402
from util_models import TinyPointNet  # noqa: E402

HP = dict(attack_lr=1e-2, num_iter=10, budget=0.03, kappa=15.0, seed=21, model_seed=3)


def main():
    _refload.install_stubs()
    knn_mod = _refload.by_path("ref_cw_knn", "CW/kNN.py")
    uknn_mod = _refload.by_path("ref_cw_uknn", "CW/UKNN.py")
    import importlib

    adv_utils = importlib.import_module("util.adv_utils")
    clip_utils = importlib.import_module("util.clip_utils")
    dist_utils = importlib.import_module("util.dist_utils")

    B, K = 4, 256
    pts = clouds(B, K, 2025, "gauss")
    rng = np.random.default_rng(9)
    nrm = rng.standard_normal((B, K, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    model = TinyPointNet(40, seed=HP["model_seed"])
    with torch.no_grad():
        label = model(torch.from_numpy(pts).transpose(1, 2)).argmax(1)
    out = dict(pts=pts, nrm=nrm, label=label.numpy())
    out.update({"hp_" + k: v for k, v in HP.items()})

    # targeted CW-kNN: drive every cloud to (label + 1) % 40, clip to an l_inf ball
    target = (label + 1) % 40
    atk = knn_mod.CWKNN(model, adv_utils.LogitsAdvLoss(kappa=HP["kappa"]), dist_utils.ChamferkNNDist(),
                        clip_utils.ClipPointsLinf(budget=HP["budget"]), attack_lr=HP["attack_lr"], num_iter=HP["num_iter"])
    torch.manual_seed(HP["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        adv, succ = atk.attack(torch.from_numpy(pts), target)
    out.update(knn_target=target.numpy(), knn_adv=adv, knn_success=int(succ))
    print("CWKNN success", succ, "moved", float(np.abs(adv - pts).max()))

    # untargeted CW-kNN with normals: move away from the label, project inner points + clip
    data6 = torch.from_numpy(np.concatenate([pts, nrm], axis=-1))
    atk = uknn_mod.CWUKNN(model, adv_utils.UntargetedLogitsAdvLoss(kappa=HP["kappa"]), dist_utils.ChamferkNNDist(),
                          clip_utils.ProjectInnerClipLinf(budget=HP["budget"]), attack_lr=HP["attack_lr"],
                          num_iter=HP["num_iter"])
    torch.manual_seed(HP["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        adv, succ = atk.attack(data6, label)
    out.update(uknn_adv=adv, uknn_success=int(succ))
    print("CWUKNN success", succ, "moved", float(np.abs(adv - pts).max()))

    # clip / projection functions and adversarial losses on their own
    g = torch.Generator().manual_seed(5)
    ori = torch.from_numpy(pts).transpose(1, 2).contiguous()
    pc = ori + 0.05 * torch.randn(ori.shape, generator=g)
    normal = torch.from_numpy(nrm).transpose(1, 2).contiguous()
    pc[:, :, :8] = ori[:, :, :8] - 0.02 * normal[:, :, :8]  # exactly opposite to the normal -> zeroed
    out.update(clip_pc=pc.numpy(), clip_ori=ori.numpy(), clip_normal=normal.numpy(),
               clip_l2=clip_utils.ClipPointsL2(budget=0.5)(pc.clone(), ori).numpy(),
               clip_linf=clip_utils.ClipPointsLinf(budget=0.03)(pc.clone(), ori).numpy(),
               clip_proj=clip_utils.ProjectInnerPoints()(pc.clone(), ori, normal).numpy(),
               clip_projlinf=clip_utils.ProjectInnerClipLinf(budget=0.03)(pc.clone(), ori, normal).numpy())
    logits = 3.0 * torch.randn(16, 40, generator=g)
    tgt = torch.randint(0, 40, (16,), generator=g)
    out.update(adv_logits=logits.numpy(), adv_targets=tgt.numpy(),
               adv_logits_loss=adv_utils.LogitsAdvLoss(kappa=5.0)(logits, tgt).numpy(),
               adv_untargeted_loss=adv_utils.UntargetedLogitsAdvLoss(kappa=5.0)(logits, tgt).numpy(),
               adv_ce_loss=adv_utils.CrossEntropyAdvLoss()(logits, tgt).numpy())
    np.savez_compressed(os.path.join(HERE, "cwknn_ref.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
