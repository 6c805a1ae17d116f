"""Host-side pieces around the CW-kNN loop: clip / projection functions and adversarial
losses against vectors produced by the unmodified reference (tests/golden/make_golden_cwknn.py)."""
import numpy as np
import torch

from hitgeom import adv_utils, clip_utils


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_clip_functions_match_reference_bitwise(golden):
    g = golden("cwknn_ref")
    pc, ori, normal = _t(g["clip_pc"]), _t(g["clip_ori"]), _t(g["clip_normal"])
    np.testing.assert_array_equal(clip_utils.ClipPointsL2(budget=0.5)(pc.clone(), ori).numpy(), g["clip_l2"])
    np.testing.assert_array_equal(clip_utils.ClipPointsLinf(budget=0.03)(pc.clone(), ori).numpy(), g["clip_linf"])
    np.testing.assert_array_equal(clip_utils.ProjectInnerPoints()(pc.clone(), ori, normal).numpy(), g["clip_proj"])
    np.testing.assert_array_equal(clip_utils.ProjectInnerClipLinf(budget=0.03)(pc.clone(), ori, normal).numpy(),
                                  g["clip_projlinf"])
    # the fixture exercises both special cases: inner points, and perturbations exactly opposite to the normal
    diff = g["clip_pc"] - g["clip_ori"]
    assert ((diff * g["clip_normal"]).sum(1) < 0).any()
    assert np.all(g["clip_proj"][:, :, :8] == g["clip_ori"][:, :, :8])


def test_project_inner_without_normals_is_identity(golden):
    g = golden("cwknn_ref")
    pc, ori = _t(g["clip_pc"]), _t(g["clip_ori"])
    assert clip_utils.ProjectInnerPoints()(pc, ori, None) is pc
    np.testing.assert_array_equal(clip_utils.ProjectInnerClipLinf(budget=0.03)(pc, ori).numpy(), g["clip_linf"])


def test_adv_losses_match_reference_bitwise(golden):
    g = golden("cwknn_ref")
    logits, tgt = _t(g["adv_logits"]), _t(g["adv_targets"])
    assert adv_utils.LogitsAdvLoss(kappa=5.0)(logits, tgt).item() == float(g["adv_logits_loss"])
    assert adv_utils.UntargetedLogitsAdvLoss(kappa=5.0)(logits, tgt).item() == float(g["adv_untargeted_loss"])
    assert adv_utils.CrossEntropyAdvLoss()(logits, tgt).item() == float(g["adv_ce_loss"])
    # column-vector targets are accepted like the reference (adv_utils.py:25-26)
    assert adv_utils.LogitsAdvLoss(kappa=5.0)(logits, tgt.view(-1, 1)).item() == float(g["adv_logits_loss"])
