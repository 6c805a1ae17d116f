"""eval_ASR's per-batch metrics (hitgeom.eval_metrics: uniform_loss, kNN_smoothing_loss, CurvStdDist; SURVEY.md 8a rows
a19, a21) against the output of the UNMODIFIED reference functions run on CPU over oracle-backed stand-ins for their
native dependencies (tests/golden/make_golden_metrics.py).  The index work underneath is bit-exact (tested per op in
test_gpu_pointnet2.py / test_gpu_knn.py); these are float reductions of those results: 1e-5 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cf(a):
    return torch.from_numpy(a).cuda().transpose(1, 2).contiguous()


def test_uniform_loss_matches_reference(golden):
    from hitgeom.eval_metrics import uniform_loss

    g = golden("metrics_ref")
    adv = _cf(g["adv"])
    got = uniform_loss(adv, k=2)
    assert got.shape == g["uniform_k2"].shape and abs(got.item() - float(g["uniform_k2"])) <= 1e-5 * abs(float(g["uniform_k2"]))
    got = uniform_loss(adv.transpose(1, 2).contiguous(), k=4)  # point-major input is accepted too (GeoA3_args.py:259-260)
    assert abs(got.item() - float(g["uniform_k4_point_major"])) <= 1e-5 * abs(float(g["uniform_k4_point_major"]))


def test_knn_smoothing_and_curvstd_match_reference(golden):
    from hitgeom.eval_metrics import CurvStdDist, kNN_smoothing_loss, kappa_std

    g = golden("metrics_ref")
    ori, adv, nrm = _cf(g["ori"]), _cf(g["adv"]), _cf(g["nrm"])
    np.testing.assert_allclose(kNN_smoothing_loss(adv, 5).cpu().numpy(), g["knn_smoothing_k5"], rtol=1e-5)
    np.testing.assert_allclose(kappa_std(adv, nrm, 4).cpu().numpy(), g["kappa_std_k4"], rtol=1e-4, atol=1e-6)
    got = CurvStdDist(k=4)(ori, adv, nrm).item()
    assert abs(got - float(g["curvstd_k4"])) <= 1e-5 * abs(float(g["curvstd_k4"]))


def test_hitadv_curvature_helpers_are_the_same_functions(golden):
    from hitgeom.eval_metrics import kappa_and_neighbours, kappa_std
    from hitgeom.hit_adv import HiT_ADV, UntargetedLogitsAdvLoss
    from util_models import TinyPointNet

    g = golden("metrics_ref")
    adv, nrm = _cf(g["adv"]), _cf(g["nrm"])
    att = HiT_ADV(TinyPointNet(40, seed=0), UntargetedLogitsAdvLoss())
    assert torch.equal(att._get_kappa_ori(adv, nrm, k=2), kappa_and_neighbours(adv, nrm, 2)[0])
    assert torch.equal(att._get_kappa_std_ori(adv, nrm, k=10), kappa_std(adv, nrm, 10))
