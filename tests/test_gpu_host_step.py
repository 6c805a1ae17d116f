"""`hg_chamfer_knn_step_host_f32` (host buffers in, loss + gradient out, chunk-pipelined over two streams) against the
device-resident path (hitgeom.dist_utils.ChamferkNNDist + autograd) on the same inputs: same kernels, so the gradient
is bit-identical whatever the chunking; the scalar loss is a mean taken in a different order (1e-6)."""
import numpy as np
import pytest
import torch

from util_inputs import clouds, jitter

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,chunk,method,k", [(10, 512, 4, "adv2ori", 5), (7, 1024, 16, "both", 4), (5, 300, 2, "ori2adv", 3),
                                                  (43, 256, 16, "adv2ori", 5)])
def test_host_step_equals_device_path(B, N, chunk, method, k):
    from hitgeom.dist_utils import ChamferkNNDist
    from hitgeom.host import ChamferKnnHostStep

    ori = clouds(B, N, 50 + B, "surface")
    adv = jitter(ori, 9)
    w = np.linspace(0.5, 2.0, B).astype(np.float32)
    for weights in (None, w):
        a = torch.from_numpy(adv).cuda().requires_grad_()
        dist = ChamferkNNDist(chamfer_method=method, knn_k=k, knn_alpha=1.05, chamfer_weight=5., knn_weight=3.)
        loss = dist(a, torch.from_numpy(ori).cuda(), weights=None if weights is None else torch.from_numpy(weights).cuda())
        loss.backward()
        step = ChamferKnnHostStep(N, chunk_clouds=chunk, chamfer_method=method, knn_k=k)
        grad = torch.empty(B, N, 3).pin_memory()
        got, per_cloud = step(torch.from_numpy(adv).pin_memory(), torch.from_numpy(ori).pin_memory(), grad, weights=weights)
        assert abs(got - loss.item()) <= 1e-6 * abs(loss.item())
        assert per_cloud.shape == (B,) and abs(per_cloud.mean() - got) <= 1e-6 * abs(got)
        assert torch.equal(grad, a.grad.cpu())
        # pageable numpy arrays work too (copies just stop overlapping), and a second call reuses the session
        grad_np = np.empty((B, N, 3), np.float32)
        got2, _ = step(adv, ori, grad_np, weights=weights)
        assert got2 == got and np.array_equal(grad_np, grad.numpy())
        step.close()


def test_host_step_rejects_bad_arguments():
    from hitgeom import HitgeomError
    from hitgeom.host import ChamferKnnHostStep

    step = ChamferKnnHostStep(64, chunk_clouds=2, knn_k=3)
    x = np.zeros((2, 64, 3), np.float32)
    with pytest.raises(HitgeomError):
        step(x.astype(np.float64), x, x.copy())
    with pytest.raises(HitgeomError):
        step(x, x[:, :32], x.copy())
    with pytest.raises(HitgeomError):
        ChamferKnnHostStep(64, knn_k=5, chamfer_method="nearest")
    with pytest.raises(HitgeomError):
        ChamferKnnHostStep(4, knn_k=5)  # k + 1 > N
