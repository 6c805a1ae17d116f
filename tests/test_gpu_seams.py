"""Parity of the torch-level seams of the PointNet++ victim (model/pointnet2_utils.py:19-107) against the golden
vectors of the reference and the oracle: all index results and square_distance bit-exact."""
import numpy as np
import pytest
import torch

from util_inputs import clouds, normwise

pytestmark = pytest.mark.gpu


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_seams_golden(golden):
    from hitgeom import model_seams as ms

    g = golden("torch_seams")
    xyz = gpu(g["xyz"])
    torch.manual_seed(7)  # same CPU generator draw as the reference run that produced the fixture
    fps = ms.farthest_point_sample(xyz, 64)
    assert fps.dtype == torch.int64
    assert np.array_equal(fps.cpu().numpy(), g["fps_idx"])
    new_xyz = ms.index_points(xyz, fps)
    assert np.array_equal(new_xyz.cpu().numpy(), g["new_xyz"])
    assert np.array_equal(ms.square_distance(new_xyz, xyz).cpu().numpy(), g["sqdist"])
    for key in g.files:
        if key.startswith("ball_"):
            r, ns = float(key.split("_")[1][1:]), int(key.split("_")[2][2:])
            out = ms.query_ball_point(r, ns, xyz, new_xyz)
            assert out.dtype == torch.int64
            assert np.array_equal(out.cpu().numpy(), g[key]), key
    grouped = ms.index_points(xyz, gpu(g["ball_r0.2_ns32"]))
    assert np.array_equal(grouped.cpu().numpy(), g["index_points_grouped"])


@pytest.mark.parametrize("B,N,S,npoint", [(64, 1024, 512, 512), (3, 512, 128, 128), (2, 100, 7, 7)])
def test_seams_vs_oracle(oracle, B, N, S, npoint):
    from hitgeom import functional as F
    from hitgeom import model_seams as ms

    xyz = clouds(B, N, 900 + N, "surface")
    start = np.random.default_rng(0).integers(0, N, B).astype(np.int64)
    fps = F.fps_torch(gpu(xyz), npoint, gpu(start))
    ofps = oracle.fps_torch(xyz, npoint, start)
    assert np.array_equal(fps.cpu().numpy(), ofps)
    new_xyz = xyz[np.arange(B)[:, None], ofps]
    for r, ns in [(0.2, 32), (0.4, 64)]:
        out = ms.query_ball_point(r, ns, gpu(xyz), gpu(new_xyz))
        assert np.array_equal(out.cpu().numpy(), oracle.query_ball_torch(r, ns, xyz, new_xyz))


@pytest.mark.parametrize("C", [3, 5, 8, 64, 131])
@pytest.mark.parametrize("shape", [(50,), (50, 9)])
def test_index_points_equals_advanced_indexing(C, shape):
    """model/pointnet2_utils.py:43-60: `points[batch_indices, idx, :]`, for the coordinate-row kernel (C = 3), the
    16-byte-chunk kernel (C a multiple of 4) and the generic one."""
    from hitgeom import model_seams as ms

    rng = np.random.default_rng(C)
    pts = torch.randn(4, 300, C, device="cuda")
    idx = gpu(rng.integers(0, 300, (4,) + shape).astype(np.int64))
    out = ms.index_points(pts, idx)
    ref = pts[torch.arange(4, device="cuda").view(4, *([1] * len(shape))).expand_as(idx), idx, :]
    assert out.shape == ref.shape and torch.equal(out, ref)


def test_index_points_backward_is_deterministic_sum():
    from hitgeom import model_seams as ms

    rng = np.random.default_rng(5)
    pts = torch.randn(3, 200, 7, device="cuda", requires_grad=True)
    idx = gpu(rng.integers(0, 200, (3, 50, 9)).astype(np.int64))
    out = ms.index_points(pts, idx)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    ref = torch.zeros(3, 200, 7, dtype=torch.float64, device="cuda")
    for b in range(3):
        ref[b].index_add_(0, idx[b].reshape(-1), w[b].reshape(-1, 7).double())
    assert normwise(pts.grad.cpu().numpy(), ref.float().cpu().numpy()) < 1e-6
    g1 = pts.grad.clone()
    pts.grad = None
    (ms.index_points(pts, idx) * w).sum().backward()
    assert torch.equal(g1, pts.grad)


def test_square_distance_is_differentiable_like_the_reference():
    """model/pointnet2_utils.py:19-40 is plain torch, so PointNetFeaturePropagation differentiates through it
    (:297-303); the kernel-backed version must hand back the same gradients (torch's expression as the yardstick)."""
    import torch

    from hitgeom.model_seams import square_distance

    g = torch.Generator().manual_seed(0)
    src = torch.randn(2, 70, 3, generator=g).cuda().requires_grad_()
    dst = torch.randn(2, 45, 3, generator=g).cuda().requires_grad_()
    w = torch.randn(2, 70, 45, generator=g).cuda()
    (square_distance(src, dst) * w).sum().backward()
    s2, d2 = src.detach().double().requires_grad_(), dst.detach().double().requires_grad_()
    ref = -2 * torch.matmul(s2, d2.permute(0, 2, 1)) + (s2 ** 2).sum(-1)[:, :, None] + (d2 ** 2).sum(-1)[:, None, :]
    (ref * w.double()).sum().backward()
    assert (src.grad.double() - s2.grad).abs().max() <= 1e-5 * s2.grad.abs().max()
    assert (dst.grad.double() - d2.grad).abs().max() <= 1e-5 * d2.grad.abs().max()
    only_dst = square_distance(src.detach(), dst)  # one-sided
    assert only_dst.requires_grad
