"""DGCNN edge features (hg_edge_feature_*, model_seams.get_graph_feature) against the golden output of the
UNMODIFIED reference function (tests/golden/dgcnn_edge.npz) and, at DGCNN's real sizes, against the oracle's torch
restatement of the reference's tensor program run on the same GPU.  Forward: bit-exact (one subtraction per
element).  Backward: the reference's `index_put_(accumulate=True)` sums each point's incoming edges with
floating-point atomics in arbitrary order, hitgeom in ascending edge order => 1e-5 norm-wise per sample, and
bit-reproducible run to run."""
import numpy as np
import pytest
import torch

from util_inputs import normwise

pytestmark = pytest.mark.gpu


def test_edge_feature_matches_reference_golden(golden):
    from hitgeom import functional as F
    from hitgeom.model_seams import get_graph_feature

    g = golden("dgcnn_edge")
    for tag, C in (("a", 3), ("b", 64), ("c", 5)):
        x = torch.from_numpy(g[f"{tag}_x"]).cuda().requires_grad_()
        idx = torch.from_numpy(g[f"{tag}_idx"]).cuda()
        feat = F.edge_feature(x, idx)
        assert np.array_equal(feat.detach().cpu().numpy(), g[f"{tag}_feat"]), tag
        gen = torch.Generator().manual_seed(70 + C)
        torch.randn(x.shape, generator=gen)
        w = torch.randn(feat.shape, generator=gen).cuda()
        (feat * w).sum().backward()
        assert normwise(x.grad.cpu().numpy(), g[f"{tag}_grad"]) < 1e-5, tag
        # through the seam (own neighbour search): same neighbour SETS as the reference's topk => same features up
        # to the order within a row; with C == 3 the order is the reference's too (distinct distances)
        k = idx.shape[2]
        feat2 = get_graph_feature(x.detach(), k=k)
        assert feat2.shape == feat.shape
        if C == 3:
            assert np.array_equal(feat2.cpu().numpy(), g[f"{tag}_feat"]), tag


@pytest.mark.parametrize("B,C,N,k", [(32, 64, 1024, 20), (4, 128, 1024, 20), (3, 3, 1024, 20), (2, 6, 300, 30),
                                     (2, 16, 4096, 20), (1, 7, 33, 5)])
def test_edge_feature_vs_torch_program_on_gpu(B, C, N, k):
    from hitgeom import functional as F
    from oracle import torch_port as tp

    gen = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(B, C, N, generator=gen).cuda()
    idx = torch.stack([torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(N)]) for _ in range(min(B, 2))])
    idx = idx.repeat((B + 1) // 2, 1, 1)[:B].contiguous().cuda()
    idx[:, :, 0] = torch.arange(N, device="cuda")[None, :]  # self edge first, like DGCNN's kNN
    idx[:, : N // 8, 1] = 0  # a hub: point 0 receives N/8 extra edges (long CSR segment)
    w = torch.randn(B, 2 * C, N, k, generator=gen).cuda()
    xr = x.clone().requires_grad_()
    ref = tp.get_graph_feature(xr, idx)
    (ref * w).sum().backward()
    xm = x.clone().requires_grad_()
    out = F.edge_feature(xm, idx)
    assert torch.equal(out, ref)
    (out * w).sum().backward()
    g1 = xm.grad.clone()
    assert normwise(g1.cpu().numpy(), xr.grad.cpu().numpy()) < 1e-5
    # float64 yardstick: no further from it than the reference's own FP32 path (x4 slack)
    xd = x.double().requires_grad_()
    (tp.get_graph_feature(xd, idx) * w.double()).sum().backward()
    e_ref = normwise(xr.grad.double().cpu().numpy(), xd.grad.cpu().numpy())
    e_me = normwise(g1.double().cpu().numpy(), xd.grad.cpu().numpy())
    assert e_me <= max(4 * e_ref, 2e-6), (e_me, e_ref)
    # deterministic: a second backward gives the same bits
    xm.grad = None
    (F.edge_feature(xm, idx) * w).sum().backward()
    assert torch.equal(xm.grad, g1)


def test_edge_feature_rejects_bad_arguments():
    from hitgeom import HitgeomError
    from hitgeom import functional as F

    x = torch.randn(2, 3, 64).cuda()
    with pytest.raises(HitgeomError):
        F.edge_feature(x, torch.zeros(2, 64, 4, dtype=torch.int32).cuda())
    with pytest.raises(RuntimeError, match="CUDA tensor"):  # the reference's CHECK_CUDA behaviour
        F.edge_feature(x.cpu(), torch.zeros(2, 64, 4, dtype=torch.int64))
