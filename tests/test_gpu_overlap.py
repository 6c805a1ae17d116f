"""hitgeom.overlap.side_branch: an independent loss term on a forked stream -- eagerly and as a parallel arm of a CUDA
graph -- gives the serial order's loss (same bits: one deterministic kernel sequence per term) and gradient (the three
contributions are accumulated in a different order: compared norm-wise at 1e-6)."""
import numpy as np
import pytest
import torch

from util_inputs import clouds, jitter

pytestmark = pytest.mark.gpu


def _step(adv, ori, overlap, temporal=False):
    from hitgeom.dist_utils import ChamferDist, HausdorffDist, KNNDist, shared_distance_pass
    from hitgeom.overlap import side_branch

    cd, hd, kd = ChamferDist(), HausdorffDist(), KNNDist(k=5).temporal_seeds(temporal)

    def fn():
        adv.grad.zero_()
        with shared_distance_pass():
            if overlap:
                with side_branch() as br:
                    l_knn = kd(adv)
                loss = cd(adv, ori) + hd(adv, ori) + br.join(l_knn)
            else:
                loss = cd(adv, ori) + hd(adv, ori) + kd(adv)
        loss.backward()
        return loss

    return fn


def _normwise(a, b):
    return float((a - b).norm() / b.norm())


def test_side_branch_eager_matches_serial():
    ori = torch.from_numpy(clouds(48, 1024, 5)).cuda()
    adv = torch.from_numpy(jitter(clouds(48, 1024, 5), 6)).cuda().requires_grad_()
    adv.grad = torch.zeros_like(adv)
    l0 = _step(adv, ori, False)()
    g0 = adv.grad.clone()
    for _ in range(3):  # (repeats: a missing dependency would show as a changing result)
        l1 = _step(adv, ori, True)()
        torch.cuda.synchronize()
        assert float(l1) == float(l0)
        assert _normwise(adv.grad, g0) < 1e-6


@pytest.mark.parametrize("temporal", [False, True])
def test_side_branch_is_captured_as_a_parallel_arm_of_a_cuda_graph(temporal):
    ori = torch.from_numpy(clouds(64, 1024, 9)).cuda()
    adv0 = torch.from_numpy(jitter(clouds(64, 1024, 9), 10)).cuda().requires_grad_()
    adv0.grad = torch.zeros_like(adv0)
    l0 = float(_step(adv0, ori, False)())
    g0 = adv0.grad.clone()
    # a fresh leaf for the graph: its gradient accumulator must be born on the warm-up stream, not on the legacy
    # default stream the eager step above ran on (PyTorch's rule for capturing a backward pass)
    adv = adv0.detach().clone().requires_grad_()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        adv.grad = torch.zeros_like(adv)
        fn = _step(adv, ori, True, temporal)
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    for _ in range(3):
        adv.grad.fill_(7.0)  # the replay zeroes and refills it
        graph.replay()
        torch.cuda.synchronize()
        assert float(out) == l0
        assert _normwise(adv.grad, g0) < 1e-6


def test_side_branch_refuses_the_current_stream():
    from hitgeom.overlap import side_branch

    with pytest.raises(RuntimeError):
        with side_branch(stream=torch.cuda.current_stream()):
            pass
