"""Run an independent part of a loss on a forked CUDA stream.

The three loss terms of an attack step (ChamferDist / HausdorffDist on (adv, ori), KNNDist on adv alone;
CW/kNN.py:104-108, util/other_utils.py:37-39) do not depend on one another, and at attack sizes (a few hundred clouds
of 1024 points) none of their kernels fills 148 SMs on its own: the small-cloud kNN kernel ends in a long
low-occupancy tail, the distance pass is 2.6 CTAs per SM.  Launched on two streams they share the machine:

    with side_branch() as br:            # forks from the current stream
        l_knn = knn_dist(adv)            # enqueued on the side stream
    loss = chamfer(adv, ori) + hausdorff(adv, ori) + br.join(l_knn)     # join: current stream waits for the branch
    loss.backward()                      # autograd runs each backward node on its forward stream: overlapped again

Works eagerly and under CUDA-graph capture (the fork and the join are event dependencies of the capturing stream, so
the branch is captured as a parallel arm of the graph).  PyTorch's rule for capturing a backward pass applies: the leaf
tensors must have run their first backward on the warm-up side stream, not on the legacy default stream (create them,
or re-create them with `.detach().requires_grad_()`, inside the warm-up stream context; tests/test_gpu_overlap.py).  Results are the same bits as the serial order: every kernel
is deterministic and the terms are combined in the same expression.
"""
import torch

_side_streams = {}


def _side_stream(device):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class side_branch:
    def __init__(self, device=None, stream=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.side = stream
        self._ctx = None

    def __enter__(self):
        self.main = torch.cuda.current_stream(self.device)
        if self.side is None:
            self.side = _side_stream(self.device)
        if self.side == self.main:
            raise RuntimeError("side_branch: the side stream is the current stream")
        self.side.wait_stream(self.main)  # everything enqueued so far (the inputs) is visible to the branch
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        self._ctx = None
        return False

    def join(self, *tensors):
        """Makes the current stream wait for the branch and hands its results over (allocator bookkeeping included)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.side)
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(cur)
        return tensors[0] if len(tensors) == 1 else tensors
