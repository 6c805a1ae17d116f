"""Drop-in mirror of the hot-path loss classes of the reference's `util/dist_utils.py`:
`ChamferDist` (:44-80), `HausdorffDist` (:83-119), `KNNDist` (:122-175), `ChamferkNNDist` (:258-294) and
`L2Dist` (:15-41, trivial, kept so that `from util.dist_utils import *` users find it).

Same constructor arguments, same `forward(adv_pc, ori_pc, weights=None, batch_avg=True)` contract, same
return values (0-d tensor, or [B] with `batch_avg=False`); weights may arrive as float64 CPU tensors
(CW/Perturb.py:148-150) and are moved with `.float().cuda()` exactly as the reference does.
"""
import torch
import torch.nn as nn

from . import functional as F
from .functional import shared_distance_pass  # noqa: F401  (opt-in: Chamfer + Hausdorff of one pair share a pass)
from .set_distance import chamfer, hausdorff


def _weights(weights, B, device):
    """`weights.float().cuda()` of the reference (dist_utils.py:76); None means unit weights, for which the
    multiplication is skipped (x * 1.0 == x exactly) instead of building a ones tensor on the host every call."""
    if weights is None:
        return None
    return weights.float().to(device)


def _finish(loss, weights, batch_avg):
    if weights is not None:
        loss = loss * weights
    if batch_avg:
        return loss.mean()
    return loss


class L2Dist(nn.Module):
    def __init__(self):
        super(L2Dist, self).__init__()

    def forward(self, adv_pc, ori_pc, weights=None, batch_avg=True):
        B = adv_pc.shape[0]
        weights = _weights(weights, B, adv_pc.device)
        dist = torch.sqrt(torch.sum((adv_pc - ori_pc) ** 2, dim=[1, 2]) + torch.tensor(1e-7))
        return _finish(dist, weights, batch_avg)


def _select(loss1, loss2, method):
    if method == 'adv2ori':
        return loss1
    elif method == 'ori2adv':
        return loss2
    return (loss1 + loss2) / 2.


class ChamferDist(nn.Module):
    def __init__(self, method='adv2ori'):
        super(ChamferDist, self).__init__()
        self.method = method

    def forward(self, adv_pc, ori_pc, weights=None, batch_avg=True):
        B = adv_pc.shape[0]
        loss1, loss2 = chamfer(adv_pc, ori_pc)  # [B], adv2ori, ori2adv
        loss = _select(loss1, loss2, self.method)
        return _finish(loss, _weights(weights, B, loss.device), batch_avg)


class HausdorffDist(nn.Module):
    def __init__(self, method='adv2ori'):
        super(HausdorffDist, self).__init__()
        self.method = method

    def forward(self, adv_pc, ori_pc, weights=None, batch_avg=True):
        B = adv_pc.shape[0]
        loss1, loss2 = hausdorff(adv_pc, ori_pc)
        loss = _select(loss1, loss2, self.method)
        return _finish(loss, _weights(weights, B, loss.device), batch_avg)


class KNNDist(nn.Module):
    """Same constructor as the reference.  `temporal_seeds(True)` (an addition) keeps the neighbour indices of the
    previous call as state and hands them to the next call as seeds: inside an attack loop (CW/kNN.py:77-111) the cloud
    moves by a learning-rate step per iteration, so they bound every k-th distance almost tightly and replace the
    spatial pre-pass.  The loss and its gradient are the same bits either way; the state is per (batch shape, device)
    and replay-safe (the kernel updates it in place)."""

    def __init__(self, k=5, alpha=1.05):
        super(KNNDist, self).__init__()
        self.k = k
        self.alpha = alpha
        self.temporal = False
        self._state = {}

    def temporal_seeds(self, on=True):
        self.temporal = bool(on)
        self._state.clear()
        return self

    def forward(self, pc, weights=None, batch_avg=True):
        """pc: [B, K, 3] or [B, 3, K] (dist_utils.py:145-147 treats shape[1] == 3 as channel-first)."""
        B = pc.shape[0]
        if pc.shape[1] == 3:
            pc = pc.transpose(2, 1)  # kernels are point-major
        state, valid = None, False
        # (getattr: install.patch_reference() binds this forward to the REFERENCE's KNNDist class, whose __init__ knows
        # nothing of the temporal state)
        if getattr(self, "temporal", False) and pc.is_cuda and pc.shape[2] == 3:
            key = (tuple(pc.shape), pc.device)
            valid = key in self._state
            if not valid:
                self._state[key] = torch.empty((pc.shape[0], pc.shape[1], self.k + 1), dtype=torch.int32, device=pc.device)
            state = self._state[key]
        loss = F.knn_outlier_loss(pc, self.k, self.alpha, state, valid)  # [B]
        return _finish(loss, _weights(weights, B, loss.device), batch_avg)


class ChamferkNNDist(nn.Module):
    def __init__(self, chamfer_method='adv2ori', knn_k=5, knn_alpha=1.05, chamfer_weight=5., knn_weight=3.):
        super(ChamferkNNDist, self).__init__()
        self.chamfer_dist = ChamferDist(method=chamfer_method)
        self.knn_dist = KNNDist(k=knn_k, alpha=knn_alpha)
        self.w1 = chamfer_weight
        self.w2 = knn_weight

    def temporal_seeds(self, on=True):
        """See KNNDist.temporal_seeds."""
        self.knn_dist.temporal_seeds(on)
        return self

    def forward(self, adv_pc, ori_pc, weights=None, batch_avg=True):
        chamfer_loss = self.chamfer_dist(adv_pc, ori_pc, weights=weights, batch_avg=batch_avg)
        knn_loss = self.knn_dist(adv_pc, weights=weights, batch_avg=batch_avg)
        loss = chamfer_loss * self.w1 + knn_loss * self.w2
        return loss
