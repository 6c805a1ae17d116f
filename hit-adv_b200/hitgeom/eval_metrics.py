"""The evaluation metrics eval_ASR accumulates per batch (util/other_utils.py:37-39,72-75) -- the callers of
`pointnet2_ops` and `pytorch3d.ops` on the reference's path:

    uniform_loss(adv_pc, percentages, radius, k)       FGM/GeoA3_args.py:258-302   (the ONE in-repo pointnet2_ops caller)
    kNN_smoothing_loss(adv_pc, k, threshold_coef)      FGM/GeoA3_args.py:240-255
    CurvStdDist(k)(ori_data, adv_data, ori_normal)     util/dist_utils.py:464-495

Same names, arguments and arithmetic; what changes is where the work happens: FPS / gather / ball query / grouping /
kNN are hitgeom kernels, and `uniform_loss` runs its loop-invariant farthest-point sampling ONCE instead of once per
ball size (GeoA3_args.py:271-272 recomputes the same 5 %-of-n sample five times).
"""
import math

import torch
import torch.nn as nn

from .pointnet2_ops import pointnet2_utils
from .pytorch3d_ops import knn_gather, knn_points


def kNN_smoothing_loss(adv_pc, k, threshold_coef=1.05):
    """adv_pc [b,3,n] -> [b]: mean over points of (mean distance to the k nearest) where it exceeds mean + coef*std."""
    pts = adv_pc.permute(0, 2, 1).contiguous()
    inter = knn_points(pts, pts, K=k + 1)
    knn_dis = inter.dists[:, :, 1:].contiguous().mean(-1)
    threshold = knn_dis.mean(-1) + threshold_coef * knn_dis.std(-1)
    condition = torch.gt(knn_dis, threshold.unsqueeze(1)).float()
    return (knn_dis * condition).mean(1)


def uniform_loss(adv_pc, percentages=(0.004, 0.006, 0.008, 0.010, 0.012), radius=1.0, k=2):
    """adv_pc [b,3,n] or [b,n,3] -> scalar: for five ball sizes, FPS(5 % of n) -> ball query -> group -> kNN inside every
    ball -> squared deviation of the mean neighbour distance from the one a uniform disk would have."""
    if adv_pc.size(1) == 3:
        adv_pc = adv_pc.permute(0, 2, 1).contiguous()
    adv_pc = adv_pc.contiguous()
    b, n, _ = adv_pc.size()
    npoint = int(n * 0.05)
    flipped = adv_pc.transpose(1, 2).contiguous()
    # loop-invariant: the reference recomputes this (deterministic) sample in every iteration
    new_xyz = pointnet2_utils.gather_operation(flipped, pointnet2_utils.furthest_point_sample(adv_pc, npoint))
    new_xyz = new_xyz.transpose(1, 2).contiguous()
    loss = None
    for p in percentages:
        p = p * 4
        nsample = int(n * p)
        r = math.sqrt(p * radius)
        disk_area = math.pi * (radius ** 2) * p / nsample
        expect_len = torch.sqrt(torch.Tensor([disk_area])).to(adv_pc.device)
        idx = pointnet2_utils.ball_query(r, nsample, adv_pc, new_xyz)
        grouped = pointnet2_utils.grouping_operation(flipped, idx).permute(0, 2, 3, 1).contiguous()  # [b,npoint,ns,3]
        grouped = torch.cat(torch.unbind(grouped, dim=1), dim=0)  # [b*npoint, ns, 3]
        inter = knn_points(grouped, grouped, K=k + 1)
        uniform_dis = torch.sqrt(torch.abs(inter.dists[:, :, 1:].contiguous()) + 1e-12).mean(dim=[-1])
        uniform_dis = (uniform_dis - expect_len) ** 2 / (expect_len + 1e-12)
        mean = torch.reshape(uniform_dis, [-1]).mean() * math.pow(p * 100, 2)
        loss = mean if loss is None else loss + mean
    return loss / len(percentages)


def _normalize(v, p=2, dim=1, eps=1e-12):
    return v / v.norm(p, dim, keepdim=True).clamp(min=eps).expand_as(v)


def kappa_and_neighbours(pc, normal, k):
    """pc, normal [b,3,n] -> (kappa [b,n], idx [b,n,k+1]): mean |cos| between the normal and the directions to the k
    nearest neighbours (HiT_ADV.py:318-327, dist_utils.py:477-487)."""
    pts = pc.permute(0, 2, 1).contiguous()
    inter = knn_points(pts, pts, K=k + 1)
    nn_pts = knn_gather(pts, inter.idx).permute(0, 3, 1, 2)[:, :, :, 1:].contiguous()  # [b,3,n,k]
    vectors = _normalize(nn_pts - pc.unsqueeze(3))
    return torch.abs((vectors * normal.unsqueeze(3)).sum(1)).mean(2), inter.idx


def kappa_std(pc, normal, k):
    """Standard deviation of kappa over each point's k nearest neighbours, [b,n] (dist_utils.py:477-491)."""
    kappa, idx = kappa_and_neighbours(pc, normal, k)
    nn_kappa = knn_gather(kappa.unsqueeze(2).contiguous(), idx).permute(0, 3, 1, 2)[:, :, :, 1:].contiguous()
    return torch.std(nn_kappa.squeeze(1), dim=2)


class CurvStdDist(nn.Module):
    """util/dist_utils.py:464-495: mean l2 distance between the per-point curvature-spread profiles of the original
    and the adversarial cloud (both measured against the ORIGINAL normals, as the reference does)."""

    def __init__(self, k=5):
        super().__init__()
        self.k = k

    def forward(self, ori_data, adv_data, ori_normal):
        pdist = torch.nn.PairwiseDistance(p=2)
        return pdist(self._get_kappa_std_ori(ori_data, ori_normal, k=self.k),
                     self._get_kappa_std_ori(adv_data, ori_normal, k=self.k)).mean()

    def _get_kappa_std_ori(self, pc, normal, k=10):
        return kappa_std(pc, normal, k)
