"""Clipping / projection steps the CW loops apply after every optimiser step (util/clip_utils.py).

Same class names, constructor arguments and `forward(pc, ori_pc[, normal])` contracts ([B,3,K] channel-first
clouds, result detached).  Plain elementwise torch on the device the clouds live on -- these run once per
iteration on B*3*K floats and are not kernels of their own; what differs from the reference is only that the
masked in-place assignments (`diff[mask] = proj[mask]`, clip_utils.py:134-141: two boolean-index gathers and
scatters, each with a device->host sync for the element count) are `torch.where` selections: same values, no sync.
"""
import torch
import torch.nn as nn


class ClipPointsL2(nn.Module):
    """util/clip_utils.py:5-31: scale the whole perturbation of a cloud back onto the l2 ball of radius budget."""

    def __init__(self, budget):
        super().__init__()
        self.budget = budget

    def forward(self, pc, ori_pc):
        with torch.no_grad():
            diff = pc - ori_pc
            norm = torch.sum(diff ** 2, dim=[1, 2]) ** 0.5
            scale = torch.clamp(self.budget / (norm + 1e-9), max=1.)
            return ori_pc + diff * scale[:, None, None]


class ClipPointsLinf(nn.Module):
    """util/clip_utils.py:62-85: clamp every coordinate of the perturbation to [-budget, budget]."""

    def __init__(self, budget):
        super().__init__()
        self.budget = budget

    def forward(self, pc, ori_pc):
        with torch.no_grad():
            return (ori_pc + torch.clamp(pc - ori_pc, min=-self.budget, max=self.budget)).detach()


class ProjectInnerPoints(nn.Module):
    """util/clip_utils.py:89-142: perturbations pointing into the object (diff . normal < 0) are projected onto
    vref = (normal x diff) x normal; those exactly opposite to the normal are zeroed."""

    def forward(self, pc, ori_pc, normal=None):
        with torch.no_grad():
            if normal is None:
                return pc
            diff = pc - ori_pc
            inner = torch.sum(diff * normal, dim=1) < 0.  # [B,K]
            vng = torch.cross(normal, diff, dim=1)
            vng_norm = torch.sum(vng ** 2, dim=1) ** 0.5
            vref = torch.cross(vng, normal, dim=1)
            vref_norm = torch.sum(vref ** 2, dim=1) ** 0.5
            proj = diff * vref / (vref_norm[:, None, :] + 1e-9)
            proj = torch.where((inner & (vng_norm < 1e-6))[:, None, :], torch.zeros_like(proj), proj)
            return ori_pc + torch.where(inner[:, None, :], proj, diff)


class ProjectInnerClipLinf(nn.Module):
    """util/clip_utils.py:145-170: project, then clip."""

    def __init__(self, budget):
        super().__init__()
        self.project_inner = ProjectInnerPoints()
        self.clip_linf = ClipPointsLinf(budget=budget)

    def forward(self, pc, ori_pc, normal=None):
        with torch.no_grad():
            return self.clip_linf(self.project_inner(pc, ori_pc, normal), ori_pc)
