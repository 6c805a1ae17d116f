"""Set-abstraction / feature-propagation modules over hitgeom's pointnet2 operators -- the public names of the
reference's `pointnet2_ops.pointnet2_modules` (pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:9-209; SURVEY.md
section 8a row a18) with the same constructor arguments, attribute names (`npoint`, `groupers`, `mlps`, `mlp`: state
dicts are interchangeable) and tensor contracts:

    PointnetSAModuleMSG(npoint, radii, nsamples, mlps, bn=True, use_xyz=True)
    PointnetSAModule(mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True)
        forward(xyz (B,N,3), features (B,C,N) | None) -> new_xyz (B,npoint,3) | None, new_features (B, sum mlp[-1], npoint)
    PointnetFPModule(mlp, bn=True)
        forward(unknown (B,n,3), known (B,m,3) | None, unknow_feats (B,C1,n) | None, known_feats (B,C2,m)) -> (B,mlp[-1],n)

The geometry (FPS, gather, ball query, grouping, three_nn, three_interpolate) runs on the native kernels; the shared
MLPs and the max-pool stay PyTorch.  Like the reference, `mlps[i][0]` is incremented by 3 IN THE CALLER'S LIST when
`use_xyz` (pointnet2_modules.py:115-116).
"""
import torch
import torch.nn as nn

from . import ops as pointnet2_utils


def build_shared_mlp(mlp_spec, bn=True):
    layers = []
    for c_in, c_out in zip(mlp_spec[:-1], mlp_spec[1:]):
        layers.append(nn.Conv2d(c_in, c_out, kernel_size=1, bias=not bn))
        if bn:
            layers.append(nn.BatchNorm2d(c_out))
        layers.append(nn.ReLU(True))
    return nn.Sequential(*layers)


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz, features):
        new_xyz = None
        if self.npoint is not None:
            centres = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), centres)
            new_xyz = new_xyz.transpose(1, 2).contiguous()
        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = mlp(grouper(xyz, new_xyz, features))  # (B, mlp[-1], npoint, nsample)
            pooled.append(grouped.max(dim=3)[0])  # == max_pool2d over the sample axis + squeeze
        return new_xyz, torch.cat(pooled, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, mlp_spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                mlp_spec[0] += 3
            self.mlps.append(build_shared_mlp(mlp_spec, bn))


class PointnetSAModule(PointnetSAModuleMSG):
    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            # the reference's `size()[0:2] + [n]` (pointnet2_modules.py:194-196) raises TypeError (torch.Size + list);
            # this is what it means
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
