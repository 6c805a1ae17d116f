"""Autograd layer over hitgeom's `_ext`, exposing the public names of the reference's
`pointnet2_ops.pointnet2_utils` (pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py:65,101,136,191,240,276 and
the `QueryAndGroup` / `GroupAll` modules :279-379) with the same call signatures:

    furthest_point_sample(xyz, npoint)            -> (B,npoint) int32          no grad
    gather_operation(features, idx)               -> (B,C,npoint)              grad -> features
    three_nn(unknown, known)                      -> (dist, idx)  dist = sqrt(d2)  no grad
    three_interpolate(features, idx, weight)      -> (B,c,n)                   grad -> features
    grouping_operation(features, idx)             -> (B,C,npoint,nsample)      grad -> features
    ball_query(radius, nsample, xyz, new_xyz)     -> (B,npoint,nsample) int32  no grad

When the reference tree itself is importable, `hitgeom.install()` simply registers `_ext` as
`pointnet2_ops._ext` and the reference's own wrapper file runs unchanged on top of it; this module is the
stand-alone equivalent (used by the tests and the benchmark on boxes without the reference).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext


def _index_op(name, native):
    """Function with integer outputs only: nothing flows back (reference: mark_non_differentiable + `()`)."""

    def forward(ctx, *args):
        out = native(*args)
        ctx.mark_non_differentiable(*(out if isinstance(out, tuple) else (out,)))
        return out

    def backward(ctx, *grads):
        return ()

    return type(name, (Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward)})


def _feature_op(name, native_fwd, native_bwd, n_extra):
    """Function differentiable in its first argument (B,C,N); `n_extra` index/weight arguments follow."""

    def forward(ctx, features, *extra):
        ctx.n_src = features.size(2)
        ctx.save_for_backward(*extra)
        return native_fwd(features, *extra)

    def backward(ctx, grad_out):
        extra = ctx.saved_tensors
        grad_features = native_bwd(grad_out.contiguous(), *extra, ctx.n_src)
        # index inputs get None, except where the reference hands back zeros_like placeholders
        return (grad_features,) + tuple(torch.zeros_like(e) if n_extra_zero else None for e in extra)

    n_extra_zero = name != "GatherOperation"  # pointnet2_utils.py:98 returns None, :188/:237 zeros_like
    return type(name, (Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward)})


FurthestPointSampling = _index_op("FurthestPointSampling", lambda xyz, npoint: _ext.furthest_point_sampling(xyz, npoint))
BallQuery = _index_op("BallQuery", lambda radius, nsample, xyz, new_xyz: _ext.ball_query(new_xyz, xyz, radius, nsample))


def _three_nn_native(unknown, known):
    dist2, idx = _ext.three_nn(unknown, known)
    return torch.sqrt(dist2), idx


ThreeNN = _index_op("ThreeNN", _three_nn_native)
GatherOperation = _feature_op("GatherOperation", _ext.gather_points, _ext.gather_points_grad, 1)
GroupingOperation = _feature_op("GroupingOperation", _ext.group_points, _ext.group_points_grad, 1)
ThreeInterpolate = _feature_op("ThreeInterpolate", _ext.three_interpolate, _ext.three_interpolate_grad, 2)

furthest_point_sample = FurthestPointSampling.apply
gather_operation = GatherOperation.apply
three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply
grouping_operation = GroupingOperation.apply
ball_query = BallQuery.apply


class GroupConcat(Function):
    """(xyz (B,N,3), new_xyz (B,S,3), features (B,C,N) | None, idx) -> (B,3+C,S,ns) = cat(xyz[idx] - new_xyz, features[idx])
    in one pass; differentiable in xyz, new_xyz and features like the reference's composition."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, idx):
        ctx.n = xyz.size(1)
        ctx.save_for_backward(idx)
        return _ext.group_concat(xyz.contiguous(), new_xyz.contiguous(),
                                 None if features is None else features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        need_xyz, need_new, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g = grad_out.contiguous()
        gx_t, gf = _ext.group_concat_grad(g, idx, ctx.n, need_xyz, need_f)
        gx = gx_t.transpose(1, 2).contiguous() if gx_t is not None else None
        gn = -g[:, :3].sum(dim=3).transpose(1, 2).contiguous() if need_new else None
        return gx, gn, gf, None


class QueryAndGroup(nn.Module):
    """ball_query -> group the coordinates and subtract the centre -> optional concat with grouped features
    (pointnet2_utils.py:279-333); with `use_xyz` the whole body after the ball query is one fused pass."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return GroupConcat.apply(xyz, new_xyz, None, idx)
        if self.use_xyz:
            return GroupConcat.apply(xyz, new_xyz, features, idx)
        return grouping_operation(features, idx)


class GroupAll(nn.Module):
    """One group holding every point: (B,3[+C],1,N)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        all_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return all_xyz
        feats = features.unsqueeze(2)
        return torch.cat([all_xyz, feats], dim=1) if self.use_xyz else feats
