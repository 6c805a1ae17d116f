"""`pointnet2_ops._ext` -- the nine functions the reference's pybind11 module exports
(pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19), same names, argument order, dtypes and
shapes, implemented by handing raw pointers to libhitgeom.so.  The reference allocates its outputs with
`torch::zeros` because its kernels skip entries (an empty ball keeps its zeros, the atomicAdd gradients need a zero
start); hitgeom's kernels write EVERY output element (same values, an empty ball is written as zeros), so the
outputs are `torch.empty` -- the memset was up to 40 % of a group_points call.

Host-side behaviour mirrored from the reference's .cpp wrappers (ball_query.cpp:8-32, group_points.cpp:12-62,
sampling.cpp:15-87, interpolate.cpp:14-99): contiguity / dtype checks raise RuntimeError (the reference's
AT_ASSERT), CPU tensors are rejected ("CPU not supported"), launches are asynchronous on the current stream.
Differences, all deliberate: a device guard is taken (the reference launches on the current device whatever
the tensor's device is), launch failures raise instead of calling exit(-1) (cuda_utils.h:30-39), and the
*_grad functions are deterministic.
"""
import torch

from .._lib import check, lib, ptr, require, stream_ptr, workspace


def _cuda_or_raise(t, name):
    if not t.is_cuda:
        raise RuntimeError("CPU not supported")  # sampling.cpp:35, ball_query.cpp:28, ...
    return t


def gather_points(points, idx):
    """points (B,C,N) f32, idx (B,M) i32 -> (B,C,M) f32   [sampling.cpp:15-40]"""
    _cuda_or_raise(points, "points")
    require(points, "points", torch.float32, 3)
    require(idx, "idx", torch.int32, 2)
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(lib().hg_p2_gather_points(B, C, N, M, ptr(points), ptr(idx), ptr(out), stream_ptr()), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """grad_out (B,C,M), idx (B,M) -> (B,C,n)   [sampling.cpp:42-65]"""
    _cuda_or_raise(grad_out, "grad_out")
    require(grad_out, "grad_out", torch.float32, 3)
    require(idx, "idx", torch.int32, 2)
    B, C, M = grad_out.shape
    out = torch.empty((B, C, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        ws = workspace(lib().hg_p2_scatter_workspace_bytes(B, n, M), grad_out.device)
        check(lib().hg_p2_gather_points_grad(B, C, n, M, ptr(grad_out), ptr(idx), ptr(out), ptr(ws), ws.numel(),
                                             stream_ptr()), "gather_points_grad")
    return out


def furthest_point_sampling(points, nsamples):
    """points (B,N,3) f32 -> (B,nsamples) i32   [sampling.cpp:66-87]"""
    _cuda_or_raise(points, "points")
    require(points, "points", torch.float32, 3)
    B, N, _ = points.shape
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        check(lib().hg_p2_furthest_point_sampling(B, N, nsamples, ptr(points), None, ptr(out), stream_ptr()),
              "furthest_point_sampling")
    return out


def three_nn(unknowns, knows):
    """unknowns (B,n,3), knows (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) i32]   [interpolate.cpp:14-40]"""
    _cuda_or_raise(unknowns, "unknowns")
    require(unknowns, "unknowns", torch.float32, 3)
    require(knows, "knows", torch.float32, 3)
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        check(lib().hg_p2_three_nn(B, n, m, ptr(unknowns), ptr(knows), ptr(dist2), ptr(idx), stream_ptr()), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """points (B,c,m), idx (B,n,3) i32, weight (B,n,3) -> (B,c,n)   [interpolate.cpp:42-70]"""
    _cuda_or_raise(points, "points")
    require(points, "points", torch.float32, 3)
    require(idx, "idx", torch.int32, 3)
    require(weight, "weight", torch.float32, 3)
    B, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, c, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(lib().hg_p2_three_interpolate(B, c, m, n, ptr(points), ptr(idx), ptr(weight), ptr(out), stream_ptr()),
              "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """grad_out (B,c,n), idx/weight (B,n,3) -> (B,c,m)   [interpolate.cpp:71-99]"""
    _cuda_or_raise(grad_out, "grad_out")
    require(grad_out, "grad_out", torch.float32, 3)
    require(idx, "idx", torch.int32, 3)
    require(weight, "weight", torch.float32, 3)
    B, c, n = grad_out.shape
    out = torch.empty((B, c, m), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        ws = workspace(lib().hg_p2_scatter_workspace_bytes(B, m, n * 3), grad_out.device)
        check(lib().hg_p2_three_interpolate_grad(B, c, n, m, ptr(grad_out), ptr(idx), ptr(weight), ptr(out), ptr(ws),
                                                 ws.numel(), stream_ptr()), "three_interpolate_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32   [ball_query.cpp:8-32]"""
    _cuda_or_raise(new_xyz, "new_xyz")
    require(new_xyz, "new_xyz", torch.float32, 3)
    require(xyz, "xyz", torch.float32, 3)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = torch.empty((B, M, nsample), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        check(lib().hg_p2_ball_query(B, N, M, float(radius), int(nsample), ptr(new_xyz), ptr(xyz), ptr(idx),
                                     stream_ptr()), "ball_query")
    return idx


def group_points(points, idx):
    """points (B,C,N), idx (B,S,ns) i32 -> (B,C,S,ns)   [group_points.cpp:12-36]"""
    _cuda_or_raise(points, "points")
    require(points, "points", torch.float32, 3)
    require(idx, "idx", torch.int32, 3)
    B, C, N = points.shape
    _, S, ns = idx.shape
    out = torch.empty((B, C, S, ns), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(lib().hg_p2_group_points(B, C, N, S, ns, ptr(points), ptr(idx), ptr(out), stream_ptr()), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """grad_out (B,C,S,ns), idx (B,S,ns) -> (B,C,n)   [group_points.cpp:38-62]"""
    _cuda_or_raise(grad_out, "grad_out")
    require(grad_out, "grad_out", torch.float32, 4)
    require(idx, "idx", torch.int32, 3)
    B, C, S, ns = grad_out.shape
    out = torch.empty((B, C, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        ws = workspace(lib().hg_p2_scatter_workspace_bytes(B, n, S * ns), grad_out.device)
        check(lib().hg_p2_group_points_grad(B, C, n, S, ns, ptr(grad_out), ptr(idx), ptr(out), ptr(ws), ws.numel(),
                                            stream_ptr()), "group_points_grad")
    return out


def group_concat(xyz, new_xyz, features, idx):
    """hitgeom extension (no counterpart in bindings.cpp): the body of QueryAndGroup.forward after its ball query
    (pointnet2_utils.py:312-333) in one pass.  xyz (B,N,3), new_xyz (B,S,3), features (B,C,N) or None, idx (B,S,ns) i32
    -> (B,3+C,S,ns): rows 0..2 = xyz[idx] - new_xyz, rows 3.. = features[idx]."""
    _cuda_or_raise(xyz, "xyz")
    require(xyz, "xyz", torch.float32, 3)
    require(new_xyz, "new_xyz", torch.float32, 3)
    require(idx, "idx", torch.int32, 3)
    B, N, _ = xyz.shape
    _, S, ns = idx.shape
    C = 0
    if features is not None:
        require(features, "features", torch.float32, 3)
        C = features.shape[1]
    out = torch.empty((B, 3 + C, S, ns), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib().hg_p2_group_concat(B, C, N, S, ns, ptr(xyz), ptr(new_xyz), ptr(features), ptr(idx), ptr(out),
                                       stream_ptr()), "group_concat")
    return out


def group_concat_grad(grad_out, idx, n, want_xyz, want_features):
    """grad_out (B,3+C,S,ns) -> (d/d xyz as (B,3,n) or None, d/d features (B,C,n) or None)."""
    _cuda_or_raise(grad_out, "grad_out")
    require(grad_out, "grad_out", torch.float32, 4)
    require(idx, "idx", torch.int32, 3)
    B, C3, S, ns = grad_out.shape
    C = C3 - 3
    gx = torch.empty((B, 3, n), dtype=torch.float32, device=grad_out.device) if want_xyz else None
    gf = torch.empty((B, C, n), dtype=torch.float32, device=grad_out.device) if (want_features and C > 0) else None
    if gx is None and gf is None:
        return None, None
    with torch.cuda.device(grad_out.device):
        ws = workspace(lib().hg_p2_scatter_workspace_bytes(B, n, S * ns), grad_out.device)
        check(lib().hg_p2_group_concat_grad(B, C, n, S, ns, ptr(grad_out), ptr(idx), ptr(gx), ptr(gf), ptr(ws), ws.numel(),
                                            stream_ptr()), "group_concat_grad")
    return gx, gf
