"""numpy/ctypes front end of the CPU oracle (oracle/hitgeom_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package never imports this module (tests/test_no_oracle_in_product.py enforces it).

Every function takes and returns numpy arrays (float32 / int32 / int64, C-contiguous) and restates one
reference function; the docstrings cite the reference file:line that the C code follows.  Work is split
over host threads cloud-by-cloud (`threads=`): ctypes drops the GIL during the C call.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_l = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile liboracle.so with gcc (a few seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "hitgeom_oracle.c")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _chunks(B, threads):
    threads = max(1, min(int(threads or 1), B if B > 0 else 1))
    edges = np.linspace(0, B, threads + 1).astype(int)
    return [(int(edges[t]), int(edges[t + 1])) for t in range(threads) if edges[t + 1] > edges[t]]


def _par(B, threads, fn):
    """Run fn(b0, b1) over contiguous cloud ranges on `threads` host threads."""
    ch = _chunks(B, threads)
    if len(ch) <= 1:
        for b0, b1 in ch:
            fn(b0, b1)
        return
    with ThreadPoolExecutor(len(ch)) as ex:
        list(ex.map(lambda r: fn(*r), ch))


def host_threads():
    return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------
# util/set_distance.py
# ------------------------------------------------------------------------------------------------------------
def pairwise_dist(x, y):
    """`_Distance.batch_pairwise_dist(x, y)` util/set_distance.py:15-32 -> P [B,Nx,Ny]."""
    x, y = _c(x, np.float32), _c(y, np.float32)
    B, Nx, D = x.shape
    Ny = y.shape[1]
    P = np.empty((B, Nx, Ny), np.float32)
    lib().orc_pairwise_dist(_p(x, _f), _p(y, _f), B, Nx, Ny, D, _p(P, _f))
    return P


def nn_bidir(gts, preds, threads=1):
    """The two `torch.min` reductions of util/set_distance.py:45-49 without storing P.

    Returns (min1 [B,N1], arg1 [B,N1], min2 [B,N2], arg2 [B,N2]); *1 = per pred/adv point over gts
    (`torch.min(P,1)`), *2 = per gt/ori point over preds (`torch.min(P,2)`)."""
    gts, preds = _c(gts, np.float32), _c(preds, np.float32)
    B, N2, D = gts.shape
    N1 = preds.shape[1]
    min1, arg1 = np.empty((B, N1), np.float32), np.empty((B, N1), np.int32)
    min2, arg2 = np.empty((B, N2), np.float32), np.empty((B, N2), np.int32)

    def run(b0, b1):
        lib().orc_nn_bidir(_p(gts[b0:b1], _f), _p(preds[b0:b1], _f), b1 - b0, N2, N1, D, _p(min1[b0:b1], _f),
                           _p(arg1[b0:b1], _i), _p(min2[b0:b1], _f), _p(arg2[b0:b1], _i))

    _par(B, threads, run)
    return min1, arg1, min2, arg2


def set_loss(min1, min2, mode):
    """Chamfer (mode 0, set_distance.py:46-49) / Hausdorff (mode 1, :66-69) reductions of the mins."""
    min1, min2 = _c(min1, np.float32), _c(min2, np.float32)
    B, N1 = min1.shape
    N2 = min2.shape[1]
    l1, l2 = np.empty(B, np.float32), np.empty(B, np.float32)
    h1, h2 = np.empty(B, np.int32), np.empty(B, np.int32)
    lib().orc_set_loss(_p(min1, _f), _p(min2, _f), B, N1, N2, mode, _p(l1, _f), _p(l2, _f), _p(h1, _i), _p(h2, _i))
    return l1, l2, h1, h2


def set_loss_bwd(gts, preds, arg1, arg2, hd1, hd2, g1, g2, mode, want_gts=False):
    """Analytic backward of chamfer / hausdorff w.r.t. preds (and gts) through the saved indices."""
    gts, preds = _c(gts, np.float32), _c(preds, np.float32)
    B, N2, D = gts.shape
    N1 = preds.shape[1]
    arg1, arg2 = _c(arg1, np.int32), _c(arg2, np.int32)
    hd1 = _c(hd1 if hd1 is not None else np.zeros(B), np.int32)
    hd2 = _c(hd2 if hd2 is not None else np.zeros(B), np.int32)
    g1, g2 = _c(g1, np.float32), _c(g2, np.float32)
    gp = np.empty_like(preds)
    gg = np.empty_like(gts) if want_gts else None
    lib().orc_set_loss_bwd(_p(gts, _f), _p(preds, _f), _p(arg1, _i), _p(arg2, _i), _p(hd1, _i), _p(hd2, _i),
                           _p(g1, _f), _p(g2, _f), B, N2, N1, D, mode, _p(gp, _f), _p(gg, _f))
    return (gp, gg) if want_gts else gp


def chamfer(preds, gts, threads=1):
    """`ChamferDistance.forward(preds, gts)` util/set_distance.py:40-50 -> (loss1 [B], loss2 [B])."""
    m1, _, m2, _ = nn_bidir(gts, preds, threads)
    l1, l2, _, _ = set_loss(m1, m2, 0)
    return l1, l2


def hausdorff(preds, gts, threads=1):
    """`HausdorffDistance.forward(preds, gts)` util/set_distance.py:58-70."""
    m1, _, m2, _ = nn_bidir(gts, preds, threads)
    l1, l2, _, _ = set_loss(m1, m2, 1)
    return l1, l2


# ------------------------------------------------------------------------------------------------------------
# util/dist_utils.py KNNDist / model/dgcnn_cls.py knn
# ------------------------------------------------------------------------------------------------------------
def knn_self(pc, k1, threads=1):
    """k1 smallest entries per row of the KNNDist / DGCNN matrix (dist_utils.py:148-156, dgcnn_cls.py:8-12).

    pc: point-major [B,K,C].  Returns (vals [B,K,k1] ascending, idx [B,K,k1] int32, lowest index on ties)."""
    pc = _c(pc, np.float32)
    B, K, C = pc.shape
    vals, idx = np.empty((B, K, k1), np.float32), np.empty((B, K, k1), np.int32)

    def run(b0, b1):
        rc = lib().orc_knn_self(_p(pc[b0:b1], _f), b1 - b0, K, C, k1, _p(vals[b0:b1], _f), _p(idx[b0:b1], _i))
        if rc:
            raise ValueError("k+1 > number of points")

    _par(B, threads, run)
    return vals, idx


def knn_outlier_fwd(vals, alpha):
    """dist_utils.py:157-167 -> (value [B,K], mask [B,K], loss [B])."""
    vals = _c(vals, np.float32)
    B, K, k1 = vals.shape
    value, mask, loss = np.empty((B, K), np.float32), np.empty((B, K), np.float32), np.empty(B, np.float32)
    lib().orc_knn_outlier_fwd(_p(vals, _f), B, K, k1, ctypes.c_float(alpha), _p(value, _f), _p(mask, _f), _p(loss, _f))
    return value, mask, loss


def knn_outlier_bwd(pc, idx, mask, g):
    pc, idx, mask, g = _c(pc, np.float32), _c(idx, np.int32), _c(mask, np.float32), _c(g, np.float32)
    B, K, C = pc.shape
    out = np.empty_like(pc)
    lib().orc_knn_outlier_bwd(_p(pc, _f), _p(idx, _i), _p(mask, _f), _p(g, _f), B, K, C, idx.shape[2], _p(out, _f))
    return out


def knn_dist(pc, k=5, alpha=1.05, threads=1):
    """`KNNDist(k, alpha).forward(pc, batch_avg=False)` with unit weights -> loss [B] (dist_utils.py:136-175)."""
    vals, idx = knn_self(pc, k + 1, threads)
    value, mask, loss = knn_outlier_fwd(vals, alpha)
    return loss, (vals, idx, value, mask)


def knn_points(p1, p2, K, threads=1):
    """pytorch3d.ops.knn_points restatement -- PARITY UNPINNED (see hitgeom_oracle.c header)."""
    p1, p2 = _c(p1, np.float32), _c(p2, np.float32)
    B, N, _ = p1.shape
    M = p2.shape[1]
    d, idx = np.empty((B, N, K), np.float32), np.empty((B, N, K), np.int64)

    def run(b0, b1):
        rc = lib().orc_knn_points(_p(p1[b0:b1], _f), _p(p2[b0:b1], _f), b1 - b0, N, M, K, _p(d[b0:b1], _f), _p(idx[b0:b1], _l))
        if rc:
            raise ValueError("K > number of points")

    _par(B, threads, run)
    return d, idx


# ------------------------------------------------------------------------------------------------------------
# model/pointnet2_utils.py (torch-level seams)
# ------------------------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """model/pointnet2_utils.py:19-40."""
    src, dst = _c(src, np.float32), _c(dst, np.float32)
    B, N, C = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().orc_square_distance(_p(src, _f), _p(dst, _f), B, N, M, C, _p(out, _f))
    return out


def fps_torch(xyz, npoint, start):
    """model/pointnet2_utils.py:63-84 with the `torch.randint` start indices (:75) passed in."""
    xyz, start = _c(xyz, np.float32), _c(start, np.int64)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), np.int64)
    lib().orc_fps_torch(_p(xyz, _f), B, N, npoint, _p(start, _l), _p(out, _l))
    return out


def query_ball_torch(radius, nsample, xyz, new_xyz):
    """model/pointnet2_utils.py:87-107; the threshold is float32(radius**2 evaluated in double)."""
    xyz, new_xyz = _c(xyz, np.float32), _c(new_xyz, np.float32)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    lib().orc_query_ball_torch(ctypes.c_float(np.float32(radius ** 2)), nsample, _p(xyz, _f), _p(new_xyz, _f), B, N, S, _p(out, _l))
    return out


# ------------------------------------------------------------------------------------------------------------
# pointnet2_ops (_ext-src/src/*.cu)
# ------------------------------------------------------------------------------------------------------------
def p2_fps(xyz, m):
    """sampling_gpu.cu:69-173 -> idx [B,m] int32."""
    xyz = _c(xyz, np.float32)
    B, n, _ = xyz.shape
    out = np.zeros((B, m), np.int32)
    lib().orc_p2_fps(_p(xyz, _f), B, n, m, _p(out, _i))
    return out


def p2_gather(points, idx):
    points, idx = _c(points, np.float32), _c(idx, np.int32)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().orc_p2_gather(_p(points, _f), _p(idx, _i), b, c, n, m, _p(out, _f))
    return out


def p2_gather_grad(grad_out, idx, n):
    grad_out, idx = _c(grad_out, np.float32), _c(idx, np.int32)
    b, c, m = grad_out.shape
    out = np.empty((b, c, n), np.float32)
    lib().orc_p2_gather_grad(_p(grad_out, _f), _p(idx, _i), b, c, n, m, _p(out, _f))
    return out


def p2_ball_query(new_xyz, xyz, radius, nsample):
    """ball_query_gpu.cu:9-44 (argument order of the `_ext` function)."""
    new_xyz, xyz = _c(new_xyz, np.float32), _c(xyz, np.float32)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = np.empty((b, m, nsample), np.int32)
    lib().orc_p2_ball_query(_p(new_xyz, _f), _p(xyz, _f), b, n, m, ctypes.c_float(radius), nsample, _p(out, _i))
    return out


def p2_group(points, idx):
    points, idx = _c(points, np.float32), _c(idx, np.int32)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = np.empty((b, c, npoints, nsample), np.float32)
    lib().orc_p2_group(_p(points, _f), _p(idx, _i), b, c, n, npoints, nsample, _p(out, _f))
    return out


def p2_group_grad(grad_out, idx, n):
    grad_out, idx = _c(grad_out, np.float32), _c(idx, np.int32)
    b, c, npoints, nsample = grad_out.shape
    out = np.empty((b, c, n), np.float32)
    lib().orc_p2_group_grad(_p(grad_out, _f), _p(idx, _i), b, c, n, npoints, nsample, _p(out, _f))
    return out


def p2_three_nn(unknown, known):
    unknown, known = _c(unknown, np.float32), _c(known, np.float32)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d, idx = np.empty((b, n, 3), np.float32), np.empty((b, n, 3), np.int32)
    with np.errstate(over="ignore"):
        lib().orc_p2_three_nn(_p(unknown, _f), _p(known, _f), b, n, m, _p(d, _f), _p(idx, _i))
    return d, idx


def p2_three_interpolate(points, idx, weight):
    points, idx, weight = _c(points, np.float32), _c(idx, np.int32), _c(weight, np.float32)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), np.float32)
    lib().orc_p2_three_interpolate(_p(points, _f), _p(idx, _i), _p(weight, _f), b, c, m, n, _p(out, _f))
    return out


def p2_three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _c(grad_out, np.float32), _c(idx, np.int32), _c(weight, np.float32)
    b, c, n = grad_out.shape
    out = np.empty((b, c, m), np.float32)
    lib().orc_p2_three_interpolate_grad(_p(grad_out, _f), _p(idx, _i), _p(weight, _f), b, c, n, m, _p(out, _f))
    return out
