"""Stand-in for the two `pytorch3d.ops` functions the reference imports (`knn_points`, `knn_gather`;
util/dist_utils.py:12, ShapeAttack/HiT_ADV.py:9, FGM/GeoA3_args.py:14; pinned pytorch3d==0.7.2).

pytorch3d is a third-party dependency whose source is not part of the reference tree, and no reference test
pins its results: PARITY UNPINNED.  This restates the documented semantics: squared L2 distances from
direct coordinate differences, the K smallest in ascending order, int64 indices, `knn_gather(x, idx)[n,l,k]
= x[n, idx[n,l,k]]`.  `lengths1/lengths2`, `norm != 2` and `version` are not supported (no caller uses them).
"""
from collections import namedtuple

import torch

from . import functional as F

_KNN = namedtuple("KNN", "dists idx knn")


def knn_gather(x, idx, lengths=None):
    """x [N,M,U], idx [N,L,K] -> [N,L,K,U]."""
    if lengths is not None:
        raise NotImplementedError("knn_gather: lengths is not supported")
    N, L, K = idx.shape
    from .model_seams import index_points

    return index_points(x, idx.reshape(N, L * K)).view(N, L, K, x.shape[-1])


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    if lengths1 is not None or lengths2 is not None or norm != 2:
        raise NotImplementedError("knn_points: lengths / norm != 2 are not supported")
    if int(K) > 64:
        raise NotImplementedError(f"knn_points: K={K} > 64 neighbours is not supported by the streaming kernel "
                                  "(register-resident lists; see INTEGRATION.md)")
    dists, idx = F.knn_points_raw(p1.detach().contiguous(), p2.detach().contiguous(), int(K))
    if p1.requires_grad or p2.requires_grad:
        # differentiable distances, recomputed from the gathered neighbours (pytorch3d's backward is the same
        # 2*(p1 - p2[idx]) gather/scatter); indices stay non-differentiable
        nn_pts = knn_gather(p2, idx)
        dists = ((p1[:, :, None, :] - nn_pts) ** 2).sum(-1)
    nn = knn_gather(p2, idx) if return_nn else None
    return _KNN(dists=dists, idx=idx, knn=nn)
