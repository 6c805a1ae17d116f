"""CPU restatement of the reference's *torch* path for the loss classes -- the same tensor program the reference
executes (three `bmm`, diagonal gather, two `min`; `matmul` + `topk`; autograd for the backward), written from
the formulas in util/set_distance.py:15-70 and util/dist_utils.py:56-80,136-175,279-294.

TEST INFRASTRUCTURE ONLY (see oracle/hitgeom_oracle.c).  Two uses:
  * bench.py's `cpu_baseline` and `--impl reference` legs time THIS on the GPU box's host cores: it is what
    the reference's CPU path costs (materialised [B,N,N] matrices, autograd-saved copies and all), which the
    matrix-free C oracle would understate;
  * tests/test_oracle_golden.py checks it against the golden vectors, so the timed program is the pinned one.
Pinned against tests/golden/loss_classes.npz and setdist_*.npz (bit-exact on this container's torch 2.11/MKL).
"""
import torch


def pairwise(x, y):
    """P[b,i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j with the squared norms read off the Gram diagonals."""
    xx = torch.bmm(x, x.transpose(2, 1))
    yy = torch.bmm(y, y.transpose(2, 1))
    zz = torch.bmm(x, y.transpose(2, 1))
    ix = torch.arange(x.shape[1], device=x.device)
    iy = torch.arange(y.shape[1], device=y.device)
    rx = xx[:, ix, ix].unsqueeze(1).expand_as(zz.transpose(2, 1))
    ry = yy[:, iy, iy].unsqueeze(1).expand_as(zz)
    return rx.transpose(2, 1) + ry - 2 * zz


def chamfer(preds, gts):
    P = pairwise(gts, preds)
    return torch.min(P, 1)[0].mean(dim=1), torch.min(P, 2)[0].mean(dim=1)


def hausdorff(preds, gts):
    P = pairwise(gts, preds)
    return torch.min(P, 1)[0].max(dim=1)[0], torch.min(P, 2)[0].max(dim=1)[0]


def _pick(l1, l2, method):
    return l1 if method == "adv2ori" else l2 if method == "ori2adv" else (l1 + l2) / 2.0


def _weighted(loss, weights, batch_avg):
    if weights is None:
        weights = torch.ones(loss.shape[0], device=loss.device)
    loss = loss * weights.float().to(loss.device)
    return loss.mean() if batch_avg else loss


def chamfer_dist(adv, ori, method="adv2ori", weights=None, batch_avg=True):
    return _weighted(_pick(*chamfer(adv, ori), method), weights, batch_avg)


def hausdorff_dist(adv, ori, method="adv2ori", weights=None, batch_avg=True):
    return _weighted(_pick(*hausdorff(adv, ori), method), weights, batch_avg)


def knn_dist(pc, k=5, alpha=1.05, weights=None, batch_avg=True):
    if pc.shape[1] != 3:
        pc = pc.transpose(2, 1)
    inner = -2.0 * torch.matmul(pc.transpose(2, 1), pc)
    xx = torch.sum(pc ** 2, dim=1, keepdim=True)
    dist = xx + inner + xx.transpose(2, 1)
    neg_value, _ = (-dist).topk(k=k + 1, dim=-1)
    value = torch.mean(-(neg_value[..., 1:]), dim=-1)
    with torch.no_grad():
        threshold = torch.mean(value, dim=-1) + alpha * torch.std(value, dim=-1)
        mask = (value > threshold[:, None]).float()
    return _weighted(torch.mean(value * mask, dim=1), weights, batch_avg)


def chamfer_knn_dist(adv, ori, weights=None, batch_avg=True, chamfer_method="adv2ori", knn_k=5, knn_alpha=1.05,
                     chamfer_weight=5.0, knn_weight=3.0):
    return (chamfer_dist(adv, ori, chamfer_method, weights, batch_avg) * chamfer_weight
            + knn_dist(adv, knn_k, knn_alpha, weights, batch_avg) * knn_weight)


def step_chamfer_knn(adv, ori):
    """One fwd+bwd of the CW-kNN distance term (CW/kNN.py:104-108): returns (loss, d loss / d adv)."""
    a = adv.detach().clone().requires_grad_()
    loss = chamfer_knn_dist(a, ori, batch_avg=False).sum()
    loss.backward()
    return loss.detach(), a.grad


def step_cd_hd_knn(adv, ori):
    """Config 1: ChamferDist + HausdorffDist + KNNDist(k=5) fwd+bwd."""
    a = adv.detach().clone().requires_grad_()
    loss = (chamfer_dist(a, ori, batch_avg=False) + hausdorff_dist(a, ori, batch_avg=False)
            + knn_dist(a, batch_avg=False)).sum()
    loss.backward()
    return loss.detach(), a.grad


def get_graph_feature(x, idx):
    """model/dgcnn_cls.py:16-43 after the neighbour search: x [B,C,N], idx [B,N,k] -> [B,2C,N,k] = cat(x[idx] - x, x),
    the same tensor program (flat gather on the transposed copy, repeat, cat, permute) on whatever device x is on.
    Pinned against tests/golden/dgcnn_edge.npz (bit-exact, forward and autograd gradient, on this container's CPU)."""
    B, C, N = x.shape
    k = idx.shape[2]
    flat = (idx + torch.arange(0, B, device=x.device).view(-1, 1, 1) * N).view(-1)
    xt = x.transpose(2, 1).contiguous()
    feature = xt.view(B * N, -1)[flat, :].view(B, N, k, C)
    xr = xt.view(B, N, 1, C).repeat(1, 1, k, 1)
    return torch.cat((feature - xr, xr), dim=3).permute(0, 3, 1, 2).contiguous()


# ---- the PointNet++ SSG victim's torch-level geometry (model/pointnet2_utils.py:19-107), same tensor programs ---------
def square_distance(src, dst):
    """:19-40  -2 src.dst^T, then += |src|^2, then += |dst|^2 (in place, in this order)."""
    d = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    d += torch.sum(src ** 2, -1).unsqueeze(2)
    d += torch.sum(dst ** 2, -1).unsqueeze(1)
    return d


def index_points(points, idx):
    """:43-60  batched advanced-index gather."""
    B = points.shape[0]
    shape = [B] + [1] * (idx.dim() - 1)
    batch = torch.arange(B, dtype=torch.long, device=points.device).view(shape).expand_as(idx)
    return points[batch, idx, :]


def farthest_point_sample(xyz, npoint, start):
    """:63-84 with the `torch.randint` start handed in: npoint rounds of gather, squared distance, masked min, argmax."""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    distance = torch.ones(B, N, device=xyz.device) * 1e10
    farthest = start.clone()
    batch = torch.arange(B, dtype=torch.long, device=xyz.device)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[batch, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = torch.max(distance, -1)[1]
    return centroids


def query_ball_point(radius, nsample, xyz, new_xyz):
    """:87-107  mark > r^2 as N, SORT the [B,S,N] index tensor, keep the first nsample, pad with the first."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    group_idx = torch.arange(N, dtype=torch.long, device=xyz.device).view(1, 1, N).repeat([B, S, 1])
    group_idx[square_distance(new_xyz, xyz) > radius ** 2] = N
    group_idx = group_idx.sort(dim=-1)[0][:, :, :nsample]
    first = group_idx[:, :, 0].view(B, S, 1).repeat([1, 1, nsample])
    mask = group_idx == N
    group_idx[mask] = first[mask]
    return group_idx
