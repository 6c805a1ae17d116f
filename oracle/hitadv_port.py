"""CPU restatement of the reference's HiT-ADV attack loop (ShapeAttack/HiT_ADV.py:44-287 and its helpers
:298-346, :489-559), written from the algorithm: the same tensor program the reference runs (repeat-based kernel
density, a `central_num`-step Python blend loop, per-iteration host bookkeeping, Adam inside a binary search),
so that (a) tests can check the B200-native loop (hit-adv_b200/hitgeom/hit_adv.py) against it on the same seeds
and (b) bench.py can time what the reference's path costs on the host cores.

TEST INFRASTRUCTURE ONLY.  Pinned against tests/golden/hitadv_ref.npz, which the UNMODIFIED reference class
produced on this container's CPU (tests/golden/make_golden_hitadv.py).  `pytorch3d.ops.knn_points` is third
party and absent: both the golden run and this port use the oracle's documented-semantics restatement.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O


def knn_points_idx(p1, p2, K):
    """[B,N,3] x [B,M,3] -> idx [B,N,K] int64 (K nearest of p2 for each p1 point, ascending distance)."""
    _, idx = O.knn_points(p1.detach().cpu().numpy(), p2.detach().cpu().numpy(), K, threads=O.host_threads())
    return torch.from_numpy(idx).to(p1.device)


def gather_points(x, idx):
    """x [B,M,U], idx [B,L,K] -> [B,L,K,U]."""
    B, L, K = idx.shape
    return torch.gather(x, 1, idx.reshape(B, L * K, 1).expand(-1, -1, x.shape[-1])).view(B, L, K, x.shape[-1])


def index_rows(points, idx):
    """points [B,N,C], idx [B,S] or [B,S,k] -> [B,S,C] / [B,S,k,C] (the reference's index_points)."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, points.shape[-1]))
    return out.view(*idx.shape, points.shape[-1])


def unit(v, dim=1, eps=1e-12):
    return v / v.norm(2, dim, keepdim=True).clamp(min=eps).expand_as(v)


def kappa(pc, normal, k):
    """HiT_ADV.py:318-325: mean |<unit(neighbour - p), n_p>| over the k nearest neighbours.  pc, normal [B,3,N]."""
    pts = pc.permute(0, 2, 1).contiguous()
    idx = knn_points_idx(pts, pts, k + 1)
    nn_pts = gather_points(pts, idx).permute(0, 3, 1, 2)[:, :, :, 1:].contiguous()  # [B,3,N,k]
    vec = unit(nn_pts - pc.unsqueeze(3))
    return torch.abs((vec * normal.unsqueeze(3)).sum(1)).mean(2), idx


def kappa_std(pc, normal, k):
    """HiT_ADV.py:327-339: std over the k neighbours' kappa."""
    kap, idx = kappa(pc, normal, k)
    nn_kappa = gather_points(kap.unsqueeze(2), idx).permute(0, 3, 1, 2)[:, :, :, 1:].contiguous()
    return torch.std(nn_kappa.squeeze(1), dim=2)


def fps(xyz, npoint):
    """HiT_ADV.py:489-510 (random start drawn from the global CPU generator)."""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    distance = torch.ones(B, N, device=xyz.device) * 1e10
    farthest = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    rows = torch.arange(B, device=xyz.device)
    for i in range(npoint):
        centroids[:, i] = farthest
        c = xyz[rows, farthest, :].view(B, 1, 3)
        d = torch.sum((xyz - c) ** 2, -1)
        m = d < distance
        distance[m] = d[m]
        farthest = torch.max(distance, -1)[1]
    return centroids


def kernel_density(central_points, pc, delta):
    """HiT_ADV.py:298-304: exp(-||x - c|| / (2 delta^2)) -> [B, J, N]."""
    J, N = central_points.shape[2], pc.shape[2]
    a = pc.unsqueeze(3).repeat(1, 1, 1, J)
    c = central_points.unsqueeze(2).repeat(1, 1, N, 1)
    norm = torch.norm(a - c, dim=1)
    return torch.exp(-norm / (2 * delta * delta).unsqueeze(1)).transpose(1, 2).contiguous()


def deform(adv, perturb, density):
    """HiT_ADV.py:162-175: the J-step weighted blend."""
    B, _, K = adv.shape
    acc = torch.zeros_like(adv)
    den = torch.zeros(B, 1, K, device=adv.device)
    for j in range(perturb.shape[1]):
        acc = acc + (adv + perturb[:, j, :].unsqueeze(2)) * density[:, j, :].unsqueeze(1)
        den = den + density[:, j, :].unsqueeze(1)
    return acc / den


def transformation_loss(perturb, delta, J, batch_avg=True):
    if batch_avg:
        return (torch.norm(perturb) + torch.norm(1 - delta)) / J
    return (torch.norm(perturb, dim=(1, 2)) + torch.norm(1 - delta, dim=1)) / J


def curv_std_loss(delta, central_kappa_std, max_delta, min_delta):
    ns = (central_kappa_std - central_kappa_std.min()) / (central_kappa_std.max() - central_kappa_std.min() + 1e-7)
    nd = (delta - min_delta) / (max_delta - min_delta + 1e-7)
    return F.cosine_similarity(ns.squeeze(-1), nd)


def chamfer_channel_first(adv, ori, weights):
    """`ChamferDist()(tmp_adv, ori, weights)` on [B,3,K] inputs: the 3x3 degenerate case of SURVEY.md R3."""
    from . import torch_port as tp

    return tp.chamfer_dist(adv, ori, "adv2ori", weights, batch_avg=True)


def untargeted_logits_loss(logits, targets, kappa_margin):
    """util/adv_utils.py:38-67."""
    one_hot = torch.zeros_like(logits).scatter_(1, targets.view(-1, 1), 1.0)
    real = torch.sum(one_hot * logits, dim=1)
    other = torch.max((1.0 - one_hot) * logits - one_hot * 10000.0, dim=1)[0]
    return torch.clamp(real - other + kappa_margin, min=0.0).mean()


def setup(model, ori, normal, target, hp):
    """HiT_ADV.py:61-97,119-124: scores, candidate centres, the chosen `central_num` centres."""
    B = ori.shape[0]
    k = hp["curv_loss_knn"]
    ks = kappa_std(ori, normal, k)
    x = ori.clone().requires_grad_()
    logits = model(x)
    logits = logits[0] if isinstance(logits, tuple) else logits
    F.cross_entropy(logits, target).backward()
    grad = x.grad.detach()
    with torch.no_grad():
        center = torch.median(ori, dim=-1)[0]
        diff = ori - center[:, :, None]
        r = torch.sum(diff ** 2, dim=1) ** 0.5
        sal = -1.0 * (r ** hp["alpha"]) * torch.sum(diff * grad, dim=1)
        nsal = (sal - sal.min()) / (sal.max() - sal.min() + 1e-7)
        nstd = (ks - ks.min()) / (ks.max() - ks.min() + 1e-7)
        score = 0.001 * nsal + nstd
        pts = ori.transpose(1, 2).contiguous()
        far_idx = fps(pts, hp["total_central_num"])
        far = index_rows(pts, far_idx)
        knn_idx = knn_points_idx(far, pts, k + 1)  # [B,T,k+1]
        far_pts = gather_points(pts, knn_idx)  # [B,T,k+1,3]
        far_score = index_rows(score.unsqueeze(2), knn_idx)  # [B,T,k+1,1]
        pick = far_score.topk(k=1, dim=2)[1].squeeze(-1)  # [B,T,1]
        total_pts = index_rows(far_pts.reshape(-1, k + 1, 3), pick.view(-1, 1)).view(B, -1, 3)  # [B,T,3]
        total_score = index_rows(far_score.view(-1, k + 1, 1), pick.view(-1, 1)).view(B, -1)
        _, sel = torch.topk(total_score, k=hp["central_num"])
        central = index_rows(total_pts, sel).transpose(1, 2).contiguous()  # [B,3,J]
        kap, _ = kappa(ori, normal, k)
        far_kap = index_rows(kap.unsqueeze(2), knn_idx)
        total_kap = index_rows(far_kap.view(-1, k + 1, 1), pick.view(-1, 1)).view(B, -1, 1)
        central_kappa_std = index_rows(total_kap, sel)  # [B,J,1]
    return score, central, central_kappa_std


def attack(model, data, target, hp):
    """Returns (best adversarial clouds [B,K,3] float64 numpy, success count, iterations run).  hp: dict with the
    reference's constructor arguments (HiT_ADV.py:18-22) plus `kappa` for the untargeted logits loss."""
    dev = data.device
    B, K = data.shape[:2]
    J = hp["central_num"]
    ori = data[:, :, :3].float().clone().transpose(1, 2).contiguous()
    normal = data[:, :, 3:].float().clone().transpose(1, 2).contiguous()
    target = target.long()
    label = target.cpu().numpy()
    _, central, central_kappa_std = setup(model, ori, normal, target, hp)
    lower = torch.zeros(B)
    scale = torch.ones(B) * hp["init_weight"]
    upper = torch.ones(B) * hp["max_weight"]
    cd_w = torch.from_numpy(np.ones((B,)) * hp["cd_weight"])
    o_bestdist = np.array([1e10] * B)
    o_bestattack = np.zeros((B, 3, K))
    iters = 0
    for _ in range(hp["binary_step"]):
        adv = ori.clone()
        perturb = (torch.rand(B, J, 3) * torch.tensor(hp["budget"])).to(dev)
        delta = (torch.ones((B, J)).to(dev) * hp["min_sigm"] + torch.rand((B, J)).to(dev) * (hp["max_sigm"] - hp["min_sigm"]))
        perturb.requires_grad_()
        delta.requires_grad_()
        bestdist = np.array([1e10] * B)
        bestscore = np.array([-1] * B)
        opt = torch.optim.Adam([{"params": perturb, "lr": hp["attack_lr"] * 5}, {"params": delta, "lr": hp["attack_lr"] * 3}],
                               weight_decay=0.0)
        for _it in range(hp["num_iter"]):
            with torch.no_grad():
                perturb.data = torch.clamp(perturb.data, min=-hp["budget"], max=hp["budget"])
                delta.data = torch.clamp(delta.data, min=hp["min_sigm"], max=hp["max_sigm"])
            dens = kernel_density(central, ori, delta)
            tmp = deform(adv, perturb, dens)
            logits = model(tmp)
            logits = logits[0] if isinstance(logits, tuple) else logits
            pred = torch.argmax(logits, dim=1)
            dist_val = transformation_loss(perturb, delta, J, batch_avg=False).detach().cpu().numpy()
            pred_val = pred.detach().cpu().numpy()
            input_val = tmp.detach().cpu().numpy()
            for e in range(B):
                if dist_val[e] < bestdist[e] and pred_val[e] != label[e]:
                    bestdist[e] = dist_val[e]
                    bestscore[e] = pred_val[e]
                if dist_val[e] < o_bestdist[e] and pred_val[e] != label[e]:
                    o_bestdist[e] = dist_val[e]
                    o_bestattack[e] = input_val[e]
            adv_loss = untargeted_logits_loss(logits, target, hp["kappa"])
            dist_loss = torch.tensor(0.0, device=dev)
            if hp["cd_weight"] != 0:
                dist_loss = dist_loss + chamfer_channel_first(tmp, ori, cd_w.to(dev))
            if hp["ker_weight"] != 0:
                dist_loss = dist_loss + transformation_loss(perturb, delta, J) * hp["ker_weight"]
            if hp["hide_weight"] != 0:
                dist_loss = dist_loss + (curv_std_loss(delta, central_kappa_std, hp["max_sigm"], hp["min_sigm"]) * hp["hide_weight"]).mean()
            loss = adv_loss + scale.float().to(dev) * dist_loss
            opt.zero_grad()
            loss.mean().backward()
            opt.step()
            iters += 1
        for e in range(B):
            if bestscore[e] != label[e] and bestscore[e] != -1 and bestdist[e] <= o_bestdist[e]:
                lower[e] = max(lower[e], scale[e])
            else:
                upper[e] = min(upper[e], scale[e])
            scale[e] = (lower[e] + upper[e]) / 2.0
    for e in range(B):
        if lower[e] == 0.0:
            o_bestattack[e] = input_val[e]
            o_bestdist[e] = dist_val[e]
    return o_bestattack.transpose((0, 2, 1)), int((lower > 0.0).sum()), iters
