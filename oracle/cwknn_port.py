"""Restatement of the reference's CW-kNN attack loops (CW/kNN.py:40-151, CW/UKNN.py:41-159) as one device-agnostic
function: same statements in the same order -- per-iteration `.item()` read-back, the two transposes per iteration,
`adv_data.data = clip(...)` -- minus the prints, timers and `.cuda()` calls.

TEST INFRASTRUCTURE ONLY (see oracle/hitgeom_oracle.c).  Pinned: with oracle/torch_port.py's ChamferkNNDist and the
reference-exact clip / loss functions it reproduces the UNMODIFIED reference classes bit for bit on this container's
CPU (tests/test_oracle_golden.py against tests/golden/cwknn_ref.npz).  Uses: that test, and bench.py's
`--workload cwknn` legs (the reference program on the host cores, and the same program on the GPU as the second bar).
"""
import torch
import torch.optim as optim


def attack(model, data, target, adv_func, dist_func, clip_func, attack_lr=1e-3, num_iter=2500, untargeted=False,
           pre_head=None, device=None):
    """-> (adv [B,K,3] float32 numpy, success count)."""
    device = torch.device(device or "cpu")
    model = model.to(device).eval()
    B, K = data.shape[:2]
    data = data.float().to(device).detach()
    data = data.transpose(1, 2).contiguous()
    ori_data = data.clone().detach()
    ori_data.requires_grad = False
    if ori_data.shape[1] == 3:
        normal = None
    else:
        normal = ori_data[:, 3:, :]
        ori_data = ori_data[:, :3, :]
    target = target.long().to(device).detach()
    adv_data = ori_data.clone().detach() + torch.randn((B, 3, K)).to(device) * 1e-7
    adv_data.requires_grad_()
    opt = optim.Adam([adv_data], lr=attack_lr, weight_decay=0.)

    def forward(x):
        logits = model(pre_head(x)) if pre_head is not None else model(x)
        return logits[0] if isinstance(logits, tuple) else logits

    for iteration in range(num_iter):
        logits = forward(adv_data)
        pred = torch.argmax(logits, dim=1)
        _ = ((pred != target) if untargeted else (pred == target)).sum().item()  # kNN.py:90 (a sync every iteration)
        adv_loss = adv_func(logits, target).mean()
        dist_loss = dist_func(adv_data.transpose(1, 2).contiguous(), ori_data.transpose(1, 2).contiguous()).mean() * K
        loss = adv_loss + dist_loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        if clip_func is not None:
            if untargeted:
                adv_data.data = clip_func(adv_data.clone().detach(), ori_data, normal)  # UKNN.py:121-122
            else:
                adv_data.data = clip_func(adv_data.clone().detach(), ori_data)
    with torch.no_grad():
        pred = torch.argmax(forward(adv_data), dim=-1)
        success_num = ((pred != target) if untargeted else (pred == target)).sum().detach().cpu().item()
    return adv_data.transpose(1, 2).contiguous().detach().cpu().numpy(), success_num
