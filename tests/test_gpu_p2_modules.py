"""hitgeom.pointnet2_ops.pointnet2_modules / pointnet2_utils against the UNMODIFIED reference
Python layer run on CPU over an oracle-backed `_ext` (tests/golden/make_golden_modules.py).  The reference's state
dicts load into hitgeom's modules unchanged.  Index work is bit-exact (new_xyz compared exactly); features go through
1x1 convolutions (cuDNN/cuBLAS on the GPU vs MKL on the CPU): 1e-5 relative on outputs, 1e-4 norm-wise on gradients."""
import numpy as np
import pytest
import torch

from util_inputs import normwise

pytestmark = pytest.mark.gpu


def _load(module, g, tag):
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{tag}_sd_")}
    module.load_state_dict(sd, strict=True)
    return module.cuda().eval()


def _run(module, g, tag, inputs, grad_of):
    res = module(*inputs)
    res = res if isinstance(res, tuple) else (res,)
    y = res[-1]
    ref = g[f"{tag}_out"]
    assert np.abs(y.detach().cpu().numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1.0), tag
    if f"{tag}_new_xyz" in g.files:
        assert np.array_equal(res[0].cpu().numpy(), g[f"{tag}_new_xyz"]), tag
    (y * torch.from_numpy(g[f"{tag}_w"]).cuda()).sum().backward()
    assert normwise(grad_of.grad.cpu().numpy(), g[f"{tag}_grad"]) < 1e-4, tag
    grad_of.grad = None


def test_sa_msg_groupall_fp_modules_match_reference(golden):
    from hitgeom.pointnet2_ops import pointnet2_modules as pm

    g = golden("p2_modules_ref")
    xyz = torch.from_numpy(g["xyz"]).cuda()
    feats = torch.from_numpy(g["feats"]).cuda()
    C = feats.shape[1]
    f = feats.clone().requires_grad_()
    spec = [C, 16, 32]
    _run(_load(pm.PointnetSAModule(spec, npoint=32, radius=0.3, nsample=16), g, "sa"), g, "sa", (xyz, f), f)
    assert spec[0] == C + 3  # the reference's in-place `mlp_spec[0] += 3` (pointnet2_modules.py:115-116)
    _run(_load(pm.PointnetSAModuleMSG(32, [0.2, 0.4], [8, 16], [[C, 16], [C, 8, 24]]), g, "msg"), g, "msg", (xyz, f), f)
    m = _load(pm.PointnetSAModule([C, 32]), g, "all")
    new_xyz, _ = m(xyz, f)
    assert new_xyz is None
    _run(m, g, "all", (xyz, f), f)
    known = xyz[:, :32].contiguous()
    # the FP case needs the fixture's own known features: redraw them from the fixture's generator sequence
    gen = torch.Generator().manual_seed(12)
    torch.randn(2, C, 256, generator=gen)
    for tag in ("sa", "msg", "all"):
        torch.randn(g[f"{tag}_out"].shape, generator=gen)
    kf = torch.randn(2, 12, 32, generator=gen).cuda().requires_grad_()
    _run(_load(pm.PointnetFPModule([12 + C, 24]), g, "fp"), g, "fp", (xyz, known, feats, kf), kf)


def test_query_and_group_xyz_gradient_matches_reference(golden):
    from hitgeom.pointnet2_ops import pointnet2_utils as pu

    g = golden("p2_modules_ref")
    x = torch.from_numpy(g["xyz"]).cuda().requires_grad_()
    new_xyz = x.detach()[:, :16].contiguous()
    grouped = pu.QueryAndGroup(0.3, 8)(x, new_xyz)
    assert np.array_equal(grouped.detach().cpu().numpy(), g["qg_out"])
    (grouped * torch.from_numpy(g["qg_w"]).cuda()).sum().backward()
    assert normwise(x.grad.cpu().numpy(), g["qg_grad"]) < 1e-5


def test_fp_module_without_known_broadcasts():
    from hitgeom.pointnet2_ops import pointnet2_modules as pm

    m = pm.PointnetFPModule([6, 8]).cuda().eval()
    out = m(torch.randn(2, 50, 3).cuda(), None, None, torch.randn(2, 6, 1).cuda())
    assert out.shape == (2, 8, 50)


@pytest.mark.parametrize("C", [0, 5, 64])
def test_fused_group_concat_equals_the_composition(C):
    """QueryAndGroup's fused body (hg_p2_group_concat) == two grouping ops + centre subtraction + cat: forward bit for bit,
    gradients w.r.t. xyz / new_xyz / features as well (same deterministic ascending-edge sums)."""
    from hitgeom.pointnet2_ops import pointnet2_utils as pu
    from util_inputs import clouds

    B, N, S, ns = 3, 500, 40, 16
    xyz = torch.from_numpy(clouds(B, N, 61, "surface")).cuda()
    new_xyz = xyz[:, :S].contiguous() + 0.01
    feats = torch.randn(B, C, N, device="cuda") if C else None
    idx = pu.ball_query(0.25, ns, xyz, new_xyz)
    leaves = []
    outs = []
    for fused in (True, False):
        x, nx = xyz.clone().requires_grad_(), new_xyz.clone().requires_grad_()
        f = feats.clone().requires_grad_() if C else None
        if fused:
            out = pu.GroupConcat.apply(x, nx, f, idx)
        else:
            rel = pu.grouping_operation(x.transpose(1, 2).contiguous(), idx) - nx.transpose(1, 2).unsqueeze(-1)
            out = torch.cat([rel, pu.grouping_operation(f, idx)], dim=1) if C else rel
        outs.append(out)
        leaves.append((x, nx, f))
    assert torch.equal(outs[0], outs[1])
    w = torch.randn_like(outs[0])
    for out in outs:
        (out * w).sum().backward()
    (x0, n0, f0), (x1, n1, f1) = leaves
    assert torch.equal(x0.grad, x1.grad)
    assert normwise(n0.grad.cpu().numpy(), n1.grad.cpu().numpy()) < 1e-6  # (a reduction over the samples: order may differ)
    if C:
        assert torch.equal(f0.grad, f1.grad)
