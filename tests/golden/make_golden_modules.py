"""tests/golden/p2_modules_ref.npz: the UNMODIFIED reference `pointnet2_ops.pointnet2_utils` autograd wrappers and
`pointnet2_ops.pointnet2_modules` (PointnetSAModule / PointnetSAModuleMSG / PointnetFPModule) run on this container's
CPU.  Their native module `pointnet2_ops._ext` needs CUDA, so the harness registers a stand-in `_ext` backed by the
ORACLE's restatements of the nine kernels (pinned bit-exactly against the reference's own CUDA kernels on a B200:
tests/golden/pointnet2_ref.npz).  What this fixture pins is the reference's PYTHON layer: wrapper semantics, the
QueryAndGroup / GroupAll composition, module wiring, state-dict layout, gradients.  Build container only."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _refload  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util_inputs import clouds  # noqa: E402


def _np(t):
    return np.ascontiguousarray(t.detach().numpy())


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def oracle_ext():
    ext = types.ModuleType("pointnet2_ops._ext")
    ext.furthest_point_sampling = lambda xyz, m: _t(O.p2_fps(_np(xyz), m))
    ext.gather_points = lambda pts, idx: _t(O.p2_gather(_np(pts), _np(idx)))
    ext.gather_points_grad = lambda g, idx, n: _t(O.p2_gather_grad(_np(g), _np(idx), n))
    ext.ball_query = lambda new_xyz, xyz, r, ns: _t(O.p2_ball_query(_np(new_xyz), _np(xyz), np.float32(r), ns))
    ext.group_points = lambda pts, idx: _t(O.p2_group(_np(pts), _np(idx)))
    ext.group_points_grad = lambda g, idx, n: _t(O.p2_group_grad(_np(g), _np(idx), n))
    ext.three_nn = lambda u, k: tuple(_t(a) for a in O.p2_three_nn(_np(u), _np(k)))
    ext.three_interpolate = lambda pts, idx, w: _t(O.p2_three_interpolate(_np(pts), _np(idx), _np(w)))
    ext.three_interpolate_grad = lambda g, idx, w, m: _t(O.p2_three_interpolate_grad(_np(g), _np(idx), _np(w), m))
    return ext


def main():
    _refload.install_stubs()
    pkg = types.ModuleType("pointnet2_ops")
    pkg.__path__ = []
    pkg._ext = oracle_ext()
    sys.modules["pointnet2_ops"] = pkg
    sys.modules["pointnet2_ops._ext"] = pkg._ext
    base = "pointnet2_ops_lib/pointnet2_ops/"
    pu = _refload.by_path("pointnet2_ops.pointnet2_utils", base + "pointnet2_utils.py")
    sys.modules["pointnet2_ops.pointnet2_utils"] = pu
    pkg.pointnet2_utils = pu
    pm = _refload.by_path("pointnet2_ops.pointnet2_modules", base + "pointnet2_modules.py")

    B, N, C = 2, 256, 8
    xyz = _t(clouds(B, N, 77, "surface"))
    g = torch.Generator().manual_seed(12)
    feats = torch.randn(B, C, N, generator=g)
    out = dict(xyz=xyz.numpy(), feats=feats.numpy())

    def run(tag, module, *inputs, grad_of):
        module.eval()
        res = module(*inputs)
        res = res if isinstance(res, tuple) else (res,)
        y = res[-1]
        w = torch.randn(y.shape, generator=g)
        (y * w).sum().backward()
        for k, v in module.state_dict().items():
            out[f"{tag}_sd_{k}"] = v.numpy()
        out[f"{tag}_w"] = w.numpy()
        out[f"{tag}_out"] = y.detach().numpy()
        if len(res) > 1 and res[0] is not None:
            out[f"{tag}_new_xyz"] = res[0].detach().numpy()
        out[f"{tag}_grad"] = grad_of.grad.numpy().copy()
        grad_of.grad = None

    torch.manual_seed(5)
    f = feats.clone().requires_grad_()
    run("sa", pm.PointnetSAModule([C, 16, 32], npoint=32, radius=0.3, nsample=16), xyz, f, grad_of=f)
    run("msg", pm.PointnetSAModuleMSG(32, [0.2, 0.4], [8, 16], [[C, 16], [C, 8, 24]]), xyz, f, grad_of=f)
    run("all", pm.PointnetSAModule([C, 32]), xyz, f, grad_of=f)  # npoint=None -> GroupAll
    known = xyz[:, :32].contiguous()
    kf = torch.randn(B, 12, 32, generator=g).requires_grad_()
    run("fp", pm.PointnetFPModule([12 + C, 24]), xyz, known, feats, kf, grad_of=kf)
    # wrappers on their own: xyz-only grouping and its gradient w.r.t. the coordinates
    x = xyz.clone().requires_grad_()
    new_xyz = xyz[:, :16].contiguous()
    grouped = pu.QueryAndGroup(0.3, 8)(x, new_xyz)
    wq = torch.randn(grouped.shape, generator=g)
    (grouped * wq).sum().backward()
    out.update(qg_out=grouped.detach().numpy(), qg_w=wq.numpy(), qg_grad=x.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "p2_modules_ref.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
