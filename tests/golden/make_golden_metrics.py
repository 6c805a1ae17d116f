"""tests/golden/metrics_ref.npz: the UNMODIFIED reference evaluation metrics -- `uniform_loss`, `kNN_smoothing_loss`
(FGM/GeoA3_args.py:240-302) and `CurvStdDist` (util/dist_utils.py:464-495) -- run on this container's CPU.
Their native dependencies are absent here (pointnet2_ops needs CUDA, pytorch3d is not installed), so the harness
registers stand-ins backed by the ORACLE's restatements (oracle/hitgeom_oracle.c: the pointnet2_ops kernels are
pinned bit-exactly against outputs of the reference's own CUDA kernels on a B200, tests/golden/pointnet2_ref.npz;
knn_points follows pytorch3d's documented semantics).  What this fixture pins is therefore the reference's PYTHON
composition around those ops: shapes, permutes, constants, reductions.  Build container only."""
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _refload  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util_inputs import clouds, jitter  # noqa: E402


def _np(t):
    return np.ascontiguousarray(t.detach().numpy())


def main():
    _refload.install_stubs()
    KNN = namedtuple("KNN", "dists idx knn")

    def knn_points(p1, p2, K=1, **kw):
        d, i = O.knn_points(_np(p1), _np(p2), K)
        return KNN(torch.from_numpy(d), torch.from_numpy(i), None)

    def knn_gather(x, idx):
        B, L, K = idx.shape
        return torch.gather(x, 1, idx.reshape(B, L * K, 1).expand(-1, -1, x.shape[-1])).view(B, L, K, x.shape[-1])

    sys.modules["pytorch3d.ops"].knn_points = knn_points
    sys.modules["pytorch3d.ops"].knn_gather = knn_gather
    pu = types.ModuleType("pointnet2_ops_lib.pointnet2_ops.pointnet2_utils")
    pu.furthest_point_sample = lambda xyz, m: torch.from_numpy(O.p2_fps(_np(xyz), m))
    pu.gather_operation = lambda feats, idx: torch.from_numpy(O.p2_gather(_np(feats), _np(idx)))
    pu.ball_query = lambda r, ns, xyz, new_xyz: torch.from_numpy(O.p2_ball_query(_np(new_xyz), _np(xyz), np.float32(r), ns))
    pu.grouping_operation = lambda feats, idx: torch.from_numpy(O.p2_group(_np(feats), _np(idx)))
    lib = types.ModuleType("pointnet2_ops_lib")
    ops = types.ModuleType("pointnet2_ops_lib.pointnet2_ops")
    lib.pointnet2_ops, ops.pointnet2_utils = ops, pu
    sys.modules.update({"pointnet2_ops_lib": lib, "pointnet2_ops_lib.pointnet2_ops": ops,
                        "pointnet2_ops_lib.pointnet2_ops.pointnet2_utils": pu})
    sys.modules.setdefault("scipy.io", __import__("scipy.io"))
    geo = _refload.by_path("ref_geoa3_args", "FGM/GeoA3_args.py")
    import importlib

    du = importlib.import_module("util.dist_utils")

    B, n = 3, 1024
    ori = clouds(B, n, 321, "surface")
    adv = jitter(ori, 5)
    rng = np.random.default_rng(3)
    nrm = rng.standard_normal((B, n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    ori_t, adv_t, nrm_t = (torch.from_numpy(a).transpose(1, 2).contiguous() for a in (ori, adv, nrm))
    out = dict(ori=ori, adv=adv, nrm=nrm)
    out["uniform_k2"] = geo.uniform_loss(adv_t, k=2).numpy()
    out["uniform_k4_point_major"] = geo.uniform_loss(adv_t.transpose(1, 2).contiguous(), k=4).numpy()
    out["knn_smoothing_k5"] = geo.kNN_smoothing_loss(adv_t, 5).numpy()
    out["curvstd_k4"] = du.CurvStdDist(k=4)(ori_t, adv_t, nrm_t).numpy()
    out["kappa_std_k4"] = du.CurvStdDist(k=4)._get_kappa_std_ori(adv_t, nrm_t, k=4).numpy()
    np.savez_compressed(os.path.join(HERE, "metrics_ref.npz"), **out)
    print({k: (v.tolist() if v.size < 4 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
