"""Generate tests/golden/pointnet2_ref.npz ON A B200 by running the reference's own pointnet2_ops CUDA kernels
(compiled unmodified for sm_100 into oracle/_ref by oracle/build_ref.py) on seeded inputs:

    gpurun -- python tests/golden/make_golden_gpu.py        # writes gpurun_out/pointnet2_ref.npz
    cp gpurun_out/pointnet2_ref.npz tests/golden/           # commit it

These fixtures pin the oracle's restatement of the nine kernels (tests/test_oracle_golden.py), which the build
container cannot execute.  Only index / gather outputs are stored (the *_grad kernels use float atomics and are
not bit-reproducible by design)."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util_inputs import clouds  # noqa: E402


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "_ext_ref.so")
    spec = importlib.util.spec_from_file_location("_ext_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    out = {}
    for tag, (B, N, m, kind) in {"a": (3, 1024, 64, "surface"), "b": (2, 600, 600, "gauss"), "c": (2, 100, 17, "gauss")}.items():
        xyz = clouds(B, N, 400 + N, kind)
        if tag == "a":
            xyz[:, 512:] = xyz[:, :512]  # exact duplicates: ties in the running distance
            xyz[:, 7] = 0.0
        out[f"fps_{tag}_xyz"] = xyz
        out[f"fps_{tag}_idx"] = ref.furthest_point_sampling(g(xyz), m).cpu().numpy()
    xyz = clouds(3, 1024, 77, "surface")
    fps = ref.furthest_point_sampling(g(xyz), 51)
    new_xyz = ref.gather_points(g(xyz.transpose(0, 2, 1)), fps).transpose(1, 2).contiguous()
    out.update(bq_xyz=xyz, bq_new_xyz=new_xyz.cpu().numpy(), bq_fps=fps.cpu().numpy())
    for r, ns in [(0.126, 16), (0.2, 32), (0.4, 64), (0.219, 49), (0.01, 8)]:
        out[f"bq_r{r}_ns{ns}"] = ref.ball_query(new_xyz, g(xyz), r, ns).cpu().numpy()
    idx = ref.ball_query(new_xyz, g(xyz), 0.2, 32)
    pts = np.random.default_rng(0).standard_normal((3, 5, 1024)).astype(np.float32)
    out.update(grp_points=pts, grp_out=ref.group_points(g(pts), idx).cpu().numpy())
    unknown, known = clouds(2, 700, 5, "surface"), clouds(2, 300, 6, "surface")
    d2, i3 = ref.three_nn(g(unknown), g(known))
    w = np.random.default_rng(1).random((2, 700, 3)).astype(np.float32)
    feats = np.random.default_rng(2).standard_normal((2, 6, 300)).astype(np.float32)
    out.update(nn_unknown=unknown, nn_known=known, nn_dist2=d2.cpu().numpy(), nn_idx=i3.cpu().numpy(), ti_w=w,
               ti_feats=feats, ti_out=ref.three_interpolate(g(feats), i3, g(w)).cpu().numpy())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "pointnet2_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
