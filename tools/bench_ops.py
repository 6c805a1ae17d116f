"""Per-operator numbers for the pointnet2_ops / seam kernels: device time, HBM GB/s against
the measured copy bandwidth (gather / group are HBM-bound), rounds/s for FPS -- next to the reference's OWN CUDA
kernels (oracle/_ref/_ext_ref.so: the unmodified sources compiled for sm_100) timed on the same GPU, same inputs.
Results: gpurun_out/bench_ops.json (summarised under profiles/)."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom import functional as F  # noqa: E402
from hitgeom import model_seams as ms  # noqa: E402
from hitgeom.pointnet2_ops import _ext  # noqa: E402


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "_ext_ref.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("_ext_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def timeit(fn, iters=20, flush=None):
    """Mean device time of fn: all iterations enqueued back to back (the GPU stays at its load clocks), one event
    pair per iteration on the launching stream, a single synchronize at the end; optional L2 flush between."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush[0].zero_()  # write 256 MiB (evicts everything) ...
            flush[1].sum()    # ... then read another 256 MiB: the dirty lines of the write are out before the timed op
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / iters


def main():
    ref = load_ref()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    flush = (torch.empty(256 << 20, dtype=torch.uint8, device="cuda"), torch.zeros(64 << 20, dtype=torch.float32, device="cuda"))
    res = []

    def row(name, ms_new, ms_ref, bytes_alg=None, extra=None):
        r = {"op": name, "ms": ms_new, "ref_ms": ms_ref, "speedup_vs_ref_cuda": (ms_ref / ms_new) if ms_ref else None}
        if bytes_alg:
            r["alg_bytes"] = bytes_alg
            r["gbs"] = bytes_alg / ms_new / 1e6
            r["hbm_frac_of_measured"] = r["gbs"] / hbm
        if extra:
            r.update(extra)
        res.append(r)
        print(r, flush=True)

    torch.manual_seed(0)
    # ---- the distance losses against the reference's OWN torch path on the SAME GPU ("the real
    # bar for C1/C4"): cuBLAS bmm + min / topk + autograd, util/set_distance.py:15-70, util/dist_utils.py:136-175 --------
    from hitgeom.dist_utils import ChamferDist, ChamferkNNDist, HausdorffDist, KNNDist
    from oracle import torch_port as tpl

    torch.backends.cuda.matmul.allow_tf32 = False  # the reference (torch 1.11 defaults) multiplies in FP32
    for (Bl, Nl) in ((388, 1024), (32, 1024)):
        ori = torch.randn(Bl, Nl, 3, device="cuda")
        ori = ori / ori.norm(dim=-1).amax(dim=1)[:, None, None]
        adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_()
        mods = {"ChamferDist": (ChamferDist(), lambda a, o: tpl.chamfer_dist(a, o), 1.0),
                "HausdorffDist": (HausdorffDist(), lambda a, o: tpl.hausdorff_dist(a, o), 1.0),
                "KNNDist(k=5)": (KNNDist(k=5), lambda a, o: tpl.knn_dist(a), 1.0),
                "ChamferkNNDist": (ChamferkNNDist(), lambda a, o: tpl.chamfer_knn_dist(a, o), 2.0)}
        for name, (mod, ref_fn, npass) in mods.items():
            def fb_new():
                adv.grad = None
                (mod(adv) if name.startswith("KNN") else mod(adv, ori)).backward()

            def fb_ref():
                adv.grad = None
                ref_fn(adv, ori).backward()

            t_new, t_ref = timeit(fb_new, flush=flush), timeit(fb_ref, iters=5, flush=flush)
            row(f"{name} fwd+bwd B={Bl} N={Nl}", t_new, t_ref,
                extra={"pair_evals_per_s": npass * Bl * Nl * Nl / t_new * 1e3,
                       "ref": "reference torch program (bmm + min/topk + autograd, FP32 matmul) on the same GPU"})

        def c1_new():
            adv.grad = None
            with F.shared_distance_pass():
                (mods["ChamferDist"][0](adv, ori) + mods["HausdorffDist"][0](adv, ori) + mods["KNNDist(k=5)"][0](adv)).backward()

        def c1_ref():
            adv.grad = None
            (tpl.chamfer_dist(adv, ori) + tpl.hausdorff_dist(adv, ori) + tpl.knn_dist(adv)).backward()

        row(f"config-1 step CD+HD+kNN fwd+bwd B={Bl} N={Nl} (eager launches)", timeit(c1_new, flush=flush),
            timeit(c1_ref, iters=5, flush=flush),
            extra={"ref": "reference torch program on the same GPU (it builds P twice + the kNN matrix)"})
    # ---- config 3 shapes: PointNet++ SSG, batch 64 x 1024 ----------------------------------------------------
    B, N = 64, 1024
    xyz = torch.randn(B, N, 3, device="cuda")
    xyz = xyz / xyz.norm(dim=-1).amax(dim=1)[:, None, None]
    for npoint in (512, 128, 51):
        t_new = timeit(lambda: _ext.furthest_point_sampling(xyz, npoint))
        t_ref = timeit(lambda: ref.furthest_point_sampling(xyz, npoint)) if ref else None
        row(f"furthest_point_sampling B={B} N={N} npoint={npoint}", t_new, t_ref, extra={"rounds_per_s": B * npoint / t_new * 1e3})
    from oracle import torch_port as tp

    start = torch.randint(0, N, (B,), device="cuda")
    row(f"torch-semantics FPS B={B} N={N} npoint=512", timeit(lambda: F.fps_torch(xyz, 512, start)),
        timeit(lambda: tp.farthest_point_sample(xyz, 512, start), iters=3), extra={"ref": "reference torch program (512 rounds of ~8 kernels) on the same GPU"})
    fps = _ext.furthest_point_sampling(xyz, 512)
    new_xyz = _ext.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    for (r, ns) in ((0.2, 32), (0.4, 64)):
        t_new = timeit(lambda: _ext.ball_query(new_xyz, xyz, r, ns))
        t_ref = timeit(lambda: ref.ball_query(new_xyz, xyz, r, ns)) if ref else None
        row(f"ball_query B={B} N={N} S=512 r={r} ns={ns}", t_new, t_ref, bytes_alg=B * (N + 512) * 12 + B * 512 * ns * 4)
    row(f"torch-semantics query_ball_point B={B} N={N} S=512 r=0.2 ns=32", timeit(lambda: ms.query_ball_point(0.2, 32, xyz, new_xyz)),
        timeit(lambda: tp.query_ball_point(0.2, 32, xyz, new_xyz)), extra={"ref": "reference torch program (int64 [B,S,N] sort) on the same GPU"})
    bidx = ms.query_ball_point(0.2, 32, xyz, new_xyz)
    row(f"torch-semantics index_points (group) B={B} N={N} S=512 ns=32 C=3", timeit(lambda: ms.index_points(xyz, bidx)),
        timeit(lambda: tp.index_points(xyz, bidx)), extra={"ref": "reference torch program (advanced indexing) on the same GPU"})
    row(f"torch-semantics square_distance B={B} S=512 N={N}", timeit(lambda: ms.square_distance(new_xyz, xyz)),
        timeit(lambda: tp.square_distance(new_xyz, xyz)), bytes_alg=4 * B * 512 * N, extra={"ref": "reference torch program (matmul + 2 adds) on the same GPU"})
    # the whole set-abstraction front end of PointNet++ SSG level 1 (model/pointnet2_utils.py:110-138 sample_and_group)
    feats = torch.randn(B, N, 3, device="cuda")

    def sag(fps_fn, ball_fn, idx_fn):
        f = fps_fn(xyz, 512, start)
        nx = idx_fn(xyz, f)
        gi = ball_fn(0.2, 32, xyz, nx)
        g = idx_fn(xyz, gi) - nx.view(B, 512, 1, 3)
        return torch.cat([g, idx_fn(feats, gi)], dim=-1)

    row(f"sample_and_group (SSG level 1) B={B} N={N} npoint=512 r=0.2 ns=32",
        timeit(lambda: sag(F.fps_torch, ms.query_ball_point, ms.index_points)),
        timeit(lambda: sag(tp.farthest_point_sample, tp.query_ball_point, tp.index_points), iters=3),
        extra={"ref": "reference torch program on the same GPU"})
    for (C, S, ns, n) in ((3, 512, 32, 1024), (131, 128, 64, 512), (64, 512, 32, 1024)):
        pts = torch.randn(B, C, n, device="cuda")
        idx = torch.randint(0, n, (B, S, ns), device="cuda", dtype=torch.int32)
        by = 4 * B * (C * n + S * ns + C * S * ns)
        t_new = timeit(lambda: _ext.group_points(pts, idx), flush=flush)
        t_ref = timeit(lambda: ref.group_points(pts, idx), flush=flush) if ref else None
        row(f"group_points B={B} C={C} n={n} S={S} ns={ns}", t_new, t_ref, bytes_alg=by)
        go = torch.randn(B, C, S, ns, device="cuda")
        t_new = timeit(lambda: _ext.group_points_grad(go, idx, n), flush=flush)
        t_ref = timeit(lambda: ref.group_points_grad(go, idx, n), flush=flush) if ref else None
        row(f"group_points_grad B={B} C={C} n={n} S={S} ns={ns}", t_new, t_ref, bytes_alg=by + 4 * B * C * n)
    # QueryAndGroup's body after the ball query (set-abstraction level 2 shape): fused pass vs the composition
    from hitgeom.pointnet2_ops import pointnet2_utils as pu

    xyz2 = new_xyz  # (B,512,3)
    nx2 = new_xyz[:, :128].contiguous()
    f2 = torch.randn(B, 128, 512, device="cuda")
    idx2 = pu.ball_query(0.4, 64, xyz2, nx2)

    def composed():
        rel = pu.grouping_operation(xyz2.transpose(1, 2).contiguous(), idx2) - nx2.transpose(1, 2).unsqueeze(-1)
        return torch.cat([rel, pu.grouping_operation(f2, idx2)], dim=1)

    row("QueryAndGroup body B=64 C=128 n=512 S=128 ns=64 (fused group_concat)", timeit(lambda: pu.GroupConcat.apply(xyz2, nx2, f2, idx2), flush=flush),
        timeit(composed, flush=flush), bytes_alg=4 * B * 131 * 128 * 64, extra={"ref": "hitgeom's own group kernels + subtraction + torch.cat (the reference's composition)"})
    unknown, known = xyz, new_xyz[:, :128].contiguous()
    t_new = timeit(lambda: _ext.three_nn(unknown, known))
    t_ref = timeit(lambda: ref.three_nn(unknown, known)) if ref else None
    row(f"three_nn B={B} n={N} m=128", t_new, t_ref)
    d2, i3 = _ext.three_nn(unknown, known)
    w = torch.rand(B, N, 3, device="cuda")
    feats = torch.randn(B, 256, 128, device="cuda")
    t_new = timeit(lambda: _ext.three_interpolate(feats, i3, w), flush=flush)
    t_ref = timeit(lambda: ref.three_interpolate(feats, i3, w), flush=flush) if ref else None
    row(f"three_interpolate B={B} c=256 m=128 n={N}", t_new, t_ref, bytes_alg=4 * B * (256 * 128 + 6 * N + 256 * N))
    # ---- config 4 shapes: DGCNN k=20, batch 32 x 1024 ----------------------------------------------------------
    for C in (3, 64, 128):
        x = torch.randn(32, C, 1024, device="cuda")
        t_new = timeit(lambda: ms.knn(x, 20))

        def ref_knn():  # the reference's torch program for DGCNN knn (model/dgcnn_cls.py:7-13) on the same GPU
            inner = -2 * torch.matmul(x.transpose(2, 1), x)
            xx = torch.sum(x ** 2, dim=1, keepdim=True)
            return (-xx - inner - xx.transpose(2, 1)).topk(k=20, dim=-1)[1]

        row(f"DGCNN knn B=32 C={C} N=1024 k=20", t_new, timeit(ref_knn), extra={"pair_evals_per_s": 32 * 1024 * 1024 / t_new * 1e3, "ref": "torch matmul+topk on the same GPU"})
    # DGCNN edge features (get_graph_feature after its kNN), the four layers of config 4
    from oracle import torch_port as tp

    for C in (3, 64, 128):
        Bq, Nq, kq = 32, 1024, 20
        x = torch.randn(Bq, C, Nq, device="cuda")
        idx = ms.knn(x, kq)
        by = 4 * Bq * (C * Nq + 2 * C * Nq * kq) + 8 * Bq * Nq * kq
        t_new = timeit(lambda: F.edge_feature(x, idx), flush=flush)
        t_ref = timeit(lambda: tp.get_graph_feature(x, idx), flush=flush)
        row(f"edge_feature B={Bq} C={C} N={Nq} k={kq}", t_new, t_ref, bytes_alg=by, extra={"ref": "reference tensor program (gather+repeat+cat+permute) on the same GPU"})
        go = torch.randn(Bq, 2 * C, Nq, kq, device="cuda")
        xg = x.clone().requires_grad_()

        def bwd_new():
            xg.grad = None
            F.edge_feature(xg, idx).backward(go)

        def bwd_ref():
            xg.grad = None
            tp.get_graph_feature(xg, idx).backward(go)

        t_new_fb, t_ref_fb = timeit(bwd_new, flush=flush), timeit(bwd_ref, flush=flush)
        row(f"edge_feature fwd+bwd B={Bq} C={C} N={Nq} k={kq}", t_new_fb, t_ref_fb, bytes_alg=2 * by,
            extra={"ref": "reference tensor program + autograd on the same GPU"})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"hbm_peak_gbs_measured": hbm, "results": res}, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
