"""Parity of the kNN kernels (KNNDist, DGCNN knn, pytorch3d-style knn_points) against golden vectors and the
oracle.  Values bit-exact; indices bit-exact (canonical lowest-index tie order == the reference's on these
inputs); loss 1e-6; gradient 1e-5 norm-wise."""
import numpy as np
import pytest
import torch

from util_inputs import clouds, jitter, normwise

pytestmark = pytest.mark.gpu


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def F():
    from hitgeom import functional

    return functional


def test_knn_topk_golden(golden, F):
    g = golden("loss_classes")
    vals, idx = F.knn_self(gpu(g["adv"]), 6)
    assert np.array_equal(vals.cpu().numpy(), g["knn_topk_vals"])
    assert np.array_equal(idx.cpu().numpy(), g["knn_topk_idx"])


@pytest.mark.parametrize("B,K,k1,kind", [(2, 6, 6, "gauss"), (3, 100, 1, "gauss"), (2, 777, 6, "surface"),
                                         (2, 1024, 5, "surface"), (1, 2500, 17, "gauss"), (2, 513, 20, "surface"),
                                         (1, 640, 32, "gauss"), (2, 300, 21, "gauss")])
def test_knn_self_vs_oracle(oracle, F, B, K, k1, kind):
    pc = jitter(clouds(B, K, 300 + K, kind), 11)
    vals, idx = F.knn_self(gpu(pc), k1)
    ov, oi = oracle.knn_self(pc, k1, threads=4)
    assert np.array_equal(vals.cpu().numpy(), ov)
    assert np.array_equal(idx.cpu().numpy(), oi)


@pytest.fixture(params=["fused-small-backward", "general-backward"])
def bwd_path(request):
    """KNNDist's backward for 3-D clouds that fit in shared memory is one fused kernel (knn_outlier_bwd_small_kernel);
    hg_tune("small_fused", 1) sends the same call down the general keys + reverse-map + gradient kernels."""
    from hitgeom._lib import lib

    lib().hg_tune(b"small_fused", 0 if request.param == "fused-small-backward" else 1)
    yield request.param
    lib().hg_tune(b"small_fused", 0)


@pytest.mark.parametrize("k,alpha", [(5, 1.05), (4, 1.05), (8, 0.5)])
def test_knn_dist_golden(golden, k, alpha, bwd_path):
    from hitgeom.dist_utils import KNNDist

    g = golden("loss_classes")
    w = torch.from_numpy(g["w"])
    for layout in ("BK3", "B3K"):
        for wt, wtag in [(None, "now"), (w, "w")]:
            a = gpu(g["adv"] if layout == "BK3" else g["adv"].transpose(0, 2, 1)).requires_grad_()
            loss = KNNDist(k=k, alpha=alpha)(a, weights=wt, batch_avg=False)
            loss.sum().backward()
            key = f"knn_k{k}_{layout}_{wtag}"
            np.testing.assert_allclose(loss.detach().cpu().numpy(), g[key], rtol=1e-6, atol=1e-12, err_msg=key)
            gr, ref = a.grad.cpu().numpy(), g[key + "_grad"]
            assert np.abs(gr - ref).max() <= 1e-5 * np.abs(ref).max(), key


def test_knn_dist_vs_oracle_config1(oracle, bwd_path):
    """Config 1 shape (388 x 1024): loss and gradient against the oracle."""
    from hitgeom.dist_utils import KNNDist

    pc = jitter(clouds(388, 1024, 1234), 99)
    a = gpu(pc).requires_grad_()
    loss = KNNDist(k=5, alpha=1.05)(a, batch_avg=False)
    loss.sum().backward()
    ol, (vals, idx, value, mask) = oracle.knn_dist(pc, 5, 1.05, threads=oracle.host_threads())
    np.testing.assert_allclose(loss.detach().cpu().numpy(), ol, rtol=1e-6)
    og = oracle.knn_outlier_bwd(pc, idx, mask, np.ones(388, np.float32))
    assert normwise(a.grad.cpu().numpy(), og) < 1e-5


@pytest.mark.parametrize("C,k", [(3, 20), (3, 5), (64, 20), (128, 20)])
def test_dgcnn_knn_golden(golden, C, k):
    from hitgeom.model_seams import knn

    g = golden("dgcnn_knn")
    x = gpu(g["x3" if C == 3 else f"x{C}"])  # [B,C,N] channel-major, as DGCNN passes it
    idx = knn(x, k)
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), g[f"idx{C}_k{k}"])


@pytest.mark.parametrize("C,K,k1", [(64, 1024, 20), (16, 257, 7), (5, 130, 30)])
def test_knn_generic_c_vs_oracle(oracle, F, C, K, k1):
    rng = np.random.default_rng(C)
    pc = rng.standard_normal((2, K, C)).astype(np.float32)
    vals, idx = F.knn_self(gpu(pc), k1)
    ov, oi = oracle.knn_self(pc, k1, threads=4)
    assert np.array_equal(vals.cpu().numpy(), ov)
    assert np.array_equal(idx.cpu().numpy(), oi)


@pytest.mark.parametrize("N,M,K", [(1024, 1024, 17), (256, 1024, 17), (51, 40, 3), (100, 300, 1)])
def test_knn_points_vs_oracle(oracle, N, M, K):
    """pytorch3d.ops.knn_points stand-in (parity unpinned upstream; pinned to the oracle's restatement)."""
    from hitgeom.pytorch3d_ops import knn_gather, knn_points

    p1 = clouds(2, N, 1, "surface")
    p2 = clouds(2, M, 2, "surface")
    out = knn_points(gpu(p1), gpu(p2), K=K, return_nn=True)
    od, oi = oracle.knn_points(p1, p2, K)
    assert np.array_equal(out.dists.cpu().numpy(), od)
    assert np.array_equal(out.idx.cpu().numpy(), oi)
    assert out.idx.dtype == torch.int64
    gathered = knn_gather(gpu(p2), out.idx).cpu().numpy()
    assert np.array_equal(gathered, p2[np.arange(2)[:, None, None], oi])
    assert np.array_equal(out.knn.cpu().numpy(), gathered)


def test_config4_dgcnn_knn_full_size(oracle):
    """BASELINE config 4 shape: DGCNN first edge-conv kNN graph, batch 32 x 1024 points, k = 20 (self included)."""
    from hitgeom.model_seams import knn

    pts = jitter(clouds(32, 1024, 4040, "surface"), 5)
    idx = knn(gpu(pts.transpose(0, 2, 1)), 20)
    _, oi = oracle.knn_self(pts, 20, threads=oracle.host_threads())
    assert np.array_equal(idx.cpu().numpy(), oi)


def test_knn_seeded_and_unseeded_agree(oracle, F):
    """The grid-seeded thresholds (N >= 512) must not change any value or index: degenerate clouds included."""
    for kind, n in (("surface", 4096), ("gauss", 600), ("gauss", 511)):
        pc = clouds(2, n, 7 + n, kind)
        pc[1, : n // 2] = pc[1, n // 2 : n // 2 * 2]  # half of cloud 1 duplicated: crowded cells, exact ties
        vals, idx = F.knn_self(gpu(pc), 6)
        ov, oi = oracle.knn_self(pc, 6, threads=4)
        assert np.array_equal(vals.cpu().numpy(), ov) and np.array_equal(idx.cpu().numpy(), oi), (kind, n)
    flat = np.zeros((1, 2048, 3), np.float32)  # every point identical: one cell, all distances equal
    vals, idx = F.knn_self(gpu(flat), 6)
    assert np.array_equal(idx.cpu().numpy()[0], np.tile(np.arange(6, dtype=np.int32), (2048, 1)))


@pytest.mark.parametrize("k1", [6, 20])
def test_knn_folded_filter_is_conservative_far_from_origin(oracle, F, k1):
    """The seeded path filters with a 4-operation folded form whose rounding differs from the reference's; its slack
    is sized from (|q| + max|c|)^2.  A small cloud far from the origin is the worst case: the expanded-form distances
    are cancellation noise (norms ~ 1e4, spacings ~ 1e-2), many candidates tie or sit within the slack band -- values
    and indices must still be the oracle's, bit for bit."""
    rng = np.random.default_rng(k1)
    for offset, scale in (((60.0, -35.0, 20.0), 0.05), ((0.5, 0.5, 0.5), 1.0), ((1000.0, 0.0, 0.0), 1.0)):
        pc = (clouds(2, 2048, 91, "gauss") * scale + np.asarray(offset, np.float32)).astype(np.float32)
        pc[1, :64] = pc[1, 64:128]  # duplicates
        pc[0] = pc[0][rng.permutation(2048)]
        vals, idx = F.knn_self(gpu(pc), k1)
        ov, oi = oracle.knn_self(pc, k1, threads=4)
        assert np.array_equal(vals.cpu().numpy(), ov) and np.array_equal(idx.cpu().numpy(), oi), (offset, scale)


def test_knn_generic_c_heavy_ties(oracle, F):
    """Feature clouds with massive exact ties (every point one of four prototypes): the row select's candidate buffer
    overflows and it falls back to the full-row rounds; values and lowest-index-first order must still be the oracle's."""
    rng = np.random.default_rng(5)
    protos = rng.standard_normal((4, 16)).astype(np.float32)
    pc = protos[rng.integers(0, 4, size=(2, 600))]  # [2,600,16]
    pc[1, :300] = protos[0]
    vals, idx = F.knn_self(gpu(pc), 20)
    ov, oi = oracle.knn_self(pc, 20, threads=2)
    assert np.array_equal(vals.cpu().numpy(), ov) and np.array_equal(idx.cpu().numpy(), oi)


def test_knn_outlier_backward_with_hub_points_vs_oracle(oracle, bwd_path):
    """Padding by repetition: hundreds of points coincide, so a few points are the neighbour of hundreds of others
    (in-degree far above the 32 the selection walk handles): the gradient is still the oracle's, promptly."""
    import time

    from hitgeom.dist_utils import KNNDist

    pc = clouds(2, 2048, 31)
    pc[0, 1000:1600] = pc[0, 5]
    pc[1, 100:2000] = pc[1, 99] + 1e-4 * np.random.default_rng(0).standard_normal((1900, 3)).astype(np.float32)
    a = gpu(pc).requires_grad_()
    loss = KNNDist(k=5, alpha=0.2)(a, batch_avg=False)
    torch.cuda.synchronize()
    t0 = time.time()
    loss.sum().backward()
    torch.cuda.synchronize()
    assert time.time() - t0 < 2.0
    ol, (vals, idx, value, mask) = oracle.knn_dist(pc, 5, 0.2, threads=2)
    assert mask.sum() > 0
    np.testing.assert_allclose(loss.detach().cpu().numpy(), ol, rtol=1e-6, atol=1e-12)
    # (sums of hundreds of terms: 1e-4, see test_backward_with_hub_points_vs_oracle)
    assert normwise(a.grad.cpu().numpy(), oracle.knn_outlier_bwd(pc, idx, mask, np.ones(2, np.float32))) < 1e-4


def test_knn_outlier_backward_fused_and_general_paths_give_the_same_bits():
    """Both backward paths add the same terms in the same (ascending edge) order: identical gradients, hubs included."""
    from hitgeom._lib import lib
    from hitgeom.dist_utils import KNNDist

    pc = jitter(clouds(24, 1024, 77), 5)
    pc[3, 200:420] = pc[3, 7]
    grads = []
    for off in (0, 1):
        lib().hg_tune(b"small_fused", off)
        try:
            a = gpu(pc).requires_grad_()
            KNNDist(k=5, alpha=0.3)(a, batch_avg=False).sum().backward()
            grads.append(a.grad.clone())
        finally:
            lib().hg_tune(b"small_fused", 0)
    assert grads[0].abs().max() > 0 and torch.equal(grads[0], grads[1])


# ---- tensor-core filter + exact evaluation (hg_knn_tc.cu): same values and indices as the FP32 path and the oracle --------------------
@pytest.mark.parametrize("tc_off", [2, 5, 1])
@pytest.mark.parametrize("C,K,k1,kind", [(64, 1024, 20, "gauss"), (128, 512, 20, "gauss"), (32, 300, 7, "gauss"),
                                         (96, 777, 32, "gauss"), (64, 512, 20, "prototypes"), (64, 640, 5, "scaled"),
                                         (128, 1024, 20, "dups")])
def test_knn_feature_clouds_tensor_core_vs_oracle(oracle, F, C, K, k1, kind, tc_off):
    """DGCNN-shaped feature clouds (C a multiple of 32, 256 <= K <= 4096) take the tcgen05 TF32 filter fused with the
    exact FP32 evaluation (`hg_tune("knn_tc", 2)` forces it for batches too small to fill the machine, as here);
    `hg_tune("knn_tc", 1)` forces the FP32 tile + row-select path.  Both must give the oracle's values
    and indices bit for bit -- including feature clouds made of a few prototypes (massive exact ties: every column is
    a hit and is evaluated exactly), widely different feature scales (the TF32 error bound is per row), duplicated
    points, a cloud size that is no multiple of the tile (K = 300, 777) and lists longer than the neighbour count
    (k = 7 in a 20-slot list, 32 in 32)."""
    from hitgeom._lib import lib

    rng = np.random.default_rng(C + K)
    pc = rng.standard_normal((2, K, C)).astype(np.float32)
    if kind == "prototypes":
        protos = rng.standard_normal((5, C)).astype(np.float32)
        pc = protos[rng.integers(0, 5, size=(2, K))]
        pc[1, : K // 2] = protos[0]
    elif kind == "scaled":
        pc *= np.exp(rng.uniform(-4, 4, size=(2, K, 1))).astype(np.float32)
    elif kind == "dups":
        pc[0, 100:400] = pc[0, 500:800]
        pc[1, :] = pc[1, rng.integers(0, 64, size=K)]
    lib().hg_tune(b"knn_tc", tc_off)
    try:
        vals, idx = F.knn_self(gpu(pc), k1)
    finally:
        lib().hg_tune(b"knn_tc", 0)
    ov, oi = oracle.knn_self(pc, k1, threads=2)
    assert np.array_equal(idx.cpu().numpy(), oi), (kind, "indices")
    assert np.array_equal(vals.cpu().numpy(), ov), (kind, "values")


@pytest.mark.parametrize("K,k1", [(1024, 20), (777, 12), (2048, 32), (300, 21)])
def test_knn_3d_long_lists_on_the_tensor_core_kernel_vs_oracle(oracle, F, K, k1):
    """3-D clouds with long neighbour lists (DGCNN layer 1, k = 20) take the tensor-core kernel on a 32-channel padded copy
    when the batch fills the machine (`hg_tune("knn_tc", 2)` forces it here): the zero channels change nothing in the
    reference's FMA chain, so values and indices are the 3-D kernels' and the oracle's bit for bit -- surface clouds with
    exact duplicates, a cloud far from the origin (large norms: the TF32 error bound is per row), identical points."""
    from hitgeom._lib import lib

    a = clouds(3, K, 300 + K, "surface")
    a[1, : K // 4] = a[1, K // 4 : 2 * (K // 4)]
    a[2, : K // 2] = a[2, 0]
    b = (clouds(2, K, 77 + K, "gauss") * 0.05 + np.asarray((60.0, -35.0, 20.0), np.float32)).astype(np.float32)
    for tag, pc in (("surface+dups", a), ("far from origin", b)):
        ov, oi = oracle.knn_self(pc, k1, threads=2)
        outs = {}
        for knob in (2, 1):  # tensor-core kernel forced / switched off (3-D small-cloud kernel)
            lib().hg_tune(b"knn_tc", knob)
            try:
                vals, idx = F.knn_self(gpu(pc), k1)
            finally:
                lib().hg_tune(b"knn_tc", 0)
            assert np.array_equal(idx.cpu().numpy(), oi), (tag, knob, "indices")
            assert np.array_equal(vals.cpu().numpy(), ov), (tag, knob, "values")
