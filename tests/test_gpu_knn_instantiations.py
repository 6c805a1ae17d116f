"""Every instantiation of the streaming 3-D kNN kernel against the oracle, bit for bit.

`launch_form` (hit-adv_b200/csrc/hg_knn3.cu) picks queries-per-lane (QT), list length (KM) and candidate pairs per
filter bit (GP) from the batch size, the cloud size and the SM count, so small seeded inputs only ever reach QT = 1.
The headline benchmark (config 5 shard, 1024 x 16384) runs `knn3_kernel<FOLD4, QT=4, KM=6, GP=4>`; config 1
(388 x 1024) runs the small-cloud kernel.  Here `hg_knn_force_shape` makes small inputs reach EVERY (QT, GP) pair, and
three batches large enough to select the big-batch variants naturally are checked as well.  Values and indices
(canonical lowest-index-first tie order) must equal `oracle.knn_self` / `oracle.knn_points` exactly; the inputs
include exact duplicates (ties), near-origin points and a cloud far from the origin (where the folded filter's slack
matters).  Reference: util/dist_utils.py:148-156, model/dgcnn_cls.py:7-13."""
import ctypes

import numpy as np
import pytest
import torch

from util_inputs import clouds, jitter

pytestmark = pytest.mark.gpu


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture()
def F():
    from hitgeom import functional

    yield functional
    functional.force_knn_shape(0, 0)
    functional.tune_knn_small(0)


def _inputs(n):
    a = clouds(3, n, 1000 + n, "surface")
    a[1, : n // 4] = a[1, n // 4 : 2 * (n // 4)]  # a quarter of cloud 1 duplicated: exact ties
    b = (clouds(2, n, 77 + n, "gauss") * 0.05 + np.asarray((60.0, -35.0, 20.0), np.float32)).astype(np.float32)
    b[0, :32] = b[0, 32:64]
    return [("surface+dups", a), ("far from origin", b)]


QT_FOR_K = {6: (1, 2, 4), 5: (1, 2, 4), 1: (4,), 20: (1, 2), 12: (2,), 32: (1,), 21: (1,)}


@pytest.mark.parametrize("n,path", [(2048, "streaming"), (3000, "streaming"), (2048, "small"), (1024, "small"),
                                    (3000, "small"), (333, "small")])
@pytest.mark.parametrize("gp", [1, 2, 4])
@pytest.mark.parametrize("k1", sorted(QT_FOR_K))
def test_self_knn_every_forced_instantiation(oracle, F, n, path, gp, k1):
    """Self-kNN (KNNDist, DGCNN layer 1).  Clouds up to 4096 points take the small-cloud kernel (Z-order seeds, deferred
    drain; GP = 1, 2, 4), larger ones -- and these, with the small-cloud path switched off -- the grid-seeded streaming
    kernel (GP = 2 or 4; a forced 1 means 4 there); every QT the list length admits."""
    if path == "streaming":
        F.tune_knn_small(-1)
    for tag, pc in _inputs(n):
        ov, oi = oracle.knn_self(pc, k1, threads=oracle.host_threads())
        for qt in QT_FOR_K[k1]:
            F.force_knn_shape(qt, gp)
            vals, idx = F.knn_self(gpu(pc), k1)
            assert np.array_equal(vals.cpu().numpy(), ov), (tag, n, path, qt, gp, k1, "values")
            assert np.array_equal(idx.cpu().numpy(), oi), (tag, n, path, qt, gp, k1, "indices")


def test_forced_shape_without_instantiation_fails_loudly(F):
    from hitgeom._lib import HitgeomError

    F.force_knn_shape(4, 4)
    with pytest.raises(HitgeomError):
        F.knn_self(gpu(clouds(1, 2048, 3)), 20)  # QT = 4 exists for lists of <= 6 entries only


@pytest.mark.parametrize("gp", [2, 4])
@pytest.mark.parametrize("k1,qts", [(6, (1, 2, 4)), (17, (1, 2)), (32, (1,))])
def test_knn_points_every_forced_instantiation(oracle, F, gp, k1, qts):
    """DIRECT form (pytorch3d.ops.knn_points stand-in; parity unpinned upstream, pinned to the oracle's restatement)."""
    from hitgeom.functional import knn_points_raw

    p1 = clouds(2, 700, 5, "surface")
    p2 = clouds(2, 2300, 6, "surface")
    p2[1, :100] = p2[1, 100:200]
    od, oi = oracle.knn_points(p1, p2, k1)
    for qt in qts:
        F.force_knn_shape(qt, gp)
        d, i = knn_points_raw(gpu(p1), gpu(p2), k1)
        assert np.array_equal(d.cpu().numpy(), od) and np.array_equal(i.cpu().numpy(), oi), (qt, gp, k1)


@pytest.mark.parametrize("gp", [2, 4])
@pytest.mark.parametrize("k1,qts", [(6, (1, 2, 4)), (20, (1, 2))])
def test_unseeded_expanded_form_every_forced_instantiation(oracle, F, gp, k1, qts):
    """The 5-operation EXPANDED filter (what hg_knn_self_f32 runs when the caller passes no workspace)."""
    from hitgeom._lib import check, lib, ptr, stream_ptr

    pc = _inputs(1500)[0][1]
    ov, oi = oracle.knn_self(pc, k1, threads=4)
    B, K, _ = pc.shape
    x = gpu(pc)
    for qt in qts:
        F.force_knn_shape(qt, gp)
        vals = torch.empty((B, K, k1), dtype=torch.float32, device="cuda")
        idx = torch.empty((B, K, k1), dtype=torch.int32, device="cuda")
        check(lib().hg_knn_self_f32(ptr(x), B, K, 3, k1, ptr(vals), ptr(idx), None, ctypes.c_size_t(0), stream_ptr()),
              "hg_knn_self_f32")
        assert np.array_equal(vals.cpu().numpy(), ov) and np.array_equal(idx.cpu().numpy(), oi), (qt, gp, k1)


@pytest.mark.parametrize("B,n,k1", [(24, 8192, 6), (12, 8192, 20), (12, 16384, 6), (80, 2048, 6), (160, 2048, 20),
                                    (388, 1024, 6), (300, 1024, 20)])
def test_big_batch_variants_selected_naturally(oracle, F, B, n, k1):
    """Batches large enough for the dispatch to pick the big-batch instantiations by itself (no override): on a
    148-SM part 24 x 8192 / k+1 = 6 and 12 x 16384 select <FOLD4, QT=4, KM=6, GP=4> -- the headline kernel of
    bench.py's config-5 shard -- 12 x 8192 / k+1 = 20 selects <FOLD4, 2, 20, 4>; up to 4096 points the small-cloud
    kernel runs (388 x 1024 is config 1)."""
    F.force_knn_shape(0, 0)
    pc = jitter(clouds(B, n, 4242 + n, "surface" if B < 100 else "gauss"), 3)
    pc[0, : n // 8] = pc[0, n // 8 : 2 * (n // 8)]
    vals, idx = F.knn_self(gpu(pc), k1)
    ov, oi = oracle.knn_self(pc, k1, threads=oracle.host_threads())
    assert np.array_equal(vals.cpu().numpy(), ov)
    assert np.array_equal(idx.cpu().numpy(), oi)


# ---- temporal seeds (hg_knn_self_temporal_f32): results never depend on the state ------------------------------------
@pytest.mark.parametrize("n,k1", [(1024, 6), (700, 20), (2048, 6), (3000, 17), (40, 32), (5000, 6), (4500, 20)])
def test_temporal_seeds_never_change_results(oracle, F, n, k1):
    """State from a previous call on a nearby cloud, stale state, and garbage state (out of range, repeated entries,
    all zeros): values and indices are the oracle's bit for bit, and the state ends up holding this call's indices."""
    rng = np.random.default_rng(n + k1)
    base = clouds(3, n, 50 + n, "surface")
    base[2, : n // 4] = base[2, n // 4 : 2 * (n // 4)]
    moved = (base + 2e-3 * rng.standard_normal(base.shape)).astype(np.float32)
    ov0, oi0 = oracle.knn_self(base, k1, threads=3)
    ov1, oi1 = oracle.knn_self(moved, k1, threads=3)
    state = torch.empty((3, n, k1), dtype=torch.int32, device="cuda")
    v, i = F.knn_self(gpu(base), k1, state=state, state_valid=False)  # first call: state is output only
    assert np.array_equal(v.cpu().numpy(), ov0) and np.array_equal(i.cpu().numpy(), oi0)
    assert torch.equal(state, i)
    v, i = F.knn_self(gpu(moved), k1, state=state, state_valid=True)  # seeds = neighbours of the previous cloud
    assert np.array_equal(v.cpu().numpy(), ov1) and np.array_equal(i.cpu().numpy(), oi1)
    assert torch.equal(state, i)
    for tag, junk in (("out of range", rng.integers(-5, 2 * n, size=(3, n, k1))),
                      ("repeated", np.repeat(rng.integers(0, n, size=(3, n, 1)), k1, axis=2)),
                      ("zeros", np.zeros((3, n, k1))),
                      ("far neighbours", rng.permuted(np.tile(np.arange(n), (3, 1)), axis=1)[:, :, None]
                       + np.arange(k1)[None, None, :] * 0 + rng.integers(0, n, size=(3, n, k1)))):
        st = torch.from_numpy(np.ascontiguousarray(junk % (4 * n) - (n if tag == "out of range" else 0)).astype(np.int32)).cuda()
        v, i = F.knn_self(gpu(moved), k1, state=st, state_valid=True)
        assert np.array_equal(v.cpu().numpy(), ov1), (tag, "values")
        assert np.array_equal(i.cpu().numpy(), oi1), (tag, "indices")
        assert torch.equal(st, i), tag


def test_knndist_temporal_state_same_bits_as_stateless():
    """KNNDist with temporal seeds inside a (mock) attack loop: loss and gradient bit-identical to the stateless module."""
    from hitgeom.dist_utils import KNNDist

    a, b = KNNDist(k=5), KNNDist(k=5).temporal_seeds(True)
    x = gpu(clouds(4, 1024, 9)).requires_grad_()
    y = x.detach().clone().requires_grad_()
    for step in range(4):
        la = a(x, batch_avg=False)
        la.sum().backward()
        lb = b(y, batch_avg=False)
        lb.sum().backward()
        assert torch.equal(la, lb) and torch.equal(x.grad, y.grad), step
        with torch.no_grad():
            x -= 0.5 * x.grad
            y -= 0.5 * y.grad
        x.grad = None
        y.grad = None


# ---- more than 32 neighbours (HiT_ADV's default curv_loss_knn = 32 asks knn_points for K = 33) -----------
@pytest.mark.parametrize("K", [33, 50, 64])
def test_knn_points_up_to_64_neighbours(oracle, F, K):
    from hitgeom.pytorch3d_ops import knn_points

    p1 = clouds(2, 300, 15, "surface")
    p2 = clouds(2, 1500, 16, "surface")
    p2[0, :60] = p2[0, 60:120]
    od, oi = oracle.knn_points(p1, p2, K)
    for gp in (2, 4):
        F.force_knn_shape(1, gp)
        out = knn_points(gpu(p1), gpu(p2), K=K)
        assert np.array_equal(out.dists.cpu().numpy(), od) and np.array_equal(out.idx.cpu().numpy(), oi), (K, gp)
    F.force_knn_shape(0, 0)
    with pytest.raises(NotImplementedError):
        knn_points(gpu(p1), gpu(p2), K=65)


@pytest.mark.parametrize("n,k1", [(900, 40), (2500, 64)])
def test_self_knn_up_to_64_neighbours(oracle, F, n, k1):
    pc = _inputs(n)[0][1]
    ov, oi = oracle.knn_self(pc, k1, threads=3)
    vals, idx = F.knn_self(gpu(pc), k1)
    assert np.array_equal(vals.cpu().numpy(), ov) and np.array_equal(idx.cpu().numpy(), oi)


def test_default_constructed_hit_adv_runs():
    """HiT_ADV(model, adv_func) with the reference's defaults (curv_loss_knn = 32 -> knn_points K = 33)."""
    from hitgeom.hit_adv import HiT_ADV, UntargetedLogitsAdvLoss
    from util_models import TinyPointNet

    pts = clouds(3, 512, 8)
    nrm = np.random.default_rng(2).standard_normal(pts.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    data = torch.from_numpy(np.concatenate([pts, nrm], -1))
    att = HiT_ADV(TinyPointNet(40, seed=1), UntargetedLogitsAdvLoss(kappa=10.0), binary_step=1, num_iter=5,
                  cd_weight=1e-4, ker_weight=1.0, hide_weight=1.0)
    torch.manual_seed(0)
    best, succ = att.attack(data, torch.zeros(3, dtype=torch.long))
    assert best.shape == (3, 512, 3) and np.isfinite(best).all()
