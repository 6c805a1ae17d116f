"""hitgeom -- B200-native (sm_100a) point-set geometry for HiT-ADV's attack hot path.

Host-side mirror of the reference's operator interface for that path (same class / function names, argument
meaning and error behaviour), calling hand-written CUDA kernels through the C ABI of include/hitgeom.h:

    hitgeom.set_distance   ChamferDistance, HausdorffDistance, chamfer, hausdorff      (util/set_distance.py)
    hitgeom.dist_utils     ChamferDist, HausdorffDist, KNNDist, ChamferkNNDist, L2Dist (util/dist_utils.py)
    hitgeom.pointnet2_ops  _ext (nine native functions) + autograd wrappers            (pointnet2_ops)
    hitgeom.model_seams    square_distance, index_points, farthest_point_sample,
                           query_ball_point, knn                  (model/pointnet2_utils.py, model/dgcnn_cls.py)
    hitgeom.pytorch3d_ops  knn_points, knn_gather                                      (pytorch3d.ops)
    hitgeom.install()      registers the above under the names the unmodified reference imports
    hitgeom.host           ChamferKnnHostStep: the CW-kNN distance step on HOST buffers, chunk-pipelined over two streams
    hitgeom.sharding       instance-sharded multi-GPU driver (one process per GPU, NCCL all-gather at the end)
    hitgeom.hit_adv        HiT_ADV attacker with the fused deformation kernel and on-device bookkeeping
                           (ShapeAttack/HiT_ADV.py; SURVEY.md 8f "next" rows #1, #2)
    hitgeom.cw_knn         CWKNN / CWUKNN attack loops, sync-free and CUDA-graph replayable (CW/kNN.py, CW/UKNN.py)
    hitgeom.clip_utils     ClipPointsL2 / ClipPointsLinf / ProjectInnerPoints / ProjectInnerClipLinf (util/clip_utils.py)
    hitgeom.adv_utils      LogitsAdvLoss / UntargetedLogitsAdvLoss / CrossEntropyAdvLoss           (util/adv_utils.py)
    hitgeom.eval_metrics   uniform_loss, kNN_smoothing_loss, CurvStdDist -- eval_ASR's per-batch metrics
                           (FGM/GeoA3_args.py:240-302, util/dist_utils.py:464-495)

There is no CPU path: importing works anywhere (so the build can be checked without a GPU), but every
operator raises unless its tensors live on a CUDA device and libhitgeom.so is present.
"""
from . import _lib
from ._lib import HitgeomError, build, device_info, exported_symbols, lib  # noqa: F401
from .install import install, patch_reference  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    import importlib

    if name in ("set_distance", "dist_utils", "pointnet2_ops", "model_seams", "pytorch3d_ops", "functional",
                "sharding", "hit_adv", "cw_knn", "clip_utils", "adv_utils", "eval_metrics", "host"):
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
