"""Rebinds the reference's hot-path seams to hitgeom, WITHOUT editing any reference file.

    import hitgeom; hitgeom.install()          # before importing the reference's util/, model/, CW/ ...

What gets registered / patched:
  * `sys.modules['pointnet2_ops']`, `['pointnet2_ops._ext']`: the native module the reference's
    pointnet2_ops/pointnet2_utils.py:8 imports (otherwise it JIT-compiles for sm_37..sm_75, which fails);
  * `sys.modules['pytorch3d']`, `['pytorch3d.ops']`: `knn_points`, `knn_gather` (util/dist_utils.py:12);
  * after the reference modules are imported, `patch_reference()` swaps
        util.set_distance.chamfer / hausdorff, util.dist_utils.{chamfer,hausdorff,ChamferDist,HausdorffDist,
        KNNDist,ChamferkNNDist}.forward, model.pointnet2_utils.{square_distance,index_points,
        farthest_point_sample,query_ball_point}, model.dgcnn_cls.{knn,get_graph_feature}
    for the kernel-backed versions.  Callers (CW/*.py, ShapeAttack/HiT_ADV.py, util/other_utils.py) keep
    constructing and calling the same classes.
"""
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(pointnet2=True, pytorch3d=True):
    from . import pointnet2_ops as p2
    from . import pytorch3d_ops as p3

    if pointnet2 and "pointnet2_ops._ext" not in sys.modules:
        pkg = sys.modules.get("pointnet2_ops") or _module("pointnet2_ops", __path__=[])
        pkg._ext = p2._ext
        sys.modules["pointnet2_ops._ext"] = p2._ext
    if pytorch3d and "pytorch3d.ops" not in sys.modules:
        pkg = sys.modules.get("pytorch3d") or _module("pytorch3d", __path__=[])
        pkg.ops = _module("pytorch3d.ops", knn_points=p3.knn_points, knn_gather=p3.knn_gather)
    return True


def patch_reference():
    """Call after the reference's modules have been imported; patches whichever of them are loaded."""
    from . import dist_utils as du
    from . import model_seams as ms
    from . import set_distance as sd

    patched = []
    m = sys.modules.get("util.set_distance")
    if m is not None:
        m.chamfer, m.hausdorff = sd.chamfer, sd.hausdorff
        patched.append("util.set_distance")
    m = sys.modules.get("util.dist_utils")
    if m is not None:
        m.chamfer, m.hausdorff = sd.chamfer, sd.hausdorff
        for cls in ("ChamferDist", "HausdorffDist", "KNNDist", "ChamferkNNDist"):
            getattr(m, cls).forward = getattr(du, cls).forward
        patched.append("util.dist_utils")
    m = sys.modules.get("model.pointnet2_utils")
    if m is not None:
        for fn in ("square_distance", "index_points", "farthest_point_sample", "query_ball_point"):
            setattr(m, fn, getattr(ms, fn))
        patched.append("model.pointnet2_utils")
    m = sys.modules.get("model.dgcnn_cls")
    if m is not None:
        m.knn = ms.knn
        m.get_graph_feature = ms.get_graph_feature
        patched.append("model.dgcnn_cls")
    return patched
