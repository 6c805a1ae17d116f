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
        import importlib.util

        sys.modules["pointnet2_ops._ext"] = p2._ext  # what pointnet2_utils.py:8 imports, whoever the parent package is
        pkg = sys.modules.get("pointnet2_ops")
        if pkg is None:
            try:
                real = importlib.util.find_spec("pointnet2_ops")
            except (ImportError, ValueError):
                real = None
            if real is None:
                # no reference package on sys.path: stand-alone package with hitgeom's wrappers under the same names
                pkg = _module("pointnet2_ops", __path__=[], pointnet2_utils=p2.pointnet2_utils,
                              pointnet2_modules=p2.pointnet2_modules)
                sys.modules["pointnet2_ops.pointnet2_utils"] = p2.pointnet2_utils
                sys.modules["pointnet2_ops.pointnet2_modules"] = p2.pointnet2_modules
            # else: the reference's own pointnet2_ops package (pointnet2_ops_lib on sys.path) imports normally and its
            # `import pointnet2_ops._ext` resolves to the entry registered above instead of JIT-compiling
        if pkg is not None:
            pkg._ext = p2._ext
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
