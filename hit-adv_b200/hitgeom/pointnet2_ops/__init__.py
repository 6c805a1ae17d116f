"""hitgeom.pointnet2_ops -- drop-in for the reference's `pointnet2_ops` package
(pointnet2_ops_lib/pointnet2_ops): `_ext` (the native module), `pointnet2_utils` (autograd wrappers) and
`pointnet2_modules` (set-abstraction / feature-propagation modules)."""
from . import _ext  # noqa: F401
from . import ops  # noqa: F401
from . import ops as pointnet2_utils  # noqa: F401  (reference module name)
from . import pointnet2_modules  # noqa: F401,E402
