"""stereo_b200 -- B200-native (sm_100a) hot path of johannesu/stereo.

Host-side mirror of the reference's MATLAB surface over the C ABI of
libstereo_b200.so (include/stereo_b200.h): ``trws`` / ``rd`` (trws.m, rd.m) and the
``dispmap_*`` classes.  All compute runs in hand-written CUDA kernels.
"""
from . import _lib  # noqa: F401
from .solvers import trws, rd, binary_fusion_grid, binary_fuse_until_convergence_grid, TrwsSolver, trws_grid_ordering, grid_from_connectivity  # noqa: F401
from .grid import construct_neighborhood, get_points  # noqa: F401
from .gridsolver import TrwsGrid, trws_grid, positions_from_labels  # noqa: F401
from . import builders  # noqa: F401
from .dispmap import dispmap_super, dispmap_ncc, dispmap_globalstereo  # noqa: F401

__all__ = ["trws", "rd", "binary_fusion_grid", "binary_fuse_until_convergence_grid", "TrwsSolver", "TrwsGrid", "trws_grid", "dispmap_super", "dispmap_ncc", "dispmap_globalstereo", "builders", "trws_grid_ordering", "grid_from_connectivity", "construct_neighborhood", "get_points"]
