"""B200-native point-geometry hot path of SVDFormer / PointSea.

Public surface = the reference's op boundary:
    chamfer_3DDist, chamfer_3DFunction                       (metrics/CD/chamfer3D/dist_chamfer_3D.py)
    furthest_point_sample, gather_operation, grouping_operation, ball_query,
    three_nn, three_interpolate, QueryAndGroup, GroupAll      (pointnet2_ops/pointnet2_utils.py)
    query_knn, fps_subsample                                  (models/model_utils.py)
`install_dropin()` makes `import metrics.CD.chamfer3D.dist_chamfer_3D` and
`import pointnet2_ops.pointnet2_utils` resolve to this package so the reference's models and
losses run unchanged.
"""
from ._lib import PointSeaError, LIB_PATH, load as load_library  # noqa: F401
from .chamfer import chamfer_3DDist, chamfer_3DFunction, chamfer_forward, chamfer_backward, chamfer_sums, chamfer_host  # noqa: F401
from .pointnet2_utils import (  # noqa: F401
    furthest_point_sample, gather_operation, grouping_operation, ball_query, three_nn, three_interpolate,
    FurthestPointSampling, GatherOperation, GroupingOperation, BallQuery, ThreeNN, ThreeInterpolate,
    QueryAndGroup, GroupAll, query_knn, fps_subsample,
)
from .dropin import install_dropin, DROPIN_PATH  # noqa: F401

__version__ = "0.1.0"
