"""B200-native point-geometry hot path of SVDFormer / PointSea.

Public surface = the reference's op boundary:
    chamfer_3DDist, chamfer_3DFunction                       (metrics/CD/chamfer3D/dist_chamfer_3D.py)
    furthest_point_sample, gather_operation, grouping_operation, ball_query,
    three_nn, three_interpolate, QueryAndGroup, GroupAll      (pointnet2_ops/pointnet2_utils.py)
    query_knn, fps_subsample, query_knn_point, index_points, group_local, sample_and_group_knn,
    EdgeConv (+ edge_features)                                (models/model_utils.py)
    calc_cd, calc_dcd, fscore                                 (utils/loss_utils.py, metrics/CD/fscore.py)
`install_dropin()` makes `import metrics.CD.chamfer3D.dist_chamfer_3D` and
`import pointnet2_ops.pointnet2_utils` resolve to this package so the reference's models and
losses run unchanged.
"""
from ._lib import PointSeaError, LIB_PATH, load as load_library  # noqa: F401
from .chamfer import chamfer_3DDist, chamfer_3DFunction, chamfer_forward, chamfer_backward, chamfer_sums, chamfer_host, chamfer_host_async, HostStep, chamfer_host_step, ChamferStep  # noqa: F401
from .pointnet2_utils import (  # noqa: F401
    furthest_point_sample, gather_operation, grouping_operation, ball_query, three_nn, three_interpolate,
    FurthestPointSampling, GatherOperation, GroupingOperation, BallQuery, ThreeNN, ThreeInterpolate,
    QueryAndGroup, GroupAll, query_knn, fps_subsample,
)
from .model_ops import (  # noqa: F401
    query_knn_point, index_points, group_local, edge_features, EdgeConv, sample_and_group_knn, knn_self, patch_model_utils,
)
from .metrics import calc_cd, calc_dcd, fscore, chamfer_metrics_raw, patch_loss_utils  # noqa: F401
from .dropin import install_dropin, DROPIN_PATH  # noqa: F401

__version__ = "0.1.0"
