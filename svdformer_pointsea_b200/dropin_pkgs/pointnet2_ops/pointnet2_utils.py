"""Drop-in for pointnet2_ops/pointnet2_utils.py (reference :34-379): every public name."""
from svdformer_pointsea_b200.pointnet2_utils import (  # noqa: F401
    FurthestPointSampling, furthest_point_sample,
    GatherOperation, gather_operation,
    ThreeNN, three_nn,
    ThreeInterpolate, three_interpolate,
    GroupingOperation, grouping_operation,
    BallQuery, ball_query,
    QueryAndGroup, GroupAll,
)
