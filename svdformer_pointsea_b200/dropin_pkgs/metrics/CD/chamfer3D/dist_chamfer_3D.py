"""Drop-in for metrics/CD/chamfer3D/dist_chamfer_3D.py (reference :26-74): same two names."""
from svdformer_pointsea_b200.chamfer import chamfer_3DFunction, chamfer_3DDist  # noqa: F401
