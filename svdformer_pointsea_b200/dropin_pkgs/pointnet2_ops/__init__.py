# The reference's pointnet2_ops/__init__.py:1 also imports pointnet2_modules (unused by every
# model); only the utils surface is shadowed here.
from . import pointnet2_utils  # noqa: F401
