"""B200-native lift-splat (camera -> BEV) for E2E Parking.

Drop-in for ``model/bev_model.py`` of qintonguav/e2e-parking-carla: ``BevModel`` keeps
the reference's constructor, parameters and ``forward/calc_bev_feature`` contract and
runs the projection in hand-written sm_100a CUDA kernels (``libls_b200.so``, C ABI in
``include/ls_b200.h``).  The package directory is spelled with underscores because
Python cannot import a hyphenated name.
"""
from .bev_model import BevModel, calculate_birds_eye_view_parameters  # noqa: F401
from . import lift_splat  # noqa: F401  (module: lift_splat.lift_splat(...) is the functional entry point)
from .lift_splat import GridSpec, LiftSplatFunction  # noqa: F401
from .depth_loss import DepthLoss  # noqa: F401
from .target_bev import add_target_bev  # noqa: F401

__all__ = ["BevModel", "calculate_birds_eye_view_parameters", "GridSpec", "LiftSplatFunction", "lift_splat", "DepthLoss", "add_target_bev"]
