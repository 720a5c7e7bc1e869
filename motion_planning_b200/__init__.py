"""motion_planning_b200 -- B200-native MPPI rollout engine behind the reference's `MPPI` class surface.

Only the hot path of moribots/motion_planning's control/src/mppi is rebuilt here:
hand-written sm_100a kernels in csrc/, a C ABI (include/mppi_b200.h), and this Python mirror of the
reference interface.  There is no CPU implementation in this package.
"""
from .mppi import MPPI, UserModel, KinematicModel, rk4, euler, bicycle_rk4, WHEEL_VEL_MAX, WHEEL_RADIUS, WHEEL_BASE  # noqa: F401
from ._capi import MppiError  # noqa: F401
from .controller import Controller, FakeDiffDrive, waypoints_from_path, lookahead_goal  # noqa: F401

__all__ = ["MPPI", "UserModel", "KinematicModel", "rk4", "euler", "bicycle_rk4", "MppiError", "WHEEL_VEL_MAX", "WHEEL_RADIUS", "WHEEL_BASE",
           "Controller", "FakeDiffDrive", "waypoints_from_path", "lookahead_goal"]
