"""B200-native batched whole-body-control hot path (FK + frame Jacobians -> task stack -> QP).

Drop-in for the per-step path of joey156/MECH5845M-WBC-for-Legged-Manipulator:
``RobotModel`` mirrors ``wrappers/Robot_Wrapper4.py`` and ``QP`` mirrors ``wrappers/QP_Wrapper.py``,
both over a batch of N robot states in float64 CUDA tensors.  The arithmetic lives in
``libwbc_b200.so`` (hand-written sm_100a kernels behind the C ABI of ``include/wbc_b200.h``).

The directory name contains hyphens, so import it through the ``wbc_b200`` shim at the repo root:
``from wbc_b200 import RobotModel, QP``.
"""
from . import _cabi
from ._cabi import WbcError, build_library, load as load_library
from .tree_table import TreeTable
from .robot_model import RobotModel, LinearTrajectory, HostDeltaEncoder, EE_FRAME_NAMES, EE_JOINT_NAMES, HIP_WAIST_JOINT_NAMES
from .qp import QP
from . import synthetic
from . import sharding
from . import mocap

__all__ = ["RobotModel", "LinearTrajectory", "HostDeltaEncoder", "mocap", "QP", "TreeTable", "WbcError", "build_library", "load_library", "synthetic", "sharding",
           "EE_FRAME_NAMES", "EE_JOINT_NAMES", "HIP_WAIST_JOINT_NAMES"]
