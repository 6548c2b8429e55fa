"""ctypes binding of libwbc_b200.so (include/wbc_b200.h).  No torch types cross this boundary:
only raw device pointers (``tensor.data_ptr()``), sizes and a ``cudaStream_t``.

The library is built in-tree by ``__graft_entry__.build()`` (or ``build_library()`` below) with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no CPU fallback: if the library is
missing, ``load()`` raises, and every compute entry point fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("WBC_B200_LIB") or os.path.join(PKG_DIR, "libwbc_b200.so")   # override: A/B builds only
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include", "wbc_b200.h")

MAX_JOINTS, MAX_NV, MAX_NQ, MAX_FRAMES, NUM_EE, MAX_NC, MAX_EXTRA = 32, 32, 33, 16, 5, 32, 24
FRAME_TRUNK = 5
TARGETS_STRIDE, MEM_STRIDE, REF_STRIDE = 18, 72, 24
RF_WORLD, RF_LOCAL, RF_LOCAL_WORLD_ALIGNED = 0, 1, 2
TASK_FR, TASK_FL, TASK_RR, TASK_RL, TASK_GRIP, TASK_TRUNK, TASK_JOINT = 1, 2, 4, 8, 16, 32, 64
JOINT_ZERO, JOINT_PREV, JOINT_MANI, JOINT_HYBRID = 0, 1, 2, 3
CON_COM, CON_TRUNK, CON_FR, CON_FL, CON_RR, CON_RL, CON_GRIP = 1, 2, 4, 8, 16, 32, 64
COMPAT_DAMPER_OFF_BY_ONE = 1
QP_SOLVED, QP_MAXITER, QP_INFEASIBLE, QP_NOT_PD = 0, 1, 2, 4
STEP_FLAG_PLAIN_INTEGRATE = 1
HOST_F64, HOST_F32 = 0, 1
HOST_FLAG_DELTA_INPUTS = 1
ABI_VERSION = 2
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
JT_UNIVERSE, JT_FREEFLYER, JT_REVOLUTE, JT_PRISMATIC = 0, 1, 2, 3

EXPORTS = [
    "wbc_abi_version", "wbc_last_error", "wbc_model_create", "wbc_model_destroy", "wbc_config_rows",
    "wbc_fk_jac", "wbc_joint_jacobians", "wbc_init_memory", "wbc_integrate", "wbc_base_estimate", "wbc_assemble",
    "wbc_qp_solve", "wbc_step",
    "wbc_rollout", "wbc_step_host", "wbc_step_host_path",
    "wbc_step_launch_info", "wbc_measure_fp64_peak",
]

i32, f64 = C.c_int32, C.c_double


class WbcTreeTable(C.Structure):
    _fields_ = [
        ("njoints", i32), ("nq", i32), ("nv", i32), ("nframes", i32),
        ("parent", i32 * MAX_JOINTS), ("jtype", i32 * MAX_JOINTS),
        ("idx_q", i32 * MAX_JOINTS), ("idx_v", i32 * MAX_JOINTS),
        ("placement_R", (f64 * 9) * MAX_JOINTS), ("placement_p", (f64 * 3) * MAX_JOINTS),
        ("axis", (f64 * 3) * MAX_JOINTS),
        ("frame_parent", i32 * MAX_FRAMES),
        ("frame_R", (f64 * 9) * MAX_FRAMES), ("frame_p", (f64 * 3) * MAX_FRAMES),
        ("lower", f64 * MAX_NQ), ("upper", f64 * MAX_NQ), ("velocity", f64 * MAX_NV),
        ("mass", f64 * MAX_JOINTS), ("com", (f64 * 3) * MAX_JOINTS),
    ]


class WbcConfig(C.Structure):
    _fields_ = [
        ("task_mask", i32), ("joint_mode", i32), ("constraint_mask", i32), ("compat_flags", i32),
        ("gripper_joint_id", i32), ("arm_base_id", i32), ("max_iter", i32), ("n_extra_rows", i32),
        ("ee_weight", (f64 * 36) * NUM_EE), ("trunk_weight", f64 * 36),
        ("cart_task_weight", f64 * 6), ("joint_task_weight", f64),
        ("ee_gain_pos", (f64 * 9) * NUM_EE), ("trunk_gain_pos", f64 * 9), ("trunk_gain_ori", f64 * 3),
        ("damper_coef", f64), ("damper_qi", f64), ("damper_qs", f64),
        ("extra_frame", i32 * MAX_EXTRA), ("extra_rf", i32 * MAX_EXTRA),
        ("extra_coeff", (f64 * 6) * MAX_EXTRA), ("extra_lo", f64 * MAX_EXTRA), ("extra_hi", f64 * MAX_EXTRA),
    ]


class WbcStepIO(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("targets", C.c_void_p), ("mem_in", C.c_void_p), ("ref", C.c_void_p),
        ("imu_quat", C.c_void_p), ("dt", f64), ("flags", C.c_int64),
        ("qdot", C.c_void_p), ("status", C.c_void_p), ("iters", C.c_void_p), ("active_set", C.c_void_p),
        ("mem_out", C.c_void_p), ("q_next", C.c_void_p), ("joint_targets", C.c_void_p),
    ]


class WbcHostIO(C.Structure):
    _fields_ = [("q", C.c_void_p), ("targets", C.c_void_p), ("mem_in", C.c_void_p), ("ref", C.c_void_p),
                ("imu_quat", C.c_void_p), ("qdot", C.c_void_p), ("status", C.c_void_p), ("iters", C.c_void_p),
                ("joint_targets", C.c_void_p), ("dtype", i32), ("flags", i32)]


class WbcAssembleOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("A", "b", "lb", "ub", "C", "Clb", "Cub", "H", "g")]


class WbcError(RuntimeError):
    pass


_lib = None


def nvcc_command(out_path=LIB_PATH, extra=()):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", *extra, "-o", out_path, os.path.join(CSRC, "wbc_kernels.cu")]


def build_library(force=False, verbose=False):
    """Compile csrc/ into libwbc_b200.so for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [INCLUDE]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = nvcc_command(extra=("-Xptxas", "-v") if verbose else ())
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise WbcError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


def load():
    """dlopen the library and declare every prototype of include/wbc_b200.h.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WbcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    lib.wbc_abi_version.restype = C.c_int
    lib.wbc_last_error.restype = C.c_char_p
    lib.wbc_model_create.argtypes = [C.POINTER(WbcTreeTable), C.POINTER(vp)]
    lib.wbc_model_destroy.argtypes = [vp]
    lib.wbc_model_destroy.restype = None
    lib.wbc_config_rows.argtypes = [C.POINTER(WbcConfig), i32, C.POINTER(i32), C.POINTER(i32)]
    lib.wbc_fk_jac.argtypes = [vp, vp, i64, C.POINTER(i32), i32, i32, vp, vp, vp]
    lib.wbc_joint_jacobians.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.wbc_init_memory.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.wbc_integrate.argtypes = [vp, vp, vp, i64, f64, vp, vp]
    lib.wbc_base_estimate.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    lib.wbc_assemble.argtypes = [vp, C.POINTER(WbcConfig), C.POINTER(WbcStepIO), i64, C.POINTER(WbcAssembleOut), vp]
    lib.wbc_qp_solve.argtypes = [i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.wbc_step.argtypes = [vp, C.POINTER(WbcConfig), C.POINTER(WbcStepIO), i64, vp]
    lib.wbc_step_host.argtypes = [vp, C.POINTER(WbcConfig), C.POINTER(WbcStepIO), C.POINTER(WbcHostIO), i64, i32, vp]
    lib.wbc_step_host_path.argtypes = [vp]
    lib.wbc_rollout.argtypes = [vp, C.POINTER(WbcConfig), C.POINTER(WbcStepIO), vp, vp, i32, i64, vp]
    lib.wbc_step_launch_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.wbc_measure_fp64_peak.argtypes = [C.POINTER(f64), vp]
    for name in EXPORTS:
        if name not in ("wbc_last_error", "wbc_model_destroy"):
            getattr(lib, name).restype = C.c_int
    if lib.wbc_abi_version() != ABI_VERSION:
        raise WbcError("libwbc_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise WbcError(f"libwbc_b200 error {rc}: {load().wbc_last_error().decode()}")
