"""oracle.c_port -- ctypes loader of ``oracle/wbc_oracle.c`` (plain-C restatement of one WBC tick).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the C port is the fast CPU checker (every state of a 4096-state
batch in well under a second) and the CPU baseline bench.py times.  It takes the same marshalled ``WbcTreeTable`` /
``WbcConfig`` structs as the CUDA library (include/wbc_b200.h), so both sides are driven by identical inputs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "wbc_oracle.c")
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "libwbc_oracle.so")

_lib = None


def build(force=False):
    """gcc -O3 -march=native -fopenmp -shared -fPIC -> oracle/_build/libwbc_oracle.so"""
    hdr = os.path.join(ROOT, "include", "wbc_b200.h")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return LIB
    os.makedirs(BUILD_DIR, exist_ok=True)
    # compile to a private name and rename: several processes (the bench's CPU-baseline workers, pytest-xdist) may find
    # the library stale at the same time, and a reader must never see a half-written file
    tmp = f"{LIB}.{os.getpid()}.tmp"
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"), SRC,
           "-o", tmp, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.wbc_oracle_step.restype = C.c_int
        _lib.wbc_oracle_threads.restype = C.c_int
    return _lib


def cabi():
    import wbc_b200
    return wbc_b200._cabi


def table_struct(name_or_table):
    """TreeTable (or its name) -> (WbcTreeTable ctypes struct, TreeTable), with the hot-frame slot order of RobotModel."""
    import wbc_b200
    from wbc_b200 import robot_model as prm
    t = name_or_table if isinstance(name_or_table, wbc_b200.TreeTable) else wbc_b200.TreeTable.load(name_or_table)
    slots = [t.getFrameId(n, "FIXED_JOINT") for n in prm.EE_FRAME_NAMES] + [t.getFrameId("imu_joint", "FIXED_JOINT")]
    return prm.RobotModel._make_table(t, slots), t


def config_struct(rm, table):
    """WbcConfig from an object carrying the reference's attribute names (oracle or product RobotModel)."""
    ab = cabi()
    c = ab.WbcConfig()
    m = 0
    for flag, bit in ((rm.task_active_FR_foot, ab.TASK_FR), (rm.task_active_FL_foot, ab.TASK_FL),
                      (rm.task_active_RR_foot, ab.TASK_RR), (rm.task_active_RL_foot, ab.TASK_RL),
                      (rm.task_active_GRIP, ab.TASK_GRIP), (rm.task_active_Trunk, ab.TASK_TRUNK)):
        if flag is True:
            m |= bit
    j = rm.task_active_Joint
    if j is True or (isinstance(j, str) and j in ("PREV", "MANI", "HYBRID")):
        m |= ab.TASK_JOINT
    c.task_mask = m
    c.joint_mode = {"PREV": ab.JOINT_PREV, "MANI": ab.JOINT_MANI, "HYBRID": ab.JOINT_HYBRID}.get(j, ab.JOINT_ZERO) \
        if isinstance(j, str) else ab.JOINT_ZERO
    k = 0
    for flag, bit in ((rm.const_active_CoM, ab.CON_COM), (rm.const_active_Trunk, ab.CON_TRUNK),
                      (rm.const_active_FR_foot, ab.CON_FR), (rm.const_active_FL_foot, ab.CON_FL),
                      (rm.const_active_RR_foot, ab.CON_RR), (rm.const_active_RL_foot, ab.CON_RL),
                      (rm.const_active_GRIP, ab.CON_GRIP)):
        if flag is True:
            k |= bit
    c.constraint_mask = k
    c.compat_flags = ab.COMPAT_DAMPER_OFF_BY_ONE if getattr(rm, "compat_damper_off_by_one", True) else 0
    c.gripper_joint_id = table.getJointId("gripper")
    c.arm_base_id = table.getJointId("waist")
    c.max_iter = int(getattr(rm, "max_qp_iterations", 200))
    for i in range(5):
        W = np.asarray(rm.EE_weight[i], dtype=float).reshape(-1)
        G = np.asarray(rm.EE_gains[i], dtype=float)[0:3, 0:3].reshape(-1)
        for e in range(36):
            c.ee_weight[i][e] = W[e]
        for e in range(9):
            c.ee_gain_pos[i][e] = G[e]
        c.cart_task_weight[i] = float(rm.cart_task_weight_EE_list[i])
    W = np.asarray(rm.trunk_weight, dtype=float).reshape(-1)
    for e in range(36):
        c.trunk_weight[e] = W[e]
    c.cart_task_weight[5] = float(rm.cart_task_weight_Trunk)
    c.joint_task_weight = float(rm.joint_task_weight)
    G = np.asarray(rm.trunk_gain, dtype=float)
    for e in range(9):
        c.trunk_gain_pos[e] = G[0:3, 0:3].reshape(-1)[e]
    for e in range(3):
        c.trunk_gain_ori[e] = G[3 + e, 3 + e]
    c.damper_coef, c.damper_qi, c.damper_qs = getattr(rm, "damper", (0.01, 0.026, 0.015))
    rows = getattr(rm, "extra_rows", [])
    c.n_extra_rows = len(rows)
    for e, (slot, rf, coeff, lo, hi) in enumerate(rows):
        c.extra_frame[e], c.extra_rf[e], c.extra_lo[e], c.extra_hi[e] = int(slot), int(rf), float(lo), float(hi)
        for r in range(6):
            c.extra_coeff[e][r] = float(coeff[r])
    return c


def step(table, cfg, q, targets, mem, ref, dt, nthreads=0, want_Ab=False):
    """One open-loop tick for N states on the CPU.  Returns dict(qdot, status, iters, active_set, mem_out[, A, b])."""
    lib = load()
    if nthreads <= 0:                      # all the cores this process may use (OMP_NUM_THREADS is often pinned to 1)
        nthreads = len(os.sched_getaffinity(0))
    q = np.ascontiguousarray(q, dtype=np.float64)
    targets = np.ascontiguousarray(targets, dtype=np.float64)
    mem = np.ascontiguousarray(mem, dtype=np.float64)
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    N, nv = q.shape[0], table.nv
    m = sum(6 for t in range(6) if cfg.task_mask & (1 << t)) + (nv if cfg.task_mask & 64 else 0)
    out = {"qdot": np.zeros((N, nv)), "status": np.zeros(N, dtype=np.int32), "iters": np.zeros(N, dtype=np.int32),
           "active_set": np.zeros((N, 2), dtype=np.uint64), "mem_out": np.zeros((N, 72))}
    A = np.zeros((N, m, nv)) if want_Ab else None
    b = np.zeros((N, m)) if want_Ab else None
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(None)
    rc = lib.wbc_oracle_step(C.byref(table), C.byref(cfg), p(q), p(targets), p(mem), p(ref), C.c_double(dt),
                             C.c_int64(N), C.c_int(nthreads), p(out["qdot"]), p(out["status"]), p(out["iters"]),
                             p(out["active_set"]), p(out["mem_out"]), p(A), p(b))
    if rc != 0:
        raise RuntimeError(f"wbc_oracle_step failed with code {rc}")
    if want_Ab:
        out["A"], out["b"] = A, b
    return out


def max_threads():
    return load().wbc_oracle_threads()
