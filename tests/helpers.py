"""Test helpers: run the CPU oracle (oracle/) on the same arrays the CUDA path consumed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pin as opin  # noqa: E402
from oracle.robot_wrapper4 import RobotModel as OracleRobotModel  # noqa: E402
from oracle.qp_wrapper import QP as OracleQP, kkt_residuals  # noqa: E402

PKG_DATA = os.path.join(ROOT, "mech5845m-wbc-for-legged-manipulator_b200", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def table_path(name):
    return os.path.join(PKG_DATA, name + ".json")


def oracle_model(name):
    return opin.Model.from_json(table_path(name))


def golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def make_oracle(name, like=None, dt=0.002):
    """Oracle RobotModel with the switches / weights of `like` (a product or oracle RobotModel)."""
    rm = OracleRobotModel(oracle_model(name), dt=dt)
    if like is not None:
        copy_settings(like, rm)
    return rm


def copy_settings(src, dst):
    for a in ("task_active_Trunk", "task_active_FR_foot", "task_active_FL_foot", "task_active_RR_foot",
              "task_active_RL_foot", "task_active_GRIP", "task_active_Joint", "const_active_CoM",
              "const_active_Trunk", "const_active_FR_foot", "const_active_FL_foot", "const_active_RR_foot",
              "const_active_RL_foot", "const_active_GRIP", "trunk_weight", "EE_weight", "cart_task_weight_EE_list",
              "cart_task_weight_Trunk", "joint_task_weight", "trunk_gain", "EE_gains", "dt"):
        setattr(dst, a, getattr(src, a))
    for a in ("extra_rows", "compat_damper_off_by_one", "max_qp_iterations"):
        if hasattr(src, a):
            setattr(dst, a, getattr(src, a))


def set_oracle_state(rm, q, mem, ref):
    """Load one state (configuration + task memory + references) into the oracle object."""
    q = np.array(q, dtype=float)
    rm.current_joint_config = q
    rm.updateState(q, feedback=False)
    rm.prev_EE_pos = [mem[3 * i:3 * i + 3].copy() for i in range(5)]
    rm.prev_EE_CoM_rot = [mem[15 + 9 * i:24 + 9 * i].reshape(3, 3).copy() for i in range(5)]
    rm.prev_trunk_ref = mem[60:63].copy()
    rm.old_ref_trunk_rot_matrix = mem[63:72].reshape(3, 3).copy()
    rm.default_EE_ori_list = [ref[3 * i:3 * i + 3].reshape(3, 1).copy() for i in range(5)]
    rm.default_trunk_ori = ref[15:18].reshape(3, 1).copy()
    rm.initial_trunk_pos = ref[18:21].copy()
    rm.initial_trunk_ori_euler = ref[21:24].reshape(3, 1).copy()
    rm.firstQP = True


def get_oracle_mem(rm):
    return np.concatenate([np.concatenate([np.reshape(p, 3) for p in rm.prev_EE_pos]),
                           np.concatenate([np.reshape(r, 9) for r in rm.prev_EE_CoM_rot]),
                           np.reshape(rm.prev_trunk_ref, 3), np.reshape(rm.old_ref_trunk_rot_matrix, 9)])


def oracle_step_one(rm, q, targets, mem, ref, imu=None, solve=True, tail=True):
    set_oracle_state(rm, q, mem, ref)
    ee = [targets[3 * i:3 * i + 3].reshape(3, 1).copy() for i in range(5)]
    tr = targets[15:18].reshape(3, 1).copy()
    rm.FR_target_cartesian_pos, rm.FL_target_cartesian_pos = ee[0], ee[1]
    rm.RR_target_cartesian_pos, rm.RL_target_cartesian_pos = ee[2], ee[3]
    out = {}
    A = rm.qpA()
    b = rm.qpb(ee, tr).reshape((A.shape[0],))
    any_con = any([rm.const_active_CoM, rm.const_active_Trunk, rm.const_active_FR_foot, rm.const_active_FL_foot,
                   rm.const_active_RR_foot, rm.const_active_RL_foot, rm.const_active_GRIP])
    if any_con:
        Ct, Clb, Cub = rm.findConstraints()
        Cm = Ct.T
    else:
        Ct = Clb = Cub = None
        Cm = np.zeros((0, A.shape[1]))
        Clb = Cub = None
    extra = getattr(rm, "extra_rows", None)
    if extra:                              # extension rows (NOT in the reference): coeff . J_frame(rf), appended to C
        Ce, le, ue = oracle_extra_rows(rm, extra)
        Cm = np.concatenate((Cm, Ce), axis=0)
        Clb = le if Clb is None else np.concatenate((Clb, le))
        Cub = ue if Cub is None else np.concatenate((Cub, ue))
        Ct = Cm.T
    lb, ub = rm.velDamperJointConstraints()
    out.update(A=A, b=b, lb=lb, ub=ub, C=Cm, Clb=np.zeros(0) if Clb is None else Clb,
               Cub=np.zeros(0) if Cub is None else Cub, mem_out=get_oracle_mem(rm))
    out["H"] = A.T @ A
    out["g"] = -A.T @ b
    if solve:
        qp = OracleQP(A, b, lb, ub, Ct, Clb, Cub, n_of_velocity_dimensions=rm.n_velocity_dimensions)
        qp.max_iter = int(getattr(rm, "max_qp_iterations", 200))
        x = qp.solveQP()
        out.update(qdot=np.array(x), status=qp.result["status"], iters=qp.result["iters"], act=qp.result["act"])
        if tail:
            joint_config = rm.jointVelocitiestoConfig(x, False)
            full = np.array(joint_config)
            base = np.array(imu, dtype=float) if imu is not None else full[3:7]
            rm.updateState(full[7:], base, running=True)
            out["q_next"] = np.array(rm.current_joint_config)
            out["q_integrated"] = full
    return out


def oracle_extra_rows(rm, rows):
    """The generic extension-row channel of WbcConfig (extra_frame / extra_rf / extra_coeff / extra_lo / extra_hi):
    row = coeff . getFrameJacobian(hot frame slot, rf).  Slots 0..4 = EE frames, 5 = trunk."""
    frames = rm.end_effector_index_list_frame + [rm.trunk_frame_index]
    Ce, lo, hi = [], [], []
    for slot, rf, coeff, l, h in rows:
        J = opin.getFrameJacobian(rm.robot_model, rm.robot_data, frames[int(slot)], int(rf))
        Ce.append(np.asarray(coeff, dtype=float) @ J)
        lo.append(float(l)); hi.append(float(h))
    return np.array(Ce), np.array(lo), np.array(hi)


def oracle_step_batch(name, like, q, targets, mem, ref, imu=None, solve=True, tail=False):
    """Loop the oracle over a batch; returns stacked arrays."""
    rm = make_oracle(name, like=like, dt=like.dt)
    res = []
    for s in range(q.shape[0]):
        res.append(oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], imu=None if imu is None else imu[s],
                                   solve=solve, tail=tail))
    keys = res[0].keys()
    return {k: np.stack([np.asarray(r[k]) for r in res]) for k in keys}


def act_to_bits(act, nv):
    """Oracle act codes (0 none, 1 lower, 2 upper, 3 eq) -> (word_box, word_rows) like the kernel's active_set."""
    wb = wr = 0
    for c, a in enumerate(act):
        lo, up = int(a) & 1, (int(a) >> 1) & 1
        if c < nv:
            wb |= (lo << (2 * c)) | (up << (2 * c + 1))
        else:
            r = c - nv
            wr |= (lo << (2 * r)) | (up << (2 * r + 1))
    return wb, wr


__all__ = ["oracle_model", "make_oracle", "oracle_step_batch", "oracle_step_one", "kkt_residuals", "golden",
           "table_path", "act_to_bits", "REFERENCE", "copy_settings", "set_oracle_state"]
