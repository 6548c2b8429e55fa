"""Synthetic (q, target, memory) batches for benchmarks and parity tests.

Everything random is drawn on the host with ``numpy.random.Generator(PCG64(seed))`` as GLOBAL arrays of
N states, then sliced per rank, so results do not depend on the number of GPUs:

  base xyz ~ U([-0.5, 0.5]^2 x [0.2, 0.5]);  base quaternion = from_euler('xyz', U(-0.3, 0.3)^3)  (x, y, z, w)
  joints   ~ U(lower + 5 % range, upper - 5 % range) from the URDF limits; gripper = 0, fingers = (+0.02, -0.02)
  targets  = FK position + N(0, sigma^2)   (sigma = 5e-4 nominal, 5e-3 "stress": bounds start to bind)
  memory   = previous targets at the FK positions, previous reference rotations at the FK rotations,
             default Euler angles / initial trunk pose from FK at the sampled state (initialiseWBC semantics)

FK for the targets is evaluated on the GPU by the product kernels (``RobotModel.initialiseWBC``).
"""
from __future__ import annotations

import numpy as np
import torch

SEED_BASE = 20260000


def _quat_from_euler_xyz(e):
    """Extrinsic xyz Euler angles -> (x, y, z, w) = qz * (qy * qx), vectorised."""
    h = 0.5 * e
    sx, cx, sy, cy, sz, cz = np.sin(h[:, 0]), np.cos(h[:, 0]), np.sin(h[:, 1]), np.cos(h[:, 1]), np.sin(h[:, 2]), np.cos(h[:, 2])
    # qy * qx
    ax, ay, az, aw = cy * sx, sy * cx, -sy * sx, cy * cx
    # qz * (.)
    x = cz * ax - sz * ay
    y = cz * ay + sz * ax
    z = cz * az + sz * aw
    w = cz * aw - sz * az
    return np.stack([x, y, z, w], axis=1)


def sample_configurations(table, N, seed, gripper_joint="gripper"):
    """[N, nq] float64 configurations inside the URDF limits."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nq = table.nq
    q = np.zeros((N, nq))
    q[:, 0:2] = rng.uniform(-0.5, 0.5, size=(N, 2))
    q[:, 2] = rng.uniform(0.2, 0.5, size=N)
    q[:, 3:7] = _quat_from_euler_xyz(rng.uniform(-0.3, 0.3, size=(N, 3)))
    lo = np.asarray(table.lower[7:], dtype=float)
    up = np.asarray(table.upper[7:], dtype=float)
    u = rng.uniform(size=(N, nq - 7))
    q[:, 7:] = lo + 0.05 * (up - lo) + u * 0.9 * (up - lo)
    g = table.getJointId(gripper_joint)
    if g < table.njoints:
        iq = table.idx_q[g]
        q[:, iq] = 0.0
        if iq + 2 < nq:
            q[:, iq + 1] = 0.02
            q[:, iq + 2] = -0.02
    return q


def sample_standing(table, N, seed, joint_sigma=0.1, tilt=0.03, gripper_joint="gripper"):
    """[N, nq] configurations around a standing pose with a near-level trunk: the closed-loop workloads.

    The reference's base estimator rotates an already world-frame offset by the trunk rotation once more
    (``trunkWorldPos``, Robot_Wrapper4.py:1303-1324), so its closed loop only makes sense for a trunk
    that is close to the identity orientation -- as in its own simulation (``sim3.py``: a standing robot facing +x).
    With the +-0.3 rad random base attitudes of ``sample_configurations`` every tick displaces the estimated base by
    centimetres and the arm runs into its velocity limits.  Here: base x, y ~ U(-0.5, 0.5), z = 0.3, attitude
    from_euler('xyz', U(-tilt, tilt)^3); legs around the standing pattern of ``Robot_Wrapper.py:26-28`` (hip 0, thigh 0.85,
    calf -1.15), arm joints around 40 % of their range, all + N(0, joint_sigma^2), clipped 5 % inside the limits;
    gripper = 0, fingers = (+0.02, -0.02)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nq = table.nq
    q = np.zeros((N, nq))
    q[:, 0:2] = rng.uniform(-0.5, 0.5, size=(N, 2))
    q[:, 2] = 0.3
    q[:, 3:7] = _quat_from_euler_xyz(rng.uniform(-tilt, tilt, size=(N, 3)))
    lo = np.asarray(table.lower[:nq], dtype=float)
    up = np.asarray(table.upper[:nq], dtype=float)
    for j in range(2, table.njoints):
        iq, name = table.idx_q[j], table.joint_names[j]
        if name.endswith("_hip_joint"):
            mean = 0.0
        elif name.endswith("_thigh_joint"):
            mean = 0.85
        elif name.endswith("_calf_joint"):
            mean = -1.15
        else:
            mean = lo[iq] + 0.4 * (up[iq] - lo[iq])
        margin = 0.05 * (up[iq] - lo[iq])
        q[:, iq] = np.clip(mean + rng.normal(0.0, joint_sigma, size=N), lo[iq] + margin, up[iq] - margin)
    g = table.getJointId(gripper_joint)
    if g < table.njoints:
        iq = table.idx_q[g]
        q[:, iq] = 0.0
        if iq + 2 < nq:
            q[:, iq + 1] = 0.02
            q[:, iq + 2] = -0.02
    return q


def sample_noise(N, seed, sigma):
    """[N, 18] target offsets: 5 EE x 3 + trunk 3."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    return rng.normal(0.0, sigma, size=(N, 18))


def load_batch(robot, q, noise):
    """Put a sampled batch onto a RobotModel: state, memory snapshot, targets.  Returns targets [N, 18] (CUDA)."""
    dev = robot.device
    qd = torch.as_tensor(q, dtype=torch.float64, device=dev).contiguous()
    robot.current_joint_config = qd
    robot.initialiseWBC(qd[:, 3:7])
    # previous reference rotations at the current rotations => zero feed-forward angular velocity
    oMf = robot._oMf
    robot._mem[:, 15:60] = oMf[:, 0:5, 0:9].reshape(robot.N, 45)
    robot._mem[:, 63:72] = oMf[:, 5, 0:9]
    targets = torch.cat([oMf[:, i, 9:12] for i in range(5)] + [oMf[:, 5, 9:12]], dim=1)
    targets = targets + torch.as_tensor(noise, dtype=torch.float64, device=dev)
    return targets.contiguous()


def config3_rows(table, mu=0.6, duty=0.005, payload=0.2, gravity=9.81):
    """Constraint rows of BASELINE config 3 ("A1 + WX200 with friction-cone + torque-limit constraints") for the generic
    extension-row channel (``RobotModel.extra_rows`` -> ``WbcConfig.extra_*``): tuples (frame slot, reference frame,
    coeff[6], lo, hi), row = coeff . J_frame(rf), lo <= row . qdot <= hi.

    NOT in the reference: its QP is purely kinematic (no contact forces, no torques), so both
    families are velocity-space stand-ins, used with the four foot equality constraints switched OFF:

      * friction pyramid, 4 faces per foot (16 rows): the LOCAL_WORLD_ALIGNED foot velocity v stays inside
        +-v_x <= mu v_z, +-v_y <= mu v_z (a contact may only lift off inside the cone; v = 0 is the apex);
      * torque-limit proxy by virtual work (7 rows): the joints of a limb deliver the power f . v against the load f
        they carry, and that power is capped by a fraction `duty` of the limb's rated power sum_k effort_k velocity_k
        (URDF <limit effort= velocity=>, pin.Model.effortLimit / velocityLimit).  Stance leg i carries a quarter of
        the weight: |(M g / 4) v_z(foot i)| <= P_leg (4 rows); the arm carries its distal links plus a payload:
        |m g v_z(gripper)| <= P_arm, and its wrist torque: |tau_wrist w_x|, |tau_wrist w_y| <= P_arm (3 rows).
    """
    big, LWA = 1e30, 2
    rows = []
    for foot in range(4):
        for cx, cy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            rows.append((foot, LWA, [cx, cy, -mu, 0, 0, 0], -big, 0.0))
    eff = np.asarray(table.effort, dtype=float)
    vel = np.asarray(table.velocity, dtype=float)
    total_mass = float(np.sum(table.mass))
    fz = total_mass * gravity / 4.0
    for foot in range(4):
        j = table.frame_parent[table.getFrameId(["FR_foot_fixed", "FL_foot_fixed", "RR_foot_fixed", "RL_foot_fixed"][foot],
                                                "FIXED_JOINT")]
        iv = table.idx_v[j]                                   # calf joint: the leg's columns are [iv - 2, iv]
        p_leg = duty * float(np.sum(eff[iv - 2:iv + 1] * vel[iv - 2:iv + 1]))
        rows.append((foot, LWA, [0, 0, fz, 0, 0, 0], -p_leg, p_leg))
    jw = table.getJointId("waist")
    jg = table.getJointId("gripper")
    if jw < table.njoints and jg < table.njoints:
        arm = range(table.idx_v[jw], table.idx_v[jg])         # waist .. wrist joints
        p_arm = duty * float(sum(eff[k] * vel[k] for k in arm))
        m_arm = float(sum(table.mass[j] for j in range(jw + 2, table.njoints))) + payload
        tau_w = float(eff[table.idx_v[jg] - 1])
        rows.append((4, LWA, [0, 0, m_arm * gravity, 0, 0, 0], -p_arm, p_arm))
        rows.append((4, LWA, [0, 0, 0, tau_w, 0, 0], -p_arm, p_arm))
        rows.append((4, LWA, [0, 0, 0, 0, tau_w, 0], -p_arm, p_arm))
    return rows


def rank_slice(N, rank, world):
    """Contiguous shard [lo, hi) of N states for `rank` of `world`."""
    per = (N + world - 1) // world
    lo = min(N, rank * per)
    return lo, min(N, lo + per)
