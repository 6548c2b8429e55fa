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


def rank_slice(N, rank, world):
    """Contiguous shard [lo, hi) of N states for `rank` of `world`."""
    per = (N + world - 1) // world
    lo = min(N, rank * per)
    return lo, min(N, lo + per)
