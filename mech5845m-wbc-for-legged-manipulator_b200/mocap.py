"""Joint-angle playback files of the reference (``tests_NOT_FOR_USE/mocap*.txt``, read by the drivers, ``sim3.py:109,321``).

One frame per line: ``index, timestamp, 12 leg joints, arm joints...``, comma separated.  The leg columns are in the
PyBullet / URDF order FR, FL, RR, RL; Pinocchio orders the tree FL, FR, RL, RR (children of ``trunk`` sort by joint name),
which is also the order of ``q[7:]`` and of the ``FL_leg, FR_leg, RL_leg, RR_leg, grip`` slices ``runWBC`` returns
(Robot_Wrapper4.py:1405-1409).  These helpers convert between the two.
"""
from __future__ import annotations

import numpy as np

_LEG_PERM = [3, 4, 5, 0, 1, 2, 9, 10, 11, 6, 7, 8]          # FR FL RR RL <-> FL FR RL RR (its own inverse)


def bullet_to_pinocchio(joints):
    """[..., 12 + n_arm] joint angles in file order (FR, FL, RR, RL, arm) -> Pinocchio order (FL, FR, RL, RR, arm)."""
    joints = np.asarray(joints, dtype=np.float64)
    out = joints.copy()
    out[..., :12] = joints[..., _LEG_PERM]
    return out


def pinocchio_to_bullet(joints):
    """Inverse of :func:`bullet_to_pinocchio` (the permutation is an involution); what ``sim3.py:109`` builds with hstack."""
    return bullet_to_pinocchio(joints)


def load_mocap(path, n_joints=None):
    """Read a playback file -> [frames, 12 + n_arm] float64 in Pinocchio order.  ``n_joints`` checks the width
    (``nq - 7`` of the model the frames are meant for)."""
    rows = []
    with open(path) as fh:
        for line in fh:
            parts = [p for p in line.replace(",", " ").split() if p]
            if len(parts) < 14:
                continue
            rows.append([float(p) for p in parts[2:]])
    frames = np.asarray(rows, dtype=np.float64)
    if n_joints is not None and frames.shape[1] != n_joints:
        raise ValueError(f"{path}: {frames.shape[1]} joint columns, expected {n_joints}")
    return bullet_to_pinocchio(frames)


def configurations(frames, base_xyz=(0.0, 0.0, 0.3), base_quat=(0.0, 0.0, 0.0, 1.0)):
    """[frames, nq]: the playback joints under a fixed free-flyer pose (x y z qx qy qz qw | joints)."""
    frames = np.asarray(frames, dtype=np.float64)
    base = np.tile(np.asarray(list(base_xyz) + list(base_quat), dtype=np.float64), (frames.shape[0], 1))
    return np.concatenate([base, frames], axis=1)
