"""oracle.rotation_port -- scalar restatement of the scipy.spatial.transform.Rotation calls on the path.

TEST INFRASTRUCTURE ONLY.  The reference calls SciPy's ``Rotation`` (version unpinned; 1.18.1 is
what this image has) at ``wrappers/Robot_Wrapper4.py:222-226, 363-367, 382-383, 714-715, 964-970,
1091-1106``.  ``oracle/robot_wrapper4.py`` calls SciPy itself, verbatim.  This file restates the
same algorithms (scipy/spatial/transform/_rotation_xp.py: ``_from_matrix_orthogonal`` :90-157,
``from_euler`` :192-224 + ``_elementary_quat_compose`` :1035-1049, ``as_matrix`` :302-333,
``as_euler`` :365-403 + ``_get_angles`` :1052-1111) as straight-line scalar code -- the exact
sequence of operations the CUDA device functions in ``csrc/wbc_device.cuh`` follow -- so the
tests can pin that sequence against SciPy on the CPU before it is trusted on the GPU.
Quaternions are (x, y, z, w), as in SciPy.
"""
import math

import numpy as np


def orthogonalize(R):
    """from_matrix's pre-step: nearest orthogonal matrix U V^T when R R^T is not close to I (atol 1e-12 off the diagonal,
    1e-12 + 1e-5 on it); Newton's polar iteration X <- (X + X^-T) / 2, the sequence the CUDA device function follows."""
    R = np.array(R, dtype=float)
    G = R @ R.T
    ok = all((abs(G[i, j] - 1.0) <= 1e-12 + 1e-5) if i == j else (abs(G[i, j]) <= 1e-12) for i in range(3) for j in range(i, 3))
    if ok:
        return R
    for _ in range(30):
        X = 0.5 * (R + np.linalg.inv(R).T)
        change = np.abs(X - R).max()
        R = X
        if change <= 1e-16:
            break
    return R


def quat_from_matrix(R):
    """Rotation.from_matrix(R).as_quat() (no sign canonicalisation); R is orthogonalised first when SciPy would."""
    R = orthogonalize(R)
    tr = R[0][0] + R[1][1] + R[2][2]
    dec = [R[0][0], R[1][1], R[2][2], tr]
    choice = 0
    for i in range(1, 4):          # argmax: first maximum wins
        if dec[i] > dec[choice]:
            choice = i
    if choice == 0:
        q = [1 - tr + 2 * R[0][0], R[1][0] + R[0][1], R[2][0] + R[0][2], R[2][1] - R[1][2]]
    elif choice == 1:
        q = [R[1][0] + R[0][1], 1 - tr + 2 * R[1][1], R[2][1] + R[1][2], R[0][2] - R[2][0]]
    elif choice == 2:
        q = [R[2][0] + R[0][2], R[2][1] + R[1][2], 1 - tr + 2 * R[2][2], R[1][0] - R[0][1]]
    else:
        q = [R[2][1] - R[1][2], R[0][2] - R[2][0], R[1][0] - R[0][1], 1 + tr]
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return np.array([q[0] / n, q[1] / n, q[2] / n, q[3] / n])


def _compose(p, q):
    """compose_quat(p, q) = p * q."""
    cx = p[1] * q[2] - p[2] * q[1]
    cy = p[2] * q[0] - p[0] * q[2]
    cz = p[0] * q[1] - p[1] * q[0]
    return np.array([
        p[3] * q[0] + q[3] * p[0] + cx,
        p[3] * q[1] + q[3] * p[1] + cy,
        p[3] * q[2] + q[3] * p[2] + cz,
        p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2],
    ])


def quat_from_euler_xyz(e):
    """Rotation.from_euler('xyz', e).as_quat(): extrinsic => q = qz * (qy * qx)."""
    qx = np.array([math.sin(e[0] / 2.0), 0.0, 0.0, math.cos(e[0] / 2.0)])
    qy = np.array([0.0, math.sin(e[1] / 2.0), 0.0, math.cos(e[1] / 2.0)])
    qz = np.array([0.0, 0.0, math.sin(e[2] / 2.0), math.cos(e[2] / 2.0)])
    return _compose(qz, _compose(qy, qx))


def matrix_from_quat(q):
    """Rotation.as_matrix()."""
    x, y, z, w = q
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return np.array([
        [x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)],
        [2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)],
        [2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2],
    ])


def _wrap(a):
    # (a + pi) % (2 pi) - pi with Python/NumPy floor-mod semantics
    two_pi = 2 * math.pi
    r = math.fmod(a + math.pi, two_pi)
    if r < 0:
        r += two_pi
    return r - math.pi


def euler_xyz_from_quat(q):
    """Rotation.as_euler('xyz'): extrinsic, asymmetric, i, j, k = 0, 1, 2, sign = +1."""
    x, y, z, w = q
    a = w - y
    b = x + z
    c = y + w
    d = z - x
    eps = 1e-7
    half_sum = math.atan2(b, a)
    half_diff = math.atan2(d, c)
    a1 = 2 * math.atan2(math.hypot(c, d), math.hypot(a, b))
    case1 = abs(a1) <= eps
    case2 = abs(a1 - math.pi) <= eps
    if not (case1 or case2):
        a0 = half_sum - half_diff
        a2 = half_sum + half_diff
    else:
        a2 = 0.0
        a0 = 2 * half_sum if case1 else -2 * half_diff
    a1 -= math.pi / 2
    return np.array([_wrap(a0), _wrap(a1), _wrap(a2)])


def euler_xyz_from_matrix(R):
    return euler_xyz_from_quat(quat_from_matrix(R))


def matrix_from_euler_xyz(e):
    return matrix_from_quat(quat_from_euler_xyz(e))
