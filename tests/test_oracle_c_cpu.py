"""The plain-C restatement (oracle/wbc_oracle.c) pinned against the NumPy/SciPy oracle: same inputs, same A / b,
same solutions, same pivoting path.  CPU only."""
import numpy as np
import pytest

from oracle import c_port
from tests import helpers as H
import bench


def _oracle(name, tasks, cons, joint):
    rm = H.make_oracle(name)
    rm.setTasks(Joint=joint, **tasks)
    rm.setConstraints(**cons)
    return rm


ALL = dict(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True)
GRIP = dict(Trunk=False, FR=False, FL=False, RR=False, RL=False, Grip=True)
NOC = dict(CoM=False, Trunk=False, FR=False, FL=False, RR=False, RL=False, Grip=False)
P2C = dict(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
COMC = dict(CoM=True, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)


@pytest.mark.parametrize("name,tasks,cons,joint,sigma", [
    ("a1_px100_pin_ver", ALL, NOC, True, 5e-3),       # P1 bootstrap pattern: bounds only
    ("a1_px100_pin_ver", GRIP, P2C, "PREV", 5e-4),    # P2 sim3 tick pattern
    ("a1_wx200", ALL, P2C, True, 5e-3),               # P3 full stack, stress sigma (inequalities enter and leave)
    ("a1_wx200", ALL, COMC, True, 5e-4),              # CoM rows on top of the P3 constraint set
])
def test_c_port_equals_numpy_oracle(name, tasks, cons, joint, sigma):
    N = 24
    q, targets, mem, ref = bench.cpu_inputs(name, N, 20260002, sigma)
    rm = _oracle(name, tasks, cons, joint)
    ts, table = c_port.table_struct(name)
    cfg = c_port.config_struct(rm, table)
    out = c_port.step(ts, cfg, q, targets, mem, ref, rm.dt, nthreads=1, want_Ab=True)
    nv = table.nv
    for s in range(N):
        r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], solve=True, tail=False)
        assert np.abs(out["A"][s] - r["A"]).max() < 1e-12
        assert np.abs(out["b"][s] - r["b"]).max() < 1e-9 * max(1.0, np.abs(r["b"]).max())
        assert np.abs(out["mem_out"][s] - r["mem_out"]).max() < 1e-13
        assert int(out["status"][s]) == int(r["status"])
        if r["status"] != 0:                # infeasible CoM box for a tilted base: both must say so, x is undefined
            continue
        assert np.abs(out["qdot"][s] - r["qdot"]).max() < 1e-8
        assert int(out["iters"][s]) == int(r["iters"])
        wb, wr = H.act_to_bits(r["act"], nv)
        assert int(out["active_set"][s, 0]) == wb and int(out["active_set"][s, 1]) == wr


def test_c_port_solutions_carry_a_kkt_certificate():
    name, N = "a1_wx200", 64
    q, targets, mem, ref = bench.cpu_inputs(name, N, 7, 5e-3)
    rm = _oracle(name, ALL, P2C, True)
    ts, table = c_port.table_struct(name)
    out = c_port.step(ts, c_port.config_struct(rm, table), q, targets, mem, ref, rm.dt, nthreads=1, want_Ab=True)
    for s in range(0, N, 4):
        r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], solve=False, tail=False)
        Hm, g = out["A"][s].T @ out["A"][s], -out["A"][s].T @ out["b"][s]
        k = H.kkt_residuals(Hm, g, r["lb"], r["ub"], r["C"], r["Clb"], r["Cub"], out["qdot"][s])
        assert max(k.values()) < 1e-6 * max(1.0, np.abs(g).max()), (s, k)


def test_config3_rows_both_oracles_take_the_same_path():
    """BASELINE config 3 (friction pyramid + torque-limit proxy rows through the extension-row channel, 27 rows): the two
    oracles -- dense re-solves (NumPy) and factor updating (C) -- agree on minimiser, pivoting path and active set on
    every state.  Mirrored pyramid faces tie exactly once their partners are active; the tie window of the entering rule
    (oracle/qp_wrapper.py: TIE_REL / TIE_ABS, the same constants as csrc/wbc_qp.cuh) makes the choice independent of
    rounding noise.  Without the window this test finds different active sets on ~5 % of the states."""
    from wbc_b200 import synthetic, TreeTable
    name, N = "a1_wx200", 96
    q, targets, mem, ref = bench.cpu_inputs(name, N, 20260003, 5e-3)
    rm = _oracle(name, ALL, dict(CoM=False, Trunk=True, FR=False, FL=False, RR=False, RL=False, Grip=False), True)
    rm.extra_rows = synthetic.config3_rows(TreeTable.load(name))
    assert len(rm.extra_rows) == 23
    ts, table = c_port.table_struct(name)
    out = c_port.step(ts, c_port.config_struct(rm, table), q, targets, mem, ref, rm.dt, nthreads=1)
    bound = 0
    for s in range(N):
        r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], solve=True, tail=False)
        assert r["C"].shape == (27, 26)
        assert int(out["status"][s]) == int(r["status"]) == 0
        assert np.abs(out["qdot"][s] - r["qdot"]).max() < 1e-8
        assert int(out["iters"][s]) == int(r["iters"])
        wb, wr = H.act_to_bits(r["act"], 26)
        assert int(out["active_set"][s, 0]) == wb and int(out["active_set"][s, 1]) == wr, s
        bound += bool(wr >> 40)
        res = _kkt_nnls(r["H"], r["g"], r["lb"], r["ub"], r["C"], r["Clb"], r["Cub"], out["qdot"][s])
        assert res < 1e-7 * max(1.0, np.abs(r["g"]).max()), (s, res)
    assert bound > N // 2                                   # the torque-limit proxy rows do bind


def _kkt_nnls(Hm, g, lb, ub, C, Clb, Cub, x, tol=1e-7):
    """KKT certificate that also holds at DEGENERATE vertices (four pyramid faces tight in a three-dimensional velocity
    space: the multipliers are not unique, so a least-squares recovery may come out with the wrong signs although valid
    ones exist): non-negative least squares over the inward normals of every tight side.  Returns the stationarity
    residual |H x + g - sum_k lam_k n_k|_inf with lam >= 0; primal feasibility is asserted."""
    from scipy.optimize import nnls
    n = len(x)
    Aall = np.vstack([np.eye(n), C])
    lo, up = np.concatenate([lb, Clb]), np.concatenate([ub, Cub])
    ax = Aall @ x
    assert (ax >= lo - tol).all() and (ax <= up + tol).all()
    cols = [Aall[c] for c in range(len(ax)) if abs(ax[c] - lo[c]) <= tol] + [-Aall[c] for c in range(len(ax)) if abs(ax[c] - up[c]) <= tol]
    grad = Hm @ x + g
    if not cols:
        return float(np.abs(grad).max())
    Nm = np.array(cols).T
    lam, _ = nnls(Nm, grad, maxiter=50 * Nm.shape[1])
    return float(np.abs(Nm @ lam - grad).max())


def test_entering_rule_tie_window():
    """Two mirrored rows violated by exactly the same amount up to rounding: the lower index must win on both oracles
    whichever of the two is `more violated' in the last bits."""
    from oracle.qp_wrapper import solve_qp
    Hm = np.eye(3)
    for eps in (0.0, 3e-16, -3e-16):
        g = -np.array([0.0, 0.0, -1.0])                     # unconstrained minimiser (0, 0, -1)
        C = np.array([[1.0, 0.0, 0.6 * (1 + eps)], [-1.0, 0.0, 0.6]])     # x + 0.6 z >= 0, -x + 0.6 z >= 0: both violated by 0.6
        r = solve_qp(Hm, g, -np.ones(3) * 10, np.ones(3) * 10, C, np.zeros(2), np.ones(2) * 1e30)
        assert r["working_set"][0][0] == 3, (eps, r["working_set"])       # row 0 (constraint id nv + 0) entered first
        assert r["status"] == 0 and np.abs(C @ r["x"]).max() < 1e-12 + 0 * eps or (C @ r["x"] >= -1e-12).all()


def test_damper_compat_flag_off_both_oracles():
    """compat_damper_off_by_one = False (the velocity damper with each joint's own coordinate): C port == NumPy oracle."""
    name, N = "a1_wx200", 16
    q, targets, mem, ref = bench.cpu_inputs(name, N, 11, 5e-4)
    rm = _oracle(name, ALL, P2C, True)
    rm.compat_damper_off_by_one = False
    lo = np.asarray(rm.robot_model.lowerPositionLimit[7:])
    q[:, 7:19] = np.where(np.arange(12) % 2 == 0, lo[:12] + 0.01, q[:, 7:19])     # every other leg joint inside the damper zone
    ts, table = c_port.table_struct(name)
    out = c_port.step(ts, c_port.config_struct(rm, table), q, targets, mem, ref, rm.dt, nthreads=1)
    rm_on = _oracle(name, ALL, P2C, True)
    differs = 0
    for s in range(N):
        r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], solve=True, tail=False)
        r_on = H.oracle_step_one(rm_on, q[s], targets[s], mem[s], ref[s], solve=False, tail=False)
        differs += np.abs(r["lb"] - r_on["lb"]).max() > 1e-6
        assert int(out["status"][s]) == int(r["status"])
        if r["status"] == 0:
            assert np.abs(out["qdot"][s] - r["qdot"]).max() < 1e-8 and int(out["iters"][s]) == int(r["iters"])
    assert differs > 0
