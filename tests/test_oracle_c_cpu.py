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
