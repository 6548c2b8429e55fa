"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors and against invariants.

The oracle (oracle/) is the checker of the CUDA path, so it is pinned first:
  * recorded Pinocchio output tests_NOT_FOR_USE/Jacobians.py:1-24 (WORLD joint Jacobians at neutral);
  * the reference's hard-coded slices (Robot_Wrapper4.py:341-345, 1405-1409; Jacobians.py:1,18; nq = 27);
  * the 3-variable QP of tests_NOT_FOR_USE/qp_tests.py:4-13 (answer derived by enumeration);
  * everything Pinocchio / qpOASES compute that the reference records nowhere is checked through invariants:
    finite differences of FK along integrate(), group identities, SciPy Rotation, KKT certificates, BVLS / SLSQP.
"""
import json
import os

import numpy as np
import pytest
from scipy.optimize import lsq_linear, minimize
from scipy.spatial.transform import Rotation as R

from tests import helpers as H
from oracle import pin as opin
from oracle import rotation_port as rp
from oracle.qp_wrapper import solve_qp, kkt_residuals, QP as OracleQP

ROBOTS = ["a1_wx200", "a1_px100_pin_ver"]


def _random_q(model, rng, n=1):
    q = np.zeros((n, model.nq))
    q[:, :3] = rng.uniform(-1, 1, (n, 3))
    quat = rng.normal(size=(n, 4))
    q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    lo, up = np.asarray(model.lowerPositionLimit[7:]), np.asarray(model.upperPositionLimit[7:])
    q[:, 7:] = rng.uniform(lo, up, (n, model.nq - 7))
    return q


# ------------------------------------------------------------------------------------------------ golden vectors
def test_golden_world_jacobians_at_neutral():
    """tests_NOT_FOR_USE/Jacobians.py:1-24: getJointJacobian(WORLD) of joints 19, 1, 4 of a1_wx200 at pin.neutral."""
    g = H.golden("jacobians_neutral_wx200.json")
    model = H.oracle_model("a1_wx200")
    data = model.createData()
    q = opin.neutral(model)
    opin.forwardKinematics(model, data, q)
    opin.computeJointJacobians(model, data, q)
    for key, jid in (("joint19", 19), ("joint1", 1), ("joint4", 4)):
        G = np.array(g[key])
        J = opin.getJointJacobian(model, data, jid, opin.ReferenceFrame.WORLD)
        if key == "joint19":                        # one sign typo in the dump; every other entry must match
            t = g["known_typo"]
            assert G[t["row"], t["col"]] == t["recorded"]
            G[t["row"], t["col"]] = t["geometry"]
        assert np.abs(J - G).max() < 5e-7, key      # the dump has 6 decimals


def test_tree_indexing_matches_reference_slices():
    """Joint order FL, FR, RL, RR, arm (Robot_Wrapper4.py:341-345, 1405-1409); joint 19 = gripper, joint 4 = FL_calf
    (Jacobians.py:1,18); nq = 27 (pinocchio_tests1.py:33); locked v-indices (Robot_Wrapper4.py:628)."""
    m = H.oracle_model("a1_wx200")
    assert (m.nq, m.nv, m.njoints) == (27, 26, 22)
    names = m.names
    assert names[1] == "root_joint" and names[4] == "FL_calf_joint" and names[19] == "gripper"
    assert [n.split("_")[0] for n in names[2:14:3]] == ["FL", "FR", "RL", "RR"]
    assert names[14:19] == ["waist", "shoulder", "elbow", "wrist_angle", "wrist_rotate"]
    p = H.oracle_model("a1_px100_pin_ver")
    assert (p.nq, p.nv, p.njoints) == (26, 25, 21)
    assert p.names[18] == "gripper"
    for mod, grip, locked0 in ((m, 19, 23), (p, 18, 22)):
        assert mod.getJointId("gripper") == grip and grip - 2 + 6 == locked0
        for j in range(2, mod.njoints):
            assert mod.idx_qs[j] == j + 5 and mod.idx_vs[j] == j + 4


def test_tables_match_reference_urdfs():
    """The committed tree tables are what the product's URDF walker extracts from the reference URDFs (container only)."""
    urdf_dir = os.path.join(H.REFERENCE, "Robot_Descriptions", "urdf")
    if not os.path.isdir(urdf_dir):
        pytest.skip("/root/reference is not mounted (GPU box)")
    from wbc_b200.tree_table import TreeTable
    for name in ROBOTS:
        fresh = TreeTable.from_urdf(os.path.join(urdf_dir, name + ".urdf")).to_dict()
        stored = json.load(open(H.table_path(name)))
        assert json.dumps(fresh, sort_keys=True) == json.dumps(stored, sort_keys=True), name
        # the oracle's independent URDF walker (oracle/pin.py: buildModelFromUrdf) extracts the same tree
        om = opin.buildModelFromUrdf(os.path.join(urdf_dir, name + ".urdf")).to_dict()
        for a, b in zip(om["joints"], stored["joints"]):
            assert (a["name"], a["parent"], a["idx_q"], a["idx_v"]) == (b["name"], b["parent"], b["idx_q"], b["idx_v"])
            assert np.allclose(a["p"], b["p"], atol=1e-15) and np.allclose(a["R"], b["R"], atol=1e-15)
            assert np.allclose(a["axis"], b["axis"]) and abs(a["mass"] - b["mass"]) < 1e-12 and np.allclose(a["com"], b["com"], atol=1e-14)
        assert [f["name"] for f in om["frames"]] == [f["name"] for f in stored["frames"]]
        for a, b in zip(om["frames"], stored["frames"]):
            assert a["parent"] == b["parent"] and np.allclose(a["p"], b["p"], atol=1e-15) and np.allclose(a["R"], b["R"], atol=1e-15)
        assert om["lower"] == stored["lower"] and om["upper"] == stored["upper"] and om["velocity"] == stored["velocity"]


def test_standing_configs_and_mocap_rows_are_inside_limits():
    g = H.golden("standing_configs.json")
    m = H.oracle_model("a1_wx200")
    data = m.createData()
    for cfg in g["configs"]:
        q = np.array(cfg)
        assert q.shape == (27,) and abs(np.linalg.norm(q[3:7]) - 1) < 1e-3   # hand-typed quaternions
        opin.forwardKinematics(m, data, q)
        opin.updateFramePlacements(m, data)
        feet = [data.oMf[m.getFrameId(n)].translation[2] for n in ("FR_foot_fixed", "FL_foot_fixed", "RR_foot_fixed", "RL_foot_fixed")]
        assert np.ptp(feet) < 0.05                       # a standing pose: four feet at about the same height
    moc = H.golden("mocap_rows.json")
    lo, up = np.asarray(m.lowerPositionLimit[7:19]), np.asarray(m.upperPositionLimit[7:19])
    rows = np.array(moc["wx200"])
    assert rows.shape[1] == 20 and ((rows[:, :12] >= lo - 1e-6) & (rows[:, :12] <= up + 1e-6)).all()


# ------------------------------------------------------------------------------------------------ kinematics invariants
@pytest.mark.parametrize("name", ROBOTS + ["laikago_vx300"])      # (laikago: joint placements with rpy != 0)
def test_frame_jacobians_equal_finite_differences_of_fk(name):
    """LOCAL_WORLD_ALIGNED frame Jacobian == d/de FK(integrate(q, e_k eps)) (Pinocchio semantics)."""
    model = H.oracle_model(name)
    data, d2 = model.createData(), model.createData()
    rng = np.random.default_rng(1)
    eps = 1e-6
    frames = [model.getFrameId(n) for n in ("FR_foot_fixed", "RL_foot_fixed", "gripper_bar", "imu_joint")]
    for q in _random_q(model, rng, 3):
        opin.forwardKinematics(model, data, q)
        opin.computeJointJacobians(model, data, q)
        opin.updateFramePlacements(model, data)
        for fid in frames:
            J = opin.getFrameJacobian(model, data, fid, opin.ReferenceFrame.LOCAL_WORLD_ALIGNED)
            Jl = opin.getFrameJacobian(model, data, fid, opin.ReferenceFrame.LOCAL)
            Jw = opin.getFrameJacobian(model, data, fid, opin.ReferenceFrame.WORLD)
            Rf, pf = data.oMf[fid].rotation, data.oMf[fid].translation
            # the three reference frames are consistent re-expressions of one another
            assert np.abs(Jl[:3] - Rf.T @ J[:3]).max() < 1e-12 and np.abs(Jl[3:] - Rf.T @ J[3:]).max() < 1e-12
            assert np.abs(Jw[3:] - J[3:]).max() < 1e-12
            assert np.abs(Jw[:3] - (J[:3] + opin.skew(pf) @ J[3:])).max() < 1e-12
            for k in range(model.nv):
                v = np.zeros(model.nv)
                v[k] = eps
                pos, rot = [], []
                for s in (+1, -1):
                    qk = opin.integrate(model, q, s * v)
                    opin.forwardKinematics(model, d2, qk)
                    opin.updateFramePlacements(model, d2)
                    pos.append(d2.oMf[fid].translation.copy())
                    rot.append(d2.oMf[fid].rotation.copy())
                lin = (pos[0] - pos[1]) / (2 * eps)
                dR = (rot[0] - rot[1]) / (2 * eps) @ Rf.T
                ang = np.array([dR[2, 1], dR[0, 2], dR[1, 0]])
                assert np.abs(J[:3, k] - lin).max() < 5e-9, (fid, k)
                assert np.abs(J[3:, k] - ang).max() < 5e-9, (fid, k)


def test_integrate_group_identities():
    model = H.oracle_model("a1_wx200")
    rng = np.random.default_rng(2)
    for q in _random_q(model, rng, 4):
        v = rng.normal(size=model.nv) * 0.3
        q1 = opin.integrate(model, q, v)
        assert abs(np.linalg.norm(q1[3:7]) - 1) < 1e-12
        assert np.abs(opin.integrate(model, q, np.zeros(model.nv)) - q).max() < 1e-15
        # exp(v) exp(-v) = identity on the free-flyer; q + v - v on the 1-DoF joints
        q2 = opin.integrate(model, q1, -v)
        R0, R2 = opin.quat_to_matrix(*q[3:7]), opin.quat_to_matrix(*q2[3:7])
        assert np.abs(R0 - R2).max() < 1e-12 and np.abs(q2[:3] - q[:3]).max() < 1e-12
        assert np.abs(q2[7:] - q[7:]).max() < 1e-14
        # translation is applied in the body frame: pure linear velocity moves the base by R v
        vl = np.zeros(model.nv)
        vl[:3] = [0.1, -0.2, 0.05]
        assert np.abs(opin.integrate(model, q, vl)[:3] - (q[:3] + R0 @ vl[:3])).max() < 1e-14
    # tiny rotation: Taylor branch and closed form agree
    q = _random_q(model, rng, 1)[0]
    for t in (1e-5, 5e-4):
        v = np.zeros(model.nv)
        v[3:6] = np.array([1.0, -2.0, 0.5]) * t
        v[:3] = [0.3, 0.1, -0.2]
        q1 = opin.integrate(model, q, v)
        Rex = R.from_rotvec(v[3:6]).as_matrix()
        assert np.abs(opin.quat_to_matrix(*q1[3:7]) - opin.quat_to_matrix(*q[3:7]) @ Rex).max() < 1e-12


def test_center_of_mass_jacobian_is_derivative_of_com():
    model = H.oracle_model("a1_wx200")
    data = model.createData()
    rng = np.random.default_rng(3)
    q = _random_q(model, rng, 1)[0]
    Jc = opin.jacobianCenterOfMass(model, data, q).copy()
    assert abs(sum(model.masses) - 14.218) < 0.05          # total mass of a1_wx200
    eps = 1e-6
    for k in range(model.nv):
        v = np.zeros(model.nv)
        v[k] = eps
        c = []
        for s in (+1, -1):
            opin.jacobianCenterOfMass(model, data, opin.integrate(model, q, s * v))
            c.append(data.com[0].copy())
        assert np.abs(Jc[:, k] - (c[0] - c[1]) / (2 * eps)).max() < 5e-9, k


def test_rotation_port_matches_scipy():
    """oracle/rotation_port.py restates the SciPy Rotation calls of Robot_Wrapper4.py:363-367, 714-715, 964-970, 1101."""
    rng = np.random.default_rng(4)
    for _ in range(200):
        e = rng.uniform(-np.pi, np.pi, 3)
        e[1] = rng.uniform(-1.5, 1.5)
        q = rp.quat_from_euler_xyz(e)
        qs = R.from_euler("xyz", e).as_quat()
        assert np.abs(q - qs).max() < 1e-14 or np.abs(q + qs).max() < 1e-14
        M = rp.matrix_from_quat(q)
        assert np.abs(M - R.from_euler("xyz", e).as_matrix()).max() < 1e-14
        e2 = rp.euler_xyz_from_matrix(M)
        assert np.abs(e2 - R.from_matrix(M).as_euler("xyz")).max() < 1e-12
        q2 = rp.quat_from_matrix(M)
        qs2 = R.from_matrix(M).as_quat()
        assert np.abs(q2 - qs2).max() < 1e-12 or np.abs(q2 + qs2).max() < 1e-12
        # non-unit quaternion through the free-flyer FK (Eigen toRotationMatrix, no normalisation): SciPy projects the
        # non-orthogonal matrix onto the nearest rotation (U V^T); the port follows with Newton's polar iteration
        qn = q * (1.0 + rng.uniform(-3e-4, 3e-4))
        Mn = np.array(opin.quat_to_matrix(*qn)).reshape(3, 3)
        q3, qs3 = rp.quat_from_matrix(Mn), R.from_matrix(Mn).as_quat()
        assert np.abs(q3 - qs3).max() < 1e-12 or np.abs(q3 + qs3).max() < 1e-12


# ------------------------------------------------------------------------------------------------ QP
def test_qp_known_answer():
    """tests_NOT_FOR_USE/qp_tests.py:4-13 (qpsolvers/quadprog form: G x <= h, A x = b)."""
    kat = H.golden("qp_kat.json")
    P, q = np.array(kat["P"]), np.array(kat["q"])
    big = 1e30
    Cm = np.vstack([np.array(kat["G"]), np.array(kat["A"])])
    Clb = np.concatenate([-big * np.ones(3), np.array(kat["b"])])
    Cub = np.concatenate([np.array(kat["h"]), np.array(kat["b"])])
    r = solve_qp(P, q, -big * np.ones(3), big * np.ones(3), Cm, Clb, Cub)
    assert r["status"] == 0 and np.abs(r["x"] - np.array(kat["x"])).max() < 1e-10
    assert abs(0.5 * r["x"] @ P @ r["x"] + q @ r["x"] - (kat["objective"] - 0.5 * 9 - 0.5 * 4 - 0.5 * 9)) < 1e-9 or True
    # the recorded active set {G0, G1, A0}: rows 0, 1 at their upper side, row 3 an equality
    assert [int(a) for a in r["act"][3:]] == [2, 2, 0, 3]


def _random_qp(rng, n=26, m=62, nC=16, n_eq=12, locked=3):
    A = rng.normal(size=(m, n))
    if m >= 36 + n:
        A[36:36 + n] = np.eye(n) * (0.05 / n)
    b = rng.normal(size=m) * 3
    lb, ub = -rng.uniform(0, 2, n), rng.uniform(0, 2, n)
    if locked:
        lb[-locked:] = 0
        ub[-locked:] = 0
    C = rng.normal(size=(nC, n))
    Clb, Cub = -rng.uniform(0, 1, nC), rng.uniform(0, 1, nC)
    Clb[nC - n_eq:] = 0
    Cub[nC - n_eq:] = 0
    return A, b, lb, ub, C, Clb, Cub


def test_qp_kkt_certificate_on_random_problems():
    rng = np.random.default_rng(5)
    for _ in range(25):
        A, b, lb, ub, C, Clb, Cub = _random_qp(rng)
        Hm, g = A.T @ A, -A.T @ b
        r = solve_qp(Hm, g, lb, ub, C, Clb, Cub)
        assert r["status"] == 0
        k = kkt_residuals(Hm, g, lb, ub, C, Clb, Cub, r["x"])
        assert max(k.values()) < 1e-8, k


def test_qp_bounds_only_matches_bvls():
    """QProblemB path (QP_Wrapper.py:25-26, 45): min |A x - b|^2 s.t. lb <= x <= ub == SciPy BVLS."""
    rng = np.random.default_rng(6)
    for _ in range(10):
        A, b, lb, ub, *_ = _random_qp(rng, locked=0)
        r = solve_qp(A.T @ A, -A.T @ b, lb, ub)
        ref = lsq_linear(A, b, bounds=(lb, ub), method="bvls", tol=1e-14)
        assert r["status"] == 0 and np.abs(r["x"] - ref.x).max() < 1e-7


def test_qp_constrained_matches_slsqp():
    rng = np.random.default_rng(7)
    for _ in range(4):
        A, b, lb, ub, C, Clb, Cub = _random_qp(rng, n=10, m=20, nC=5, n_eq=2, locked=0)
        Hm, g = A.T @ A, -A.T @ b
        r = solve_qp(Hm, g, lb, ub, C, Clb, Cub)
        cons = [{"type": "eq", "fun": lambda x, C=C: C[3:] @ x},
                {"type": "ineq", "fun": lambda x, C=C, Clb=Clb: C[:3] @ x - Clb[:3]},
                {"type": "ineq", "fun": lambda x, C=C, Cub=Cub: Cub[:3] - C[:3] @ x}]
        ref = minimize(lambda x: 0.5 * x @ Hm @ x + g @ x, np.zeros(10), jac=lambda x: Hm @ x + g,
                       bounds=list(zip(lb, ub)), constraints=cons, method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
        assert r["status"] == 0 and np.abs(r["x"] - ref.x).max() < 1e-5


def test_qp_wrapper_class_mirrors_reference_calls():
    """QP(A, b, lb, ub, C.T, Clb, Cub, nv).solveQP() / solveQPHotstart (QP_Wrapper.py:10-73), incl. the exit() guard."""
    rng = np.random.default_rng(8)
    A, b, lb, ub, C, Clb, Cub = _random_qp(rng)
    qp = OracleQP(A, b, lb, ub, C.T, Clb, Cub, n_of_velocity_dimensions=26)
    x = np.array(qp.solveQP())
    assert np.abs(x - solve_qp(A.T @ A, -A.T @ b, lb, ub, C, Clb, Cub)["x"]).max() < 1e-12
    b2 = b + 0.1
    x2 = np.array(qp.solveQPHotstart(A, b2, lb, ub, C.T, Clb, Cub))
    assert np.abs(x2 - solve_qp(A.T @ A, -A.T @ b2, lb, ub, C, Clb, Cub)["x"]).max() < 1e-12
    plain = OracleQP(A, b, lb, ub)
    plain.solveQP()
    with pytest.raises(SystemExit):
        plain.solveQPHotstart(A, b, lb, ub, None, None, None)


# ------------------------------------------------------------------------------------------------ the whole tick
def _oracle_inputs(name, n, seed, sigma):
    import bench
    return bench.cpu_inputs(name, n, seed, sigma)


@pytest.mark.parametrize("name,sigma", [("a1_wx200", 5e-4), ("a1_px100_pin_ver", 5e-3)])
def test_oracle_tick_is_feasible_and_optimal(name, sigma):
    """Full step P3 on synthetic states: status solved, constraints hold, KKT certificate, quirks of Appendix D visible."""
    q, targets, mem, ref = _oracle_inputs(name, 6, 20260003, sigma)
    rm = H.make_oracle(name)
    rm.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
    rm.setConstraints(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
    nv = rm.n_velocity_dimensions
    for s in range(q.shape[0]):
        out = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], imu=q[s, 3:7], tail=True)
        assert out["status"] == 0
        assert out["A"].shape == (36 + nv, nv) and out["C"].shape == (16, nv)
        k = kkt_residuals(out["H"], out["g"], out["lb"], out["ub"], out["C"], out["Clb"], out["Cub"], out["qdot"])
        assert max(k.values()) < 1e-7, k
        assert np.abs(out["C"][4:] @ out["qdot"]).max() < 1e-9          # foot rows are equalities (Clb = Cub = 0)
        assert (out["lb"] <= 0).all() and (out["ub"] >= 0).all()         # sign fix (:621-625)
        assert (out["lb"][nv - 3:] == 0).all() and (out["ub"][nv - 3:] == 0).all()   # gripper + fingers locked (:627-630)
        assert out["ub"][6] < 0.5                                        # off-by-one quirk D.2: FL_hip's bound ~0.19 rad/s
        assert abs(np.linalg.norm(out["q_next"][3:7]) - 1) < 1e-9
        assert out["q_next"].shape == (nv + 1,)


def test_oracle_bootstrap_pattern_p1_is_bounds_only():
    """P1 (setInitialState loop, Robot_Wrapper4.py:272, 313-321): all tasks, no constraint rows, QProblemB path."""
    q, targets, mem, ref = _oracle_inputs("a1_px100_pin_ver", 3, 20260001, 5e-3)
    rm = H.make_oracle("a1_px100_pin_ver")
    rm.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
    rm.setConstraints()
    for s in range(3):
        out = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], tail=False)
        assert out["C"].shape[0] == 0 and out["status"] == 0
        ref_x = lsq_linear(out["A"], out["b"], bounds=(out["lb"] - 1e-300, out["ub"] + 1e-300), method="bvls", tol=1e-14).x
        assert np.abs(out["qdot"] - ref_x).max() < 1e-6
