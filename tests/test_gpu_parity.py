"""-m gpu: parity of the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): FK poses / frame Jacobians 1e-10 abs, QP solutions 1e-6 with
an identical active set.  All arithmetic is float64.
"""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

FK_TOL = 1e-10
QP_TOL = 1e-6


def _maxabs(a):
    a = np.asarray(a)
    return float(np.abs(a).max()) if a.size else 0.0


def _robot(name, N, tasks, cons, joint=True):
    import wbc_b200
    r = wbc_b200.RobotModel(name, batch=N, device="cuda:0")
    r.setTasks(Trunk=tasks["Trunk"], FR=tasks["FR"], FL=tasks["FL"], RR=tasks["RR"], RL=tasks["RL"], Grip=tasks["Grip"],
               Joint=joint)
    r.setConstraints(**cons)
    return r


P1_TASKS = dict(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True)
P2_TASKS = dict(Trunk=False, FR=False, FL=False, RR=False, RL=False, Grip=True)
NO_CONS = dict(CoM=False, Trunk=False, FR=False, FL=False, RR=False, RL=False, Grip=False)
P2_CONS = dict(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)


def _load(robot, N, seed, sigma):
    from wbc_b200 import synthetic
    q = synthetic.sample_configurations(robot.robot_model, N, seed)
    noise = synthetic.sample_noise(N, seed, sigma)
    targets = synthetic.load_batch(robot, q, noise)
    return q, targets


@pytest.mark.parametrize("name", ["a1_wx200", "a1_px100_pin_ver", "laikago_vx300"])
def test_fk_and_frame_jacobians_match_oracle(name):
    import wbc_b200
    from wbc_b200 import synthetic
    N = 256
    robot = wbc_b200.RobotModel(name, batch=N, device="cuda:0")
    q = synthetic.sample_configurations(robot.robot_model, N, 11)
    q[N // 2:, 3:7] = np.random.default_rng(5).normal(size=(N - N // 2, 4))     # uniform SO(3) half
    q[N // 2:, 3:7] /= np.linalg.norm(q[N // 2:, 3:7], axis=1, keepdims=True)
    robot.updateState(torch.as_tensor(q, device="cuda:0"), feedback=False)
    model = H.oracle_model(name)
    data = model.createData()
    got = {rf: robot.frameJacobians(rf) for rf in (0, 1, 2)}
    oMi = robot._oMi.cpu().numpy()
    Jw = robot.J.cpu().numpy()
    frames = robot.end_effector_index_list_frame + [robot.trunk_frame_index]
    worst = 0.0
    for s in range(0, N, 3):
        H.opin.forwardKinematics(model, data, q[s])
        H.opin.computeJointJacobians(model, data, q[s])
        H.opin.updateFramePlacements(model, data)
        worst = max(worst, np.abs(Jw[s] - data.J).max())
        for j in range(1, model.njoints):
            ref = np.concatenate([data.oMi[j].rotation.reshape(-1), data.oMi[j].translation])
            worst = max(worst, np.abs(oMi[s, j] - ref).max())
        for rf in (0, 1, 2):
            oMf, J = got[rf]
            for k, fid in enumerate(frames):
                Jr = H.opin.getFrameJacobian(model, data, fid, rf)
                worst = max(worst, np.abs(J[s, k].cpu().numpy() - Jr).max())
                ref = np.concatenate([data.oMf[fid].rotation.reshape(-1), data.oMf[fid].translation])
                worst = max(worst, np.abs(oMf[s, k].cpu().numpy() - ref).max())
    assert worst < FK_TOL, worst


def test_golden_jacobians_neutral_wx200():
    """Reference's recorded Pinocchio output tests_NOT_FOR_USE/Jacobians.py:1-24 through the CUDA path."""
    import wbc_b200
    g = H.golden("jacobians_neutral_wx200.json")
    robot = wbc_b200.RobotModel("a1_wx200", batch=1, device="cuda:0")
    Jw = robot.J[0].cpu().numpy()                       # constructor leaves the model at pin.neutral
    t = robot.robot_model
    for key, jid in (("joint19", 19), ("joint1", 1), ("joint4", 4)):
        G = np.array(g[key])
        mask = np.array([(t.support_mask(jid) >> k) & 1 for k in range(t.nv)], dtype=float)
        Jj = Jw * mask
        if key == "joint19":
            typo = g["known_typo"]
            assert G[typo["row"], typo["col"]] == typo["recorded"]
            G[typo["row"], typo["col"]] = typo["geometry"]
        assert np.abs(Jj - G).max() < 5e-7, key        # the dump is printed with 6 decimals


@pytest.mark.parametrize("name,tasks,cons,joint,sigma", [
    ("a1_px100_pin_ver", P1_TASKS, NO_CONS, True, 5e-3),      # P1 bootstrap pattern: bounds only
    ("a1_px100_pin_ver", P2_TASKS, P2_CONS, "PREV", 5e-4),    # P2 sim3 tick pattern
    ("a1_wx200", P1_TASKS, P2_CONS, True, 5e-3),              # P3 full stack + constraints (stress sigma)
    ("a1_wx200", P1_TASKS, P2_CONS, True, 5e-4),              # P3 nominal sigma
    ("a1_px100_pin_ver", P2_TASKS, P2_CONS, "HYBRID", 5e-4),  # what sim3.py:145 actually runs (f2): FD manipulability gradient
    ("a1_wx200", P1_TASKS, P2_CONS, "HYBRID", 5e-4),
    ("a1_wx200", P2_TASKS, P2_CONS, "MANI", 5e-4),
    ("laikago_vx300", P1_TASKS, P2_CONS, True, 5e-4),         # third robot of sim3.py:13 (joint rpy != 0, no `gripper` joint:
                                                               #  the reference's look-up returns njoints, nothing gets locked)
])
def test_assembly_and_qp_match_oracle(name, tasks, cons, joint, sigma):
    N = 96 if joint in (True, "PREV") else (32 if joint == "HYBRID" else 12)    # the FD modes cost 12 / 52 oracle FK passes per state
    robot = _robot(name, N, tasks, cons, joint)
    q, targets = _load(robot, N, 20260002, sigma)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    asm = robot.assemble(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18])
    imu = torch.as_tensor(q[:, 3:7], device="cuda:0")
    qdot = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], imu_quat=imu, advance=True).cpu().numpy()
    ref = H.oracle_step_batch(name, robot, q, targets.cpu().numpy(), mem0.cpu().numpy(), ref0.cpu().numpy(),
                              imu=q[:, 3:7], tail=True)
    for k in ("A", "lb", "ub", "C"):
        assert _maxabs(asm[k].cpu().numpy() - ref[k]) < FK_TOL, k
    for k in ("b", "Clb", "Cub", "g"):          # these carry 1/dt = 500: scale the tolerance with the magnitude
        a, r = asm[k].cpu().numpy(), ref[k]
        assert _maxabs(a - r) < 1e-9 * max(1.0, _maxabs(r)), k
    assert np.abs(asm["H"].cpu().numpy() - ref["H"]).max() < 1e-10
    status = robot.last_status.cpu().numpy()
    assert (status == 0).all() and (ref["status"] == 0).all()
    assert np.abs(qdot - ref["qdot"]).max() < QP_TOL
    # identical active set and identical pivoting path
    act = robot.last_active_set.cpu().numpy().astype(np.uint64)
    nv = robot.n_velocity_dimensions
    for s in range(N):
        wb, wr = H.act_to_bits(ref["act"][s], nv)
        assert int(act[s, 0]) == wb and int(act[s, 1]) == wr, s
    assert (robot.last_iters.cpu().numpy() == ref["iters"]).all()
    # task memory and next configuration (integrate + base estimate)
    assert np.abs(robot._mem.cpu().numpy() - ref["mem_out"]).max() < 1e-12
    assert np.abs(robot.current_joint_config.cpu().numpy() - ref["q_next"]).max() < 1e-8


def test_qp_kat_and_random_problems():
    """QP drop-in: the reference's only QP example (tests_NOT_FOR_USE/qp_tests.py) + random dense problems."""
    import wbc_b200
    from oracle.qp_wrapper import solve_qp
    kat = H.golden("qp_kat.json")
    P, qv = np.array(kat["P"]), np.array(kat["q"])
    L = np.linalg.cholesky(P)
    A = L.T                                           # A^T A = P
    b = -np.linalg.solve(L, qv)                        # -A^T b = q
    big = 1e30
    Cm = np.vstack([np.array(kat["G"]), np.array(kat["A"])])
    Clb = np.concatenate([-big * np.ones(3), np.array(kat["b"])])
    Cub = np.concatenate([np.array(kat["h"]), np.array(kat["b"])])
    qp = wbc_b200.QP(A, b, -big * np.ones(3), big * np.ones(3), Cm.T, Clb, Cub, n_of_velocity_dimensions=3)
    x = qp.solveQP().cpu().numpy()
    assert np.abs(x - np.array(kat["x"])).max() < 1e-9
    assert int(qp.status[0]) == 0

    rng = np.random.default_rng(3)
    N, n, m, nC = 128, 26, 62, 16
    A = rng.normal(size=(N, m, n)); A[:, 36:, :] = 0
    A[:, 36:, :] = np.eye(n)[None] * (0.05 / 26)
    b = rng.normal(size=(N, m)) * 3
    lb = -rng.uniform(0, 2, (N, n)); ub = rng.uniform(0, 2, (N, n)); lb[:, 23:] = 0; ub[:, 23:] = 0
    lb[::2, 5] = ub[::2, 5] = 0.3; lb[::3, 17] = ub[::3, 17] = -0.2     # variables fixed at non-zero values (eliminated up front)
    Cm = rng.normal(size=(N, nC, n)); Clb = -rng.uniform(0, 1, (N, nC)); Cub = rng.uniform(0, 1, (N, nC))
    Clb[:, 4:] = 0; Cub[:, 4:] = 0
    qp = wbc_b200.QP(A, b, lb, ub, np.transpose(Cm, (0, 2, 1)), Clb, Cub, n_of_velocity_dimensions=n)
    x = qp.solveQP().cpu().numpy()
    assert (qp.status.cpu().numpy() == 0).all()
    for s in range(N):
        Hs, gs = A[s].T @ A[s], -A[s].T @ b[s]
        r = solve_qp(Hs, gs, lb[s], ub[s], Cm[s], Clb[s], Cub[s])
        assert np.abs(x[s] - r["x"]).max() < QP_TOL
        k = H.kkt_residuals(Hs, gs, lb[s], ub[s], Cm[s], Clb[s], Cub[s], x[s])
        assert max(k.values()) < 1e-7, (s, k)
        assert int(qp.iters[s]) == r["iters"]
    # bounds-only path (QProblemB)
    qp2 = wbc_b200.QP(A, b, lb, ub, n_of_velocity_dimensions=n)
    x2 = qp2.solveQP().cpu().numpy()
    for s in range(0, N, 8):
        r = solve_qp(A[s].T @ A[s], -A[s].T @ b[s], lb[s], ub[s])
        assert np.abs(x2[s] - r["x"]).max() < QP_TOL


def test_full_size_properties_config2():
    """BASELINE config 2 size (PX100, 4096 states): size-independent properties of every solution."""
    N = 4096
    robot = _robot("a1_px100_pin_ver", N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260001, 5e-3)
    asm = robot.assemble(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], want=("C", "Clb", "Cub", "lb", "ub", "H", "g"))
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False)
    assert (robot.last_status == 0).all()
    Cx = torch.einsum("nrk,nk->nr", asm["C"], x)
    assert (Cx >= asm["Clb"] - 1e-8).all() and (Cx <= asm["Cub"] + 1e-8).all()
    assert (x >= asm["lb"] - 1e-9).all() and (x <= asm["ub"] + 1e-9).all()
    # foot rows are equalities: J_foot qdot = 0
    assert Cx[:, 4:].abs().max() < 1e-8
    # stationarity on the free subspace: the gradient lies in the span of the active normals
    grad = torch.einsum("nij,nj->ni", asm["H"], x) + asm["g"]
    act = robot.last_active_set.cpu().numpy().astype(np.uint64)
    Cn, gn, xn = asm["C"].cpu().numpy(), grad.cpu().numpy(), x.cpu().numpy()
    nv = robot.n_velocity_dimensions
    for s in range(0, N, 64):
        normals = [np.eye(nv)[k] for k in range(nv) if (int(act[s, 0]) >> (2 * k)) & 3]
        normals += [Cn[s, r] for r in range(Cn.shape[1]) if (int(act[s, 1]) >> (2 * r)) & 3]
        Nm = np.array(normals).T
        lam, *_ = np.linalg.lstsq(Nm, gn[s], rcond=None)
        assert np.abs(Nm @ lam - gn[s]).max() < 1e-7, s


def test_full_size_sweep_config4_one_million_states():
    """BASELINE configs[3]: the A1 + WX200 sweep over 2^20 states (the bench shards it over eight GPUs; here all of it on one).
    Size-independent properties: every QP solved; the answer of a state does not depend on the launch it rode in (slices
    relaunched as batches of their own -- other grid, other rounds, other barrier groups -- give the same bits); a stride
    sample satisfies its bounds and rows and has J_foot qdot = 0; a stride sample equals the C oracle with the same pivoting
    path and active set; a checksum of checksums over eight shards equals the checksum of the whole."""
    from oracle import c_port
    N = 1 << 20
    robot = _robot("a1_wx200", N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260003, 5e-4)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False).clone()
    st, it, act = robot.last_status.clone(), robot.last_iters.clone(), robot.last_active_set.clone()
    assert (st == 0).all() and torch.isfinite(x).all()
    assert abs(float(sum(x[k * (N // 8):(k + 1) * (N // 8)].abs().sum() for k in range(8))) - float(x.abs().sum())) < 1e-6
    for lo, n in ((0, 4096), (123457, 2369), (N - 777, 777)):
        sub = _robot("a1_wx200", n, P1_TASKS, P2_CONS, True)
        sub.current_joint_config = robot.current_joint_config[lo:lo + n].clone()
        sub._mem.copy_(mem0[lo:lo + n]); sub._ref.copy_(ref0[lo:lo + n])
        t = targets[lo:lo + n]
        xs = sub.step(t[:, :15].reshape(n, 5, 3), t[:, 15:18], advance=False)
        assert torch.equal(xs, x[lo:lo + n]) and torch.equal(sub.last_iters, it[lo:lo + n])
        assert torch.equal(sub.last_active_set, act[lo:lo + n])
        if lo == 0:
            asm = sub.assemble(t[:, :15].reshape(n, 5, 3), t[:, 15:18], want=("C", "Clb", "Cub", "lb", "ub"))
            Cx = torch.einsum("nrk,nk->nr", asm["C"], xs)
            assert (Cx >= asm["Clb"] - 1e-8).all() and (Cx <= asm["Cub"] + 1e-8).all()
            assert (xs >= asm["lb"] - 1e-9).all() and (xs <= asm["ub"] + 1e-9).all()
            assert Cx[:, 4:].abs().max() < 1e-8
    idx = torch.arange(0, N, 257, device="cuda:0")
    ts, table = c_port.table_struct("a1_wx200")
    ref = c_port.step(ts, c_port.config_struct(robot, table), q[idx.cpu().numpy()], targets[idx].cpu().numpy(),
                      mem0[idx].cpu().numpy(), ref0[idx].cpu().numpy(), robot.dt)
    assert (ref["status"] == 0).all()
    assert np.abs(x[idx].cpu().numpy() - ref["qdot"]).max() < QP_TOL
    assert (it[idx].cpu().numpy() == ref["iters"]).all()
    assert (act[idx].cpu().numpy().astype(np.uint64) == ref["active_set"]).all()


def test_config2_every_state_against_the_c_oracle():
    """BASELINE config 2: A1 + PX100, 4096 random states, every solution compared with the CPU loop
    (oracle/wbc_oracle.c, itself pinned against the NumPy/SciPy oracle on the CPU)."""
    from oracle import c_port
    N = 4096
    for sigma in (5e-4, 5e-3):
        robot = _robot("a1_px100_pin_ver", N, P1_TASKS, P2_CONS, True)
        q, targets = _load(robot, N, 20260001, sigma)
        mem0, ref0 = robot._mem.clone(), robot._ref.clone()
        x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=True).cpu().numpy()
        ts, table = c_port.table_struct("a1_px100_pin_ver")
        ref = c_port.step(ts, c_port.config_struct(robot, table), q, targets.cpu().numpy(), mem0.cpu().numpy(),
                          ref0.cpu().numpy(), robot.dt)
        assert (robot.last_status.cpu().numpy() == 0).all() and (ref["status"] == 0).all()
        assert np.abs(x - ref["qdot"]).max() < QP_TOL
        assert (robot.last_iters.cpu().numpy() == ref["iters"]).all()
        act = robot.last_active_set.cpu().numpy().astype(np.uint64)
        assert (act == ref["active_set"]).all()
        assert np.abs(robot._mem.cpu().numpy() - ref["mem_out"]).max() < 1e-12


def test_closed_loop_rollout_matches_oracle():
    """BASELINE config 5 in miniature: Euler-integrated closed loop (integrate + base estimate + second FK pass,
    Robot_Wrapper4.py:1397-1402, 1297-1327) over K ticks; every tick's configuration compared with the oracle loop."""
    name, N, K = "a1_wx200", 24, 8
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260005, 5e-4)
    rng = np.random.default_rng(5)
    drift = torch.as_tensor(rng.normal(0, 2e-4, size=(K, N, 18)), device="cuda:0").cumsum(0)   # slowly moving targets
    traj = targets[None] + drift
    mem0, ref0 = robot._mem.clone().cpu().numpy(), robot._ref.clone().cpu().numpy()
    qh, vh, sh = robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18], record=True)
    assert (sh == 0).all()
    # the single-call C-ABI rollout (wbc_rollout: one persistent launch for the horizon, state advanced in place) lands on the same state
    twin = _robot(name, N, P1_TASKS, P2_CONS, True)
    twin.current_joint_config = torch.as_tensor(q, device="cuda:0").clone()
    twin._mem.copy_(torch.as_tensor(mem0, device="cuda:0")); twin._ref.copy_(torch.as_tensor(ref0, device="cuda:0"))
    v_last = twin.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18])
    assert torch.equal(twin.current_joint_config, qh[-1]) and torch.equal(v_last, vh[-1]) and torch.equal(twin._mem, robot._mem)
    qh, vh = qh.cpu().numpy(), vh.cpu().numpy()
    rm = H.make_oracle(name, like=robot, dt=robot.dt)
    traj_h = traj.cpu().numpy()
    worst_q = worst_v = 0.0
    for s in range(0, N, 3):
        qs, mem = q[s].copy(), mem0[s].copy()
        for k in range(K):
            r = H.oracle_step_one(rm, qs, traj_h[k, s], mem, ref0[s], imu=None, solve=True, tail=True)
            worst_v = max(worst_v, np.abs(vh[k, s] - r["qdot"]).max())
            worst_q = max(worst_q, np.abs(qh[k, s] - r["q_next"]).max())
            qs, mem = r["q_next"], H.get_oracle_mem(rm)
    assert worst_v < QP_TOL and worst_q < 1e-8, (worst_v, worst_q)


@pytest.mark.parametrize("name", ["a1_wx200", "a1_px100_pin_ver"])
def test_closed_loop_reach_trajectory_matches_oracle(name):
    """A horizon of the length and shape the reference's driver runs (sim3.py:287-327: a standing robot, feet planted, the
    gripper reaching along a trajectory while the trunk sways): 120 closed-loop ticks with the IMU quaternion fed back, on
    the standing-pose sampler.  Every tick of every robot is compared with the oracle loop: solution, configuration, status,
    iteration count and active set -- so bounds and trunk-box rows that become active and inactive along the way do so at
    the same tick on both sides.  The horizon runs as one persistent launch (wbc_rollout) and tick by tick."""
    from wbc_b200 import synthetic
    N, K = 6, 120
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q = synthetic.sample_standing(robot.robot_model, N, 20260061)
    targets = synthetic.load_batch(robot, q, np.zeros((N, 18)))
    mem0, ref0 = robot._mem.clone().cpu().numpy(), robot._ref.clone().cpu().numpy()
    t = torch.arange(1, K + 1, dtype=torch.float64, device="cuda:0")[:, None] * robot.dt
    amp = torch.as_tensor(np.random.default_rng(7).uniform(0.5, 1.0, size=(N,)), device="cuda:0")[None, :]
    traj = targets[None].repeat(K, 1, 1)
    traj[:, :, 12] += 0.10 * amp * torch.sin(2 * np.pi * 2.0 * t)           # gripper x: 10 cm reach, 2 Hz
    traj[:, :, 14] += 0.06 * amp * (1 - torch.cos(2 * np.pi * 2.0 * t))     # gripper z
    traj[:, :, 16] += 0.02 * amp * torch.sin(2 * np.pi * 1.0 * t)           # trunk sway y
    traj[:, :, 17] -= 0.03 * amp * (1 - torch.cos(2 * np.pi * 1.0 * t))     # trunk z: crouch towards the trunk box
    imu = torch.as_tensor(q[:, 3:7], device="cuda:0")[None].repeat(K, 1, 1)
    ee, tr = traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18]
    qh, vh, sh = robot.rollout(ee, tr, imu_quat_traj=imu, record=True)
    twin = _robot(name, N, P1_TASKS, P2_CONS, True)
    twin.current_joint_config = torch.as_tensor(q, device="cuda:0").clone()
    twin._mem.copy_(torch.as_tensor(mem0, device="cuda:0")); twin._ref.copy_(torch.as_tensor(ref0, device="cuda:0"))
    twin.rollout(ee, tr, imu_quat_traj=imu)
    assert torch.equal(twin.current_joint_config, qh[-1]) and torch.equal(twin._mem, robot._mem)
    assert (sh == 0).all()
    # per-tick iteration counts / active sets of the device loop (record=True steps tick by tick)
    robot.current_joint_config = torch.as_tensor(q, device="cuda:0").clone()
    robot._mem.copy_(torch.as_tensor(mem0, device="cuda:0"))
    its, acts = [], []
    for k in range(K):
        robot.step(ee[k], tr[k], imu_quat=imu[k], advance=True)
        its.append(robot.last_iters.cpu().numpy().copy()); acts.append(robot.last_active_set.cpu().numpy().astype(np.uint64).copy())
    assert torch.equal(robot.current_joint_config, qh[-1])
    qh, vh = qh.cpu().numpy(), vh.cpu().numpy()
    rm = H.make_oracle(name, like=robot, dt=robot.dt)
    traj_h, imu_h = traj.cpu().numpy(), imu.cpu().numpy()
    worst_q = worst_v = 0.0
    n_ineq = 0
    for s in range(N):
        qs, mem = q[s].copy(), mem0[s].copy()
        for k in range(K):
            r = H.oracle_step_one(rm, qs, traj_h[k, s], mem, ref0[s], imu=imu_h[k, s], solve=True, tail=True)
            assert r["status"] == 0, (s, k)
            worst_v = max(worst_v, np.abs(vh[k, s] - r["qdot"]).max())
            worst_q = max(worst_q, np.abs(qh[k, s] - r["q_next"]).max())
            wb, wr = H.act_to_bits(r["act"], robot.n_velocity_dimensions)
            assert r["iters"] == int(its[k][s]) and wb == int(acts[k][s, 0]) and wr == int(acts[k][s, 1]), (s, k)
            n_ineq += r["iters"] - 15
            qs, mem = r["q_next"], H.get_oracle_mem(rm)
    print(name, "reach trajectory: worst |dv|", worst_v, "worst |dq|", worst_q, "inequality iterations", n_ineq)
    assert worst_v < QP_TOL and worst_q < 1e-8, (worst_v, worst_q)
    assert n_ineq > 0                  # the horizon does drive constraints active


def test_config5_full_size_with_infeasible_states(monkeypatch):
    """BASELINE config 5 at full size -- 16,384 robots x 100 closed-loop ticks -- on the RANDOM-attitude sampler, where the
    reference's base estimator (quirk D.10) walks a few per cent of the robots into infeasible QPs.  The horizon must stay
    finite and the statuses must say which robots failed; the reduced front and the general front must tell the same story:
    every robot fails for the first time at the same tick on both (what happens to a robot AFTER its first infeasible QP is
    not comparable: the iterate a front holds when it proves infeasibility is its own), and the robots that never fail end
    on the same configuration with the same iteration counts (median difference 3e-15; a few robots in a thousand run close
    enough to the edge for the fronts' rounding to send them apart -- the closed loop of quirk D.10 is not contractive
    there -- so the assertion is on 99.5 %).  The one-launch horizon (wbc_rollout) ends where the
    tick-by-tick loop ends.  Then the per-tick iteration budget: with a cap of 40 no QP runs longer, capped states are
    flagged."""
    name, N, K = "a1_wx200", 16384, 100
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(20260008)
    drift = torch.zeros(K, N, 18, dtype=torch.float64, device="cuda:0")
    drift[:, :, 12:18] = torch.randn(K, N, 6, dtype=torch.float64, device="cuda:0", generator=gen).mul_(1e-4).cumsum(0)
    ends = {}
    for label, off in (("reduced", "0"), ("general", "1")):
        monkeypatch.setenv("WBC_B200_NO_REDUCED", off)
        robot = _robot(name, N, P1_TASKS, P2_CONS, True)
        q, targets = _load(robot, N, 20260003, 5e-4)
        q0, mem0 = robot.current_joint_config.clone(), robot._mem.clone()
        traj = targets[None] + drift
        ee, tr = traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18]
        qh, vh, sh = robot.rollout(ee, tr, record=True)
        assert torch.isfinite(qh).all() and torch.isfinite(vh).all()
        bad = sh != 0
        first_bad = torch.where(bad.any(dim=0), bad.int().argmax(dim=0), torch.full((N,), K, device="cuda:0"))
        ends[label] = (qh[-1].clone(), first_bad, robot.last_iters.clone())
        if label == "reduced":                   # the whole horizon as one persistent launch
            robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0)
            robot.rollout(ee, tr)
            assert torch.equal(robot.current_joint_config, qh[-1]) and torch.equal(robot.last_status, sh[-1])
        del qh, vh, sh
    monkeypatch.delenv("WBC_B200_NO_REDUCED")
    (qa, fa, ia), (qb, fb, ib) = ends["reduced"], ends["general"]
    never = fa == K
    nbad = int((~never).sum())
    print("config 5, random attitudes: robots that meet an unsolved QP within the horizon:", nbad, "of", N)
    assert 0 < nbad < 0.1 * N                    # the sampler does produce infeasible robots, and only a few per cent
    assert torch.equal(fa, fb)
    # (the two fronts differ by rounding, ~3e-9 per solve; over 100 closed-loop ticks that may flip a near-tie of the
    #  pivoting rule on a handful of robots -- an extra add / drop pair, same minimiser)
    same_iters = float((ia[never] == ib[never]).double().mean())
    dq = (qa[never] - qb[never]).abs().amax(dim=1)
    close = float((dq < 1e-6).double().mean())
    print("never-failing robots: identical last-tick iteration count on", same_iters, " |dq| < 1e-6 after", K, "ticks on", close,
          " outliers:", int((dq >= 1e-6).sum()), "worst", float(dq.max()), "median", float(dq.median()))
    assert same_iters > 0.995 and close > 0.995
    # iteration budget per tick
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    robot.max_qp_iterations = 40
    q, targets = _load(robot, N, 20260003, 5e-4)
    traj = targets[None] + drift
    robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18])
    assert int(robot.last_iters.max()) <= 40 and torch.isfinite(robot.current_joint_config).all()
    capped = (robot.last_status & 1) != 0
    assert bool((robot.last_iters[capped] == 40).all())


def test_config3_extension_rows_full_size():
    """BASELINE config 3: A1 + WX200, 65,536 states, friction-pyramid + torque-limit proxy rows through the generic
    extension-row channel (NOT in the reference, whose QP is purely kinematic: parity unpinned by construction; both
    oracles implement the same channel, `synthetic.config3_rows` defines the rows).  4 trunk + 16 pyramid + 7 power
    rows = 27 rows exercise the full-width (nC > 16) solver layout at a mean of ~27 working-set changes per state.
    EVERY state is compared with the C oracle: same minimiser, same pivoting path, same active set -- the pyramid's
    mirrored faces tie exactly once their partners are active, which is what the shared tie window of the entering
    rule (wbc_qp.cuh / oracle/qp_wrapper.py) is for; without it 5 % of the states ended on another (equally valid)
    active set of the same degenerate vertex."""
    import wbc_b200
    from wbc_b200 import synthetic
    from oracle import c_port
    N = 65536
    robot = _robot("a1_wx200", N, P1_TASKS, dict(CoM=False, Trunk=True, FR=False, FL=False, RR=False, RL=False, Grip=False), True)
    robot.extra_rows = synthetic.config3_rows(robot.robot_model)
    q, targets = _load(robot, N, 20260003, 5e-3)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    asm = robot.assemble(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], want=("C", "Clb", "Cub", "lb", "ub"))
    assert asm["C"].shape[1] == 27
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False)
    assert (robot.last_status == 0).all()
    Cx = torch.einsum("nrk,nk->nr", asm["C"], x)
    assert (Cx >= asm["Clb"] - 1e-7).all() and (Cx <= asm["Cub"] + 1e-7).all()
    assert (x >= asm["lb"] - 1e-9).all() and (x <= asm["ub"] + 1e-9).all()
    act_rows = robot.last_active_set[:, 1]
    pyramid = ((act_rows >> 8) & 0xFFFFFFFF) != 0
    power = ((act_rows >> 40) & 0x3FFF) != 0
    assert float(pyramid.double().mean()) > 0.5 and float(power.double().mean()) > 0.5     # both families do bind
    ts, table = c_port.table_struct("a1_wx200")
    ref = c_port.step(ts, c_port.config_struct(robot, table), q, targets.cpu().numpy(), mem0.cpu().numpy(),
                      ref0.cpu().numpy(), robot.dt)
    assert (ref["status"] == 0).all()
    assert np.abs(x.cpu().numpy() - ref["qdot"]).max() < QP_TOL
    same_path = robot.last_iters.cpu().numpy() == ref["iters"]
    same_set = (robot.last_active_set.cpu().numpy().astype(np.uint64) == ref["active_set"]).all(axis=1)
    print("config3: identical pivoting path", same_path.mean(), "identical active set", same_set.mean(),
          "mean iterations", ref["iters"].mean())
    assert same_path.all() and same_set.all(), (int((~same_path).sum()), int((~same_set).sum()))
    # the NumPy / SciPy oracle (dense re-solves, no factor updating) on a sample of the same states
    rm = H.make_oracle("a1_wx200", like=robot, dt=robot.dt)
    tg, m0, r0 = targets.cpu().numpy(), mem0.cpu().numpy(), ref0.cpu().numpy()
    act = robot.last_active_set.cpu().numpy().astype(np.uint64)
    for s in range(0, N, 1024):
        r = H.oracle_step_one(rm, q[s], tg[s], m0[s], r0[s], solve=True, tail=False)
        wb, wr = H.act_to_bits(r["act"], robot.n_velocity_dimensions)
        assert int(act[s, 0]) == wb and int(act[s, 1]) == wr and r["iters"] == int(robot.last_iters[s]), s


@pytest.mark.parametrize("name", ["a1_wx200", "a1_px100_pin_ver"])
@pytest.mark.parametrize("kind", ["ineq_rows", "eq_row_fallback", "general_front"])
def test_reduced_front_against_c_oracle_and_general_front(name, kind, monkeypatch):
    """The null-space front of the QP (csrc/wbc_qp_red.inc: foot rows eliminated up front) against the C oracle, which
    enters the twelve rows one by one: same minimiser, same iteration count, same active set.  `ineq_rows`: feet + two
    extra inequality rows on the trunk frame (rows without leg support ride along as column selections);
    `eq_row_fallback`: one extra row with lo == hi makes the equality set differ from the foot rows, so every state
    takes the in-kernel fallback to the general front; `general_front`: the same problem with the reduced front
    switched off (WBC_B200_NO_REDUCED=1), i.e. the A/B baseline."""
    import wbc_b200
    from oracle import c_port
    if kind == "general_front":
        monkeypatch.setenv("WBC_B200_NO_REDUCED", "1")
    N = 2048
    robot = _robot(name, N, P1_TASKS, dict(CoM=False, Trunk=False, FR=True, FL=True, RR=True, RL=True, Grip=False), True)
    rf = wbc_b200._cabi.RF_LOCAL_WORLD_ALIGNED
    if kind == "eq_row_fallback":
        robot.extra_rows = [(5, rf, [0, 0, 1, 0, 0, 0], 0.01, 0.01), (5, rf, [0, 0, 0, 0, 0, 1], -0.2, 0.2)]
    else:
        robot.extra_rows = [(5, rf, [0, 0, 1, 0, 0, 0], -0.05, 0.05), (5, rf, [0, 0, 0, 0, 0, 1], -0.2, 0.2)]
    q, targets = _load(robot, N, 20260007, 5e-3)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False)
    assert (robot.last_status == 0).all()
    ts, table = c_port.table_struct(name)
    ref = c_port.step(ts, c_port.config_struct(robot, table), q, targets.cpu().numpy(), mem0.cpu().numpy(),
                      ref0.cpu().numpy(), robot.dt)
    assert (ref["status"] == 0).all()
    assert np.abs(x.cpu().numpy() - ref["qdot"]).max() < QP_TOL
    same_path = robot.last_iters.cpu().numpy() == ref["iters"]
    same_set = (robot.last_active_set.cpu().numpy().astype(np.uint64) == ref["active_set"]).all(axis=1)
    assert float((robot.last_active_set[:, 1] & 0xF000000).ne(0).double().mean()) > 0.05     # the extra rows do bind
    assert same_path.mean() > 0.995 and same_set.mean() > 0.995, (same_path.mean(), same_set.mean())


def test_reduced_front_falls_back_on_singular_legs(monkeypatch):
    """States whose leg Jacobian is (nearly) singular -- a straight knee, calf angle 0 -- fail the conditioning check of
    the reduced front and take the general front inside the same launch; every state, singular or not, must agree with
    the launch that has the reduced front switched off."""
    N = 1024
    name = "a1_wx200"
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260013, 5e-3)
    qd = robot.current_joint_config.clone()
    calf = [robot.robot_model.idx_q[robot.robot_model.getJointId(n)] for n in ("FL_calf_joint", "RR_calf_joint")]
    qd[::7, calf[0]] = 0.0                         # every 7th state: FL knee straight
    qd[3::11, calf[1]] = 1e-7                      # some: RR knee within 1e-7 rad of straight
    robot.current_joint_config.copy_(qd)            # (task memory / references stay those of the sampled states)
    ee, tr = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    x_red = robot.step(ee, tr, advance=False).clone()
    st_red, it_red, act_red = robot.last_status.clone(), robot.last_iters.clone(), robot.last_active_set.clone()
    monkeypatch.setenv("WBC_B200_NO_REDUCED", "1")
    x_gen = robot.step(ee, tr, advance=False).clone()
    assert torch.equal(st_red, robot.last_status)
    ok = robot.last_status == 0                    # (a straight knee makes the foot rows dependent: the general front
    assert ok.double().mean() > 0.8                #  may flag such states; both launches must flag the same ones)
    assert (x_red[ok] - x_gen[ok]).abs().max() < 1e-7
    assert torch.equal(it_red[ok], robot.last_iters[ok]) and torch.equal(act_red[ok], robot.last_active_set[ok])


def test_bootstrap_matches_oracle():
    """f4: the constructor bootstrap (setInitialState, Robot_Wrapper4.py:196-351) batched -- 60 of its 2000 ticks
    (linear EE trajectories, bounds-only QP, plain integrate) and the final re-basing, against the oracle."""
    import wbc_b200
    from oracle.robot_wrapper4 import RobotModel as ORM
    name, K = "a1_wx200", 60
    robot = wbc_b200.RobotModel(name, batch=3, device="cuda:0")
    qf = robot.setInitialState(bootstrap_steps=K).cpu().numpy()
    orm = ORM(H.oracle_model(name), run_bootstrap=True, bootstrap_steps=K)     # the constructor runs it (:161), as the reference does
    assert (robot.last_status == 0).all()
    assert np.abs(qf - np.asarray(orm.current_joint_config)[None]).max() < 1e-8
    assert np.abs(robot.FL_leg.cpu().numpy() - np.asarray(orm.FL_leg)[None]).max() < 1e-8
    assert np.abs(robot.grip.cpu().numpy() - np.asarray(orm.grip)[None]).max() < 1e-8


def test_com_rows_match_oracle():
    """f3: CoMConstraint (Robot_Wrapper4.py:669-694) = rows x, y of jacobianCenterOfMass with the box between the RR and
    FL feet.  The reference's recorded CoM Jacobian (Jacobians.py:27-42) is stale against the current URDF masses, so
    the pin is the oracle (whose CoM Jacobian is checked against the derivative of the CoM on the CPU).  Tilted bases
    can make the box empty: both sides must then agree that the QP is infeasible."""
    from oracle import c_port
    name, N = "a1_wx200", 64
    cons = dict(CoM=True, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
    robot = _robot(name, N, P1_TASKS, cons, True)
    q, targets = _load(robot, N, 20260002, 5e-4)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    asm = robot.assemble(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], want=("C", "Clb", "Cub"))
    ref = H.oracle_step_batch(name, robot, q, targets.cpu().numpy(), mem0.cpu().numpy(), ref0.cpu().numpy(), solve=False)
    assert asm["C"].shape[1] == 18
    assert _maxabs(asm["C"].cpu().numpy() - ref["C"]) < FK_TOL
    for k in ("Clb", "Cub"):
        assert _maxabs(asm[k].cpu().numpy() - ref[k]) < 1e-9 * max(1.0, _maxabs(ref[k]))
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False).cpu().numpy()
    ts, table = c_port.table_struct(name)
    cref = c_port.step(ts, c_port.config_struct(robot, table), q, targets.cpu().numpy(), mem0.cpu().numpy(),
                       ref0.cpu().numpy(), robot.dt)
    st = robot.last_status.cpu().numpy()
    assert ((st & 2) == (cref["status"] & 2)).all()
    ok = st == 0
    assert ok.sum() >= N // 4
    assert np.abs(x[ok] - cref["qdot"][ok]).max() < QP_TOL
    # Infeasible QPs have a DEFINED output too: the iterate the dual method held when it proved that no step exists
    # (the reference uses whatever qpOASES leaves in its primal vector, QP_Wrapper.py:45-51).  Kernel and oracle walk the
    # same path, so they stop at the same iteration on the same point.
    bad = (st & 2) != 0
    if bad.any():
        assert (robot.last_iters.cpu().numpy()[bad] == cref["iters"][bad]).all()
        scale = max(1.0, float(np.abs(cref["qdot"][bad]).max()))
        assert np.abs(x[bad] - cref["qdot"][bad]).max() < QP_TOL * scale


@pytest.mark.parametrize("resident", [True, False])
def test_step_host_matches_device_step(resident):
    """wbc_step_host (host buffers in, host buffers out, sliced copy / compute pipeline inside the C ABI) returns exactly
    what the device-resident tick returns, for a batch that does not divide into the slices and for more slices than
    three streams."""
    name, N = "a1_wx200", 4099
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260011, 5e-3)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False).clone()
    st, it = robot.last_status.clone(), robot.last_iters.clone()
    host_in = {"q": robot.current_joint_config.cpu().pin_memory(), "targets": targets.cpu().pin_memory(),
               "mem": mem0.cpu().pin_memory(), "ref": ref0.cpu().pin_memory()}
    host_out = {"qdot": torch.full((N, robot.n_velocity_dimensions), float("nan"), dtype=torch.float64).pin_memory(),
                "status": torch.full((N,), -1, dtype=torch.int32).pin_memory(),
                "iters": torch.full((N,), -1, dtype=torch.int32).pin_memory()}
    if not resident:                     # everything travels: scramble the device copies first
        robot.current_joint_config.zero_(); robot._targets.zero_(); robot._mem.zero_(); robot._ref.zero_()
    else:                                # task memory / references resident; q and targets travel
        robot.current_joint_config.zero_(); robot._targets.zero_()
    for chunks in (0, 1, 7):              # zero-copy, one staged slice, more slices than streams
        host_out["qdot"].fill_(float("nan"))
        h2d, d2h = robot.step_host(host_in, host_out, chunks=chunks, resident_state=resident)
        torch.cuda.synchronize()
        assert torch.equal(host_out["qdot"], x.cpu())
        assert torch.equal(host_out["status"], st.cpu()) and torch.equal(host_out["iters"], it.cpu())
        assert h2d == N * 8 * ((27 + 18) if resident else (27 + 18 + 72 + 24)) and d2h == N * (26 * 8 + 8)


def test_ragged_and_empty_batches():
    """Batch-size invariance: a state's answer does not depend on how many other states ride in the launch
    (1, 13, one more than a full wave of 148 x 12 warps), and an empty batch is a no-op."""
    import ctypes as C
    import wbc_b200
    from wbc_b200 import _cabi as cabi
    name = "a1_wx200"
    big_n = 148 * 12 + 1
    big = _robot(name, big_n, P1_TASKS, P2_CONS, True)
    q, targets = _load(big, big_n, 20260007, 5e-3)
    mem0, ref0 = big._mem.clone(), big._ref.clone()
    xb = big.step(targets[:, :15].reshape(big_n, 5, 3), targets[:, 15:18], advance=False).clone()
    itb, actb = big.last_iters.clone(), big.last_active_set.clone()
    assert (big.last_status == 0).all()
    for n in (1, 13):
        small = _robot(name, n, P1_TASKS, P2_CONS, True)
        small.current_joint_config.copy_(big.current_joint_config[-n:])
        small._mem.copy_(mem0[-n:]); small._ref.copy_(ref0[-n:])
        xs = small.step(targets[-n:, :15].reshape(n, 5, 3), targets[-n:, 15:18], advance=False)
        assert torch.equal(xs, xb[-n:])                       # bit-identical: same code path, same data
        assert torch.equal(small.last_iters, itb[-n:]) and torch.equal(small.last_active_set, actb[-n:])
    # N = 0 through the C ABI: accepted, nothing launched
    io = big._io(targets=targets, qdot=big.qdot, status=big.last_status, iters=big.last_iters)
    cfg = big._config()
    assert cabi.load().wbc_step(big._model, C.byref(cfg), C.byref(io), 0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(big.qdot, xb)


def test_general_placements_and_axes(tmp_path):
    """The A1 URDFs only have pure-translation placements and unit axes, which take exact shortcuts in the kernels.
    A perturbed copy of the tree (rotated joint placements and frame offsets, a skew revolute axis, as in
    laikago_vx300.urdf) drives the general code paths; FK, Jacobians and one full tick against the oracle."""
    import json
    import wbc_b200
    from scipy.spatial.transform import Rotation as R
    d = json.load(open(H.table_path("a1_wx200")))
    rng = np.random.default_rng(12)
    for j in (3, 6, 9, 15, 17):                                   # thighs and two arm joints: rotated placements
        d["joints"][j]["R"] = R.from_euler("xyz", rng.uniform(-0.4, 0.4, 3)).as_matrix().reshape(-1).tolist()
    ax = np.array([0.3, -0.5, 0.8]); ax /= np.linalg.norm(ax)
    d["joints"][16]["axis"] = ax.tolist()                          # elbow about a skew axis
    for f in d["frames"]:
        if f["name"] in ("FR_foot_fixed", "gripper_bar", "imu_joint"):
            f["R"] = R.from_euler("xyz", rng.uniform(-0.5, 0.5, 3)).as_matrix().reshape(-1).tolist()
    path = tmp_path / "a1_wx200_general.json"
    json.dump(d, open(path, "w"))
    N = 64
    robot = wbc_b200.RobotModel(str(path), batch=N, device="cuda:0")
    robot.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
    robot.setConstraints(**P2_CONS)
    model = H.opin.Model.from_json(str(path))
    data = model.createData()
    q, targets = _load(robot, N, 31, 5e-4)
    got = {rf: robot.frameJacobians(rf) for rf in (0, 1, 2)}
    frames = robot.end_effector_index_list_frame + [robot.trunk_frame_index]
    worst = 0.0
    for s in range(0, N, 5):
        H.opin.forwardKinematics(model, data, q[s])
        H.opin.computeJointJacobians(model, data, q[s])
        H.opin.updateFramePlacements(model, data)
        for rf in (0, 1, 2):
            oMf, J = got[rf]
            for k, fid in enumerate(frames):
                worst = max(worst, np.abs(J[s, k].cpu().numpy() - H.opin.getFrameJacobian(model, data, fid, rf)).max())
                ref = np.concatenate([data.oMf[fid].rotation.reshape(-1), data.oMf[fid].translation])
                worst = max(worst, np.abs(oMf[s, k].cpu().numpy() - ref).max())
    assert worst < FK_TOL, worst
    # one fused tick on the general tree
    from oracle.robot_wrapper4 import RobotModel as ORM
    mem0, ref0 = robot._mem.clone().cpu().numpy(), robot._ref.clone().cpu().numpy()
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False).cpu().numpy()
    rm = ORM(model, dt=robot.dt)
    H.copy_settings(robot, rm)
    tg = targets.cpu().numpy()
    for s in range(0, N, 7):
        r = H.oracle_step_one(rm, q[s], tg[s], mem0[s], ref0[s], solve=True, tail=False)
        assert int(robot.last_status[s]) == int(r["status"])
        if r["status"] == 0:
            assert np.abs(x[s] - r["qdot"]).max() < QP_TOL
            assert int(robot.last_iters[s]) == r["iters"]
