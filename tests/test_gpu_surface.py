"""-m gpu: the drop-in SURFACE of the reference, method by method, through the CUDA path against the CPU oracle.

tests/test_gpu_parity.py drives the fused tick through `assemble` / `step` / `rollout`; this module calls what a user
of the reference calls -- `initialiseWBC`, `runWBC`, `qpA`, `qpb`, `endEffectorA2`, `trunkA`,
`velDamperJointConstraints`, `EEConstraint`, `trunkConstraint`, `CoMConstraint`, `findConstraints` (which returns `C.T`),
`jointVelocitiestoConfig`, `updateState(running=True)`, `trunkWorldPos`, `staticReachMode`, `QP.solveQP`,
`QP.solveQPHotstart` (wrappers/Robot_Wrapper4.py, wrappers/QP_Wrapper.py) -- and compares each with the oracle method
of the same name on the same arrays, plus the configurations the reference can be switched into that the default-weight
tests never reach: non-identity weights, unequal gains (gain-list order quirk), the gripper constraint,
the velocity damper without its off-by-one, QPs that hit the iteration cap, a Hessian that is not positive definite.

Tolerances as in BASELINE.json north_star: FK / Jacobian quantities 1e-10 abs, QP solutions 1e-6 with an identical
active set, quantities that carry 1/dt = 500 relative 1e-9.
"""
import numpy as np
import pytest
import torch

from tests import helpers as H
from tests.test_gpu_parity import _robot, _load, _maxabs, P1_TASKS, P2_TASKS, P2_CONS, NO_CONS, FK_TOL, QP_TOL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _np(t):
    return t.detach().cpu().numpy()


def _uniform_quats(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _mem_of(rm):
    return H.get_oracle_mem(rm)


def _ref_of(rm):
    return np.concatenate([np.concatenate([np.reshape(e, 3) for e in rm.default_EE_ori_list]),
                           np.reshape(rm.default_trunk_ori, 3), np.reshape(rm.initial_trunk_pos, 3),
                           np.reshape(rm.initial_trunk_ori_euler, 3)])


# ------------------------------------------------------------------------------------------------ initialiseWBC
@pytest.mark.parametrize("name", ["a1_wx200", "a1_px100_pin_ver"])
def test_initialise_wbc_matches_oracle(name):
    """wbc_init_memory (initialiseWBC, Robot_Wrapper4.py:354-383: from_matrix -> as_euler('xyz'), R_trunk^T R_EE, the
    trunk snapshot) against the oracle's initialiseWBC on 256 random states, half of them with a uniform-SO(3) base.
    Every other test takes the task memory / references from this kernel and hands them to both sides, so an error
    here would cancel there."""
    import wbc_b200
    from wbc_b200 import synthetic
    N = 256
    robot = wbc_b200.RobotModel(name, batch=N, device=DEV)
    q = synthetic.sample_configurations(robot.robot_model, N, 77)
    q[N // 2:, 3:7] = _uniform_quats(np.random.default_rng(9), N - N // 2)
    qd = torch.as_tensor(q, device=DEV)
    robot.updateState(qd, feedback=False)
    robot.initialiseWBC(qd[:, 3:7])
    mem, ref = _np(robot._mem), _np(robot._ref)
    rm = H.make_oracle(name)
    worst_pos = worst_rot = worst_eul = 0.0
    checked = 0
    for s in range(N):
        rm.current_joint_config = q[s].copy()
        rm.updateState(q[s], feedback=False)
        rm.initialiseWBC(q[s, 3:7])
        om, orf = _mem_of(rm), _ref_of(rm)
        worst_pos = max(worst_pos, np.abs(mem[s, 0:15] - om[0:15]).max(), np.abs(mem[s, 60:63] - om[60:63]).max(),
                        np.abs(ref[s, 18:21] - orf[18:21]).max())
        worst_rot = max(worst_rot, np.abs(mem[s, 15:60] - om[15:60]).max(), np.abs(mem[s, 63:72] - om[63:72]).max())
        # Euler angles: as_euler is ill-conditioned next to the gimbal lock (second angle +-pi/2); compare away from it
        eul_o = np.concatenate([orf[0:18], orf[21:24]]).reshape(7, 3)
        eul_g = np.concatenate([ref[s, 0:18], ref[s, 21:24]]).reshape(7, 3)
        ok = np.abs(np.cos(eul_o[:, 1])) > 1e-3
        d = np.abs(eul_g[ok] - eul_o[ok])
        d = np.minimum(d, np.abs(d - 2 * np.pi))                      # the same angle either side of the +-pi cut
        if d.size:
            worst_eul = max(worst_eul, d.max())
            checked += int(ok.sum())
    assert worst_pos < FK_TOL and worst_rot < FK_TOL, (worst_pos, worst_rot)
    assert worst_eul < 1e-8 and checked > 6 * N, (worst_eul, checked)
    # the properties / accessors a caller reads (sim3.py:109-128, 164-172)
    assert torch.equal(robot.prev_EE_pos.reshape(N, 15), robot._mem[:, :15])
    assert torch.equal(robot.initial_trunk_pos, robot._ref[:, 18:21])


# ------------------------------------------------------------------------------------------------ the method surface
def _oracle_at(rm, q, mem, ref):
    H.set_oracle_state(rm, q, mem, ref)
    return rm


def test_every_mirrored_method_matches_the_oracle_method():
    """Call the reference's method names on the batched mirror and on the oracle, state by state."""
    import wbc_b200
    name, N = "a1_wx200", 24
    cons = dict(CoM=True, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
    robot = _robot(name, N, P1_TASKS, cons, "PREV")
    q, targets = _load(robot, N, 20260021, 5e-4)
    mem0, ref0 = _np(robot._mem).copy(), _np(robot._ref).copy()
    tg = _np(targets)
    ee_t, tr_t = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    nv, nq = robot.n_velocity_dimensions, robot.n_configuration_dimensions
    rm = H.make_oracle(name, like=robot, dt=robot.dt)

    # -- A rows ------------------------------------------------------------------------------------
    for i in range(5):
        robot.endEffectorA2(i)
    robot.trunkA()
    A = _np(robot.qpA())
    lb, ub = (_np(x) for x in robot.velDamperJointConstraints())
    Ct, Clb, Cub = robot.findConstraints()
    assert tuple(Ct.shape) == (N, nv, 18) and tuple(Clb.shape) == (N, 18)
    assert Ct.stride(1) == 1 and Ct.stride(2) == nv          # the transposed VIEW of C, as the reference's C.T (:836)
    per = {"ee": [robot.EEConstraint(i) for i in range(4)], "trunk": robot.trunkConstraint(), "com": robot.CoMConstraint()}
    worst = {}

    def upd(key, got, want, rel=False):
        d = _maxabs(np.asarray(got) - np.asarray(want))
        if rel:
            d /= max(1.0, _maxabs(want))
        worst[key] = max(worst.get(key, 0.0), d)

    for s in range(N):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        ee = [tg[s, 3 * i:3 * i + 3].reshape(3, 1) for i in range(5)]
        rm.FR_target_cartesian_pos, rm.FL_target_cartesian_pos, rm.RR_target_cartesian_pos, rm.RL_target_cartesian_pos = ee[:4]
        for i in range(5):
            rm.endEffectorA2(i)
            upd("endEffectorA2", _np(robot.EE_A_list[i][s]), rm.EE_A_list[i])
        rm.trunkA()
        upd("trunkA", _np(robot.trunk_A[s]), rm.trunk_A)
        upd("qpA", A[s], rm.qpA())
        olb, oub = rm.velDamperJointConstraints()
        upd("lb", lb[s], olb); upd("ub", ub[s], oub)
        oCt, oClb, oCub = rm.findConstraints()
        upd("findConstraints.C", _np(Ct[s]), oCt)
        upd("findConstraints.Clb", _np(Clb[s]), oClb, rel=True); upd("findConstraints.Cub", _np(Cub[s]), oCub, rel=True)
        for i in range(4):
            oC, ol, ou = rm.EEConstraint(i)
            upd("EEConstraint", _np(per["ee"][i][0][s]), oC)
            assert not _np(per["ee"][i][1][s]).any() and not _np(per["ee"][i][2][s]).any()
        oC, ol, ou = rm.trunkConstraint()
        upd("trunkConstraint.C", _np(per["trunk"][0][s]), oC)
        upd("trunkConstraint.Clb", _np(per["trunk"][1][s]), ol, rel=True); upd("trunkConstraint.Cub", _np(per["trunk"][2][s]), ou, rel=True)
        oC, ol, ou = rm.CoMConstraint()
        upd("CoMConstraint.C", _np(per["com"][0][s]), oC)
        upd("CoMConstraint.Clb", _np(per["com"][1][s]), ol, rel=True); upd("CoMConstraint.Cub", _np(per["com"][2][s]), ou, rel=True)
    for k, v in worst.items():
        assert v < (1e-9 if k.endswith(("Clb", "Cub")) else FK_TOL), (k, v)

    # -- qpb mutates the task memory (:995-996, :1151-1152) -----------------------------------------
    b = _np(robot.qpb(ee_t, tr_t))
    assert b.shape == (N, A.shape[1], 1)
    mem1 = _np(robot._mem)
    for s in range(N):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        ob = rm.qpb([tg[s, 3 * i:3 * i + 3].reshape(3, 1) for i in range(5)], tg[s, 15:18].reshape(3, 1))
        assert _maxabs(b[s] - ob) < 1e-9 * max(1.0, _maxabs(ob)), s
        assert _maxabs(mem1[s] - _mem_of(rm)) < 1e-12, s

    # -- trunkWorldPos (:1297-1327) reads the foot targets of the last tick (FR_target_cartesian_pos ..., :1332-1335) ----
    base = _np(robot.trunkWorldPos())
    for s in range(N):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        ee = [tg[s, 3 * i:3 * i + 3].reshape(3, 1) for i in range(5)]
        rm.FR_target_cartesian_pos, rm.FL_target_cartesian_pos, rm.RR_target_cartesian_pos, rm.RL_target_cartesian_pos = ee[:4]
        assert _maxabs(base[s] - rm.trunkWorldPos()) < FK_TOL, s

    # -- jointVelocitiestoConfig (:440-449), both flavours; updateState(running=True) (:387-428) ------
    rng = np.random.default_rng(4)
    v = rng.normal(0, 0.3, size=(N, nv))
    q_int = _np(robot.jointVelocitiestoConfig(torch.as_tensor(v, device=DEV), update_model=False))
    imu = _uniform_quats(rng, N)
    for s in range(N):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        assert _maxabs(q_int[s] - rm.jointVelocitiestoConfig(v[s], False)) < 1e-12, s
    robot.updateState(torch.as_tensor(q_int[:, 7:], device=DEV), torch.as_tensor(imu, device=DEV), running=True)
    q_run = _np(robot.current_joint_config)
    oMf_trunk = _np(robot.robot_data.oMf[robot.trunk_frame_index].translation)
    for s in range(N):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        ee = [tg[s, 3 * i:3 * i + 3].reshape(3, 1) for i in range(5)]
        rm.FR_target_cartesian_pos, rm.FL_target_cartesian_pos, rm.RR_target_cartesian_pos, rm.RL_target_cartesian_pos = ee[:4]
        rm.updateState(q_int[s, 7:], imu[s], running=True)
        assert _maxabs(q_run[s] - rm.current_joint_config) < FK_TOL, s
        assert _maxabs(oMf_trunk[s] - rm.robot_data.oMf[rm.trunk_frame_index].translation) < FK_TOL, s
    # update_model=True on an initialised controller = integrate + updateState(running=True) with feedback off (:443-445)
    robot.updateState(torch.as_tensor(q, device=DEV), feedback=False)
    assert robot.jointVelocitiestoConfig(torch.as_tensor(v, device=DEV), update_model=True) is None
    q_upd = _np(robot.current_joint_config)
    for s in range(0, N, 3):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        ee = [tg[s, 3 * i:3 * i + 3].reshape(3, 1) for i in range(5)]
        rm.FR_target_cartesian_pos, rm.FL_target_cartesian_pos, rm.RR_target_cartesian_pos, rm.RL_target_cartesian_pos = ee[:4]
        rm.jointVelocitiestoConfig(v[s], True)
        assert _maxabs(q_upd[s] - rm.current_joint_config) < FK_TOL, s


@pytest.mark.parametrize("joint", [True, "HYBRID"])
def test_run_wbc_three_ticks_matches_oracle_run_wbc(joint):
    """runWBC itself (Robot_Wrapper4.py:1330-1412): first tick QP(...).solveQP(), later ticks solveQPHotstart, the
    mutated task memory, integrate, IMU feedback, base re-estimate -- three consecutive ticks with moving targets and a
    changing IMU quaternion, against the oracle's runWBC driven the same way.  "HYBRID" is what sim3.py:145 selects."""
    name, N, K = "a1_px100_pin_ver", 12 if joint is True else 6, 3
    tasks = P1_TASKS if joint is True else P2_TASKS
    robot = _robot(name, N, tasks, P2_CONS, joint)
    q, targets = _load(robot, N, 20260023, 5e-4)
    mem0, ref0 = _np(robot._mem).copy(), _np(robot._ref).copy()
    rng = np.random.default_rng(8)
    drift = rng.normal(0, 2e-4, size=(K, N, 18)).cumsum(0)
    from scipy.spatial.transform import Rotation as R
    imus = np.stack([(R.from_quat(q[:, 3:7]) * R.from_rotvec(rng.normal(0, 2e-3, size=(N, 3)))).as_quat() for _ in range(K)])
    tg = _np(targets)
    got = []
    for k in range(K):
        t = torch.as_tensor(tg + drift[k], device=DEV)
        legs = robot.runWBC(torch.as_tensor(imus[k], device=DEV), t[:, :15].reshape(N, 5, 3), t[:, 15:18])
        assert len(legs) == 5 and tuple(legs[0].shape) == (N, 3) and legs[4].shape[1] == robot.n_configuration_dimensions - 19
        got.append((np.concatenate([_np(x) for x in legs], axis=1), _np(robot.qdot).copy(), _np(robot.last_status).copy(),
                    _np(robot.current_joint_config).copy(), _np(robot._mem).copy()))
    assert robot.firstQP is False
    rm = H.make_oracle(name, like=robot, dt=robot.dt)
    worst_v = worst_q = worst_m = 0.0
    for s in range(N):
        _oracle_at(rm, q[s], mem0[s], ref0[s])
        for k in range(K):
            t = tg[s] + drift[k, s]
            ee = [t[3 * i:3 * i + 3].reshape(3, 1) for i in range(5)]
            out = rm.runWBC(imus[k, s], ee, t[15:18].reshape(3, 1))
            assert rm.qp.result["status"] == got[k][2][s] == 0, (s, k)
            worst_v = max(worst_v, _maxabs(got[k][1][s] - rm.last["qdot"]))
            worst_q = max(worst_q, _maxabs(got[k][3][s] - rm.current_joint_config),
                          _maxabs(got[k][0][s] - np.concatenate([np.asarray(x) for x in out])))
            worst_m = max(worst_m, _maxabs(got[k][4][s] - _mem_of(rm)))
    assert worst_v < QP_TOL and worst_q < 1e-8 and worst_m < 1e-12, (worst_v, worst_q, worst_m)


def test_qp_wrapper_solve_and_hotstart_match_oracle():
    """QP(A, b, lb, ub, C, Clb, Cub).solveQP() then .solveQPHotstart(A', b', ...) (QP_Wrapper.py:10-73) with the
    matrices of two consecutive WBC ticks; C handed over as the reference does, as C.T; bounds-only QPs refuse to
    hot-start exactly where the reference exits (:57-59)."""
    import wbc_b200
    from oracle.qp_wrapper import QP as OQP
    name, N = "a1_wx200", 48
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260029, 5e-3)
    ee_t, tr_t = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    t2 = targets + 1e-3
    ticks = []
    for t in (targets, t2):
        A = robot.qpA()
        b = robot.qpb(t[:, :15].reshape(N, 5, 3), t[:, 15:18])
        Ct, Clb, Cub = robot.findConstraints()
        lb, ub = robot.velDamperJointConstraints()
        ticks.append((A, b, lb, ub, Ct, Clb, Cub))
    qp = wbc_b200.QP(*ticks[0], n_of_velocity_dimensions=robot.n_velocity_dimensions)
    x0 = _np(qp.solveQP()).copy()
    it0, st0 = _np(qp.iters).copy(), _np(qp.status).copy()
    x1 = _np(qp.solveQPHotstart(*ticks[1])).copy()
    it1, st1 = _np(qp.iters).copy(), _np(qp.status).copy()
    for s in range(N):
        a = [_np(x[s]) for x in ticks[0]]
        oq = OQP(a[0], a[1].reshape(-1), a[2], a[3], a[4], a[5], a[6], n_of_velocity_dimensions=a[0].shape[1])
        ox0 = oq.solveQP()
        assert oq.result["status"] == st0[s] == 0 and oq.result["iters"] == it0[s]
        assert _maxabs(x0[s] - ox0) < QP_TOL
        a = [_np(x[s]) for x in ticks[1]]
        ox1 = oq.solveQPHotstart(a[0], a[1].reshape(-1), a[2], a[3], a[4], a[5], a[6])
        assert oq.result["status"] == st1[s] == 0 and oq.result["iters"] == it1[s]
        assert _maxabs(x1[s] - ox1) < QP_TOL
    qb = wbc_b200.QP(ticks[0][0], ticks[0][1], ticks[0][2], ticks[0][3], n_of_velocity_dimensions=robot.n_velocity_dimensions)
    qb.solveQP()
    with pytest.raises(SystemExit):
        qb.solveQPHotstart(*ticks[1])


# ------------------------------------------------------------------------------------------------ non-default settings
def _spd(rng, scale=1.0):
    M = rng.normal(size=(6, 6))
    return scale * (M @ M.T / 6 + 0.5 * np.eye(6))


@pytest.mark.parametrize("name", ["a1_wx200", "a1_px100_pin_ver"])
@pytest.mark.parametrize("setting", ["static_reach", "spd_weights", "unequal_gains", "all"])
def test_weights_gains_and_static_reach_mode(name, setting):
    """staticReachMode() (Robot_Wrapper4.py:1415-1464: foot task weight 100, joint weight 0.001, GRIP gain 0.05 against
    0.8 for the feet), random SPD 6x6 task weights (the W (J w) / (W J) w products of :476-490 are skipped by the kernel
    when every W is the identity), unequal position gains (the gain list is built FL, FR, RL, RR, GRIP but indexed with
    the FR, FL, RR, RL, GRIP frame index, :125 / :908 -- SURVEY App. D.8): assembly and solution against the oracle."""
    N = 64
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    rng = np.random.default_rng(21)
    if setting in ("static_reach", "all"):
        robot.staticReachMode()
    if setting in ("spd_weights", "all"):
        robot.EE_weight = [_spd(rng) for _ in range(5)]
        robot.trunk_weight = _spd(rng)
        robot.cart_task_weight_EE_list = [1.5, 0.7, 1.2, 0.9, 2.0]
        robot.cart_task_weight_Trunk = 0.6
    if setting in ("unequal_gains", "all"):
        robot.FL_gain, robot.FR_gain = np.identity(6) * 0.3, np.identity(6) * 0.7
        robot.RL_gain, robot.RR_gain = np.identity(6) * 0.45, np.identity(6) * 0.9
        robot.GRIP_gain = np.diag([0.2, 0.4, 0.6, 0.5, 0.5, 0.5]) + 0.05 * np.ones((6, 6))
        robot.EE_gains = [robot.FL_gain, robot.FR_gain, robot.RL_gain, robot.RR_gain, robot.GRIP_gain]   # :125
        robot.trunk_gain = np.diag([0.35, 0.55, 0.75, 0.2, 0.4, 0.6])
    q, targets = _load(robot, N, 20260031, 5e-4)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    ee_t, tr_t = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    asm = robot.assemble(ee_t, tr_t)
    x = _np(robot.step(ee_t, tr_t, advance=False))
    ref = H.oracle_step_batch(name, robot, q, _np(targets), _np(mem0), _np(ref0))
    scaleA = max(1.0, _maxabs(ref["A"]))
    assert _maxabs(_np(asm["A"]) - ref["A"]) < FK_TOL * scaleA
    for k in ("b", "g"):
        assert _maxabs(_np(asm[k]) - ref[k]) < 1e-9 * max(1.0, _maxabs(ref[k])), k
    assert _maxabs(_np(asm["H"]) - ref["H"]) < 1e-10 * max(1.0, _maxabs(ref["H"]))
    st = _np(robot.last_status)
    assert (st == ref["status"]).all() and (st == 0).all()
    assert _maxabs(x - ref["qdot"]) < QP_TOL * max(1.0, _maxabs(ref["qdot"]))
    act = _np(robot.last_active_set).astype(np.uint64)
    nv = robot.n_velocity_dimensions
    same = sum(int(act[s, 0]) == H.act_to_bits(ref["act"][s], nv)[0] and int(act[s, 1]) == H.act_to_bits(ref["act"][s], nv)[1]
               for s in range(N))
    assert same == N and (_np(robot.last_iters) == ref["iters"]).all(), same


@pytest.mark.parametrize("cons", [
    dict(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True),      # 19 rows: the full-width solver layout
    dict(CoM=False, Trunk=True, FR=False, FL=False, RR=False, RL=False, Grip=True),  # 7 rows, no foot rows
    dict(CoM=False, Trunk=False, FR=True, FL=True, RR=True, RL=True, Grip=True),     # 15 rows: feet + gripper, reduced front off
])
def test_gripper_constraint_rows(cons):
    """setConstraints(Grip=True): EEConstraint(4) (Robot_Wrapper4.py:757-761 with i = 4, stacked last by findConstraints
    :816) pins the gripper frame's WORLD linear velocity.  The trunk task keeps working against it."""
    name, N = "a1_wx200", 64
    robot = _robot(name, N, P1_TASKS, cons, True)
    q, targets = _load(robot, N, 20260037, 5e-4)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    ee_t, tr_t = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    asm = robot.assemble(ee_t, tr_t, want=("C", "Clb", "Cub"))
    Cg = robot.EEConstraint(4)[0]
    x = robot.step(ee_t, tr_t, advance=False)
    ref = H.oracle_step_batch(name, robot, q, _np(targets), _np(mem0), _np(ref0))
    assert _maxabs(_np(asm["C"]) - ref["C"]) < FK_TOL
    assert _maxabs(_np(asm["C"][:, -3:]) - _np(Cg)) == 0.0                   # the gripper rows come last
    for k in ("Clb", "Cub"):
        assert _maxabs(_np(asm[k]) - ref[k]) < 1e-9 * max(1.0, _maxabs(ref[k]))
    st = _np(robot.last_status)
    assert (st == ref["status"]).all() and (st == 0).all()
    assert _maxabs(_np(x) - ref["qdot"]) < QP_TOL
    assert torch.einsum("nrk,nk->nr", asm["C"][:, -3:], x).abs().max() < 1e-8     # the gripper does not move
    assert (_np(robot.last_iters) == ref["iters"]).all()
    act = _np(robot.last_active_set).astype(np.uint64)
    for s in range(N):
        wb, wr = H.act_to_bits(ref["act"][s], robot.n_velocity_dimensions)
        assert int(act[s, 0]) == wb and int(act[s, 1]) == wr, s


def test_velocity_damper_without_the_off_by_one():
    """compat_damper_off_by_one = False: joint i's limits are compared with joint i's coordinate instead of the
    reference's q[i] of the un-shifted configuration (Robot_Wrapper4.py:597-613).  States sampled
    right at the limits, so the damper zone is active on many joints and the two indexings differ."""
    name, N = "a1_wx200", 96
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260041, 5e-4)
    t = robot.robot_model
    rng = np.random.default_rng(2)
    lo, up = np.asarray(t.lower[7:t.nq]), np.asarray(t.upper[7:t.nq])
    near = rng.uniform(size=(N, t.nq - 7)) < 0.4
    side = rng.uniform(size=(N, t.nq - 7)) < 0.5
    qn = q.copy()
    edge = np.where(side, lo + rng.uniform(0, 0.03, size=near.shape), up - rng.uniform(0, 0.03, size=near.shape))
    qn[:, 7:] = np.where(near, edge, q[:, 7:])
    qn[:, t.nq - 3:] = q[:, t.nq - 3:]                                   # gripper + fingers stay where the sampler put them
    robot.updateState(torch.as_tensor(qn, device=DEV), feedback=False)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    ee_t, tr_t = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    res = {}
    for flag in (True, False):
        robot.compat_damper_off_by_one = flag
        lb, ub = robot.velDamperJointConstraints()
        x = robot.step(ee_t, tr_t, advance=False).clone()
        ref = H.oracle_step_batch(name, robot, qn, _np(targets), _np(mem0), _np(ref0))
        assert _maxabs(_np(lb) - ref["lb"]) < FK_TOL and _maxabs(_np(ub) - ref["ub"]) < FK_TOL, flag
        ok = (ref["status"] == 0)
        assert (_np(robot.last_status) == ref["status"]).all() and ok.mean() > 0.9
        assert _maxabs(_np(x)[ok] - ref["qdot"][ok]) < QP_TOL
        res[flag] = _np(lb).copy()
    assert _maxabs(res[True] - res[False]) > 1e-3                         # the quirk does change the bounds here


# ------------------------------------------------------------------------------------------------ solver reports
def test_iteration_cap_status_and_iterate():
    """max_qp_iterations (the reference passes nWSR = 100000, QP_Wrapper.py:20; the mirror's default is 200, never
    reached on this path): with a cap below what some states need, those states report WBC_QP_MAXITER and return the
    iterate the method held when it stopped -- the same one the oracle holds with the same cap."""
    name, N = "a1_wx200", 256
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260043, 2e-2)                           # large target noise: many bounds bind
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    ee_t, tr_t = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
    robot.step(ee_t, tr_t, advance=False)
    full_iters = _np(robot.last_iters).copy()
    assert (_np(robot.last_status) == 0).all()
    cap = int(np.median(full_iters))
    assert full_iters.max() > cap >= 15
    robot.max_qp_iterations = cap
    x = _np(robot.step(ee_t, tr_t, advance=False))
    st, it = _np(robot.last_status), _np(robot.last_iters)
    assert ((st == 1) == (full_iters > cap)).all() and (it <= cap).all() and (st[full_iters <= cap] == 0).all()
    ref = H.oracle_step_batch(name, robot, q, _np(targets), _np(mem0), _np(ref0))
    assert (ref["status"] == st).all() and (ref["iters"] == it).all()
    assert _maxabs(x - ref["qdot"]) < QP_TOL
    with pytest.raises(Exception):
        robot.check_status()


def test_not_positive_definite_hessian_is_flagged():
    """Without the joint-posture task H = A^T A of the gripper task alone has rank <= 6: the Cholesky pivot clamp fires
    and the state is flagged WBC_QP_NOT_PD on both sides (the reference hands qpOASES the singular H and ignores what
    comes back, QP_Wrapper.py:45-51).  Through the QP drop-in (run-time-size solver and the nv = 26 register solver)
    and through the fused tick (general and reduced front)."""
    import wbc_b200
    from oracle.qp_wrapper import solve_qp
    rng = np.random.default_rng(6)
    for n in (9, 26):
        N, m = 32, n + 4
        A = rng.normal(size=(N, m, n))
        A[::2, :, n - 2:] = 0.0                 # every other problem: two variables that no row touches -> exact zero pivots
        b = rng.normal(size=(N, m))
        lb, ub = -np.ones((N, n)), np.ones((N, n))
        qp = wbc_b200.QP(A, b, lb, ub, n_of_velocity_dimensions=n)
        x = _np(qp.solveQP())
        st = _np(qp.status)
        assert np.isfinite(x).all(), n
        assert ((st[::2] & 4) != 0).all() and (st[1::2] == 0).all(), (n, st)
        for s in range(N):
            r = solve_qp(A[s].T @ A[s], -A[s].T @ b[s], lb[s], ub[s])
            assert (r["status"] & 4) == (st[s] & 4), (n, s)
            if s % 2:
                assert _maxabs(x[s] - r["x"]) < QP_TOL
    # the fused tick, bounds only: gripper task alone and no joint-posture term -> the leg columns of A are zero
    robot = _robot("a1_wx200", 64, P2_TASKS, NO_CONS, False)                 # Joint=False: no (w/nv)^2 I term
    q, targets = _load(robot, 64, 20260047, 5e-4)
    x = robot.step(targets[:, :15].reshape(64, 5, 3), targets[:, 15:18], advance=False)
    st = _np(robot.last_status)
    assert ((st & 4) != 0).all() and torch.isfinite(x).all()


# ------------------------------------------------------------------------------------------------ host-buffer runWBC tick
@pytest.mark.parametrize("chunks", [0, 1, 5])
def test_step_host_closed_loop_equals_rollout(chunks):
    """wbc_step_host in closed-loop form IS the runWBC tick for a caller that holds NumPy arrays: per tick the IMU
    quaternion and the targets come from (pinned) host memory, the joint targets and the solver report go back, and
    the configuration / task memory advance in place on the device.  K calls must land exactly where `rollout` lands
    (zero-copy launch, one staged slice, more slices than streams)."""
    name, N, K = "a1_wx200", 2500, 5
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260053, 5e-4)
    q0, mem0, ref0 = robot.current_joint_config.clone(), robot._mem.clone(), robot._ref.clone()
    gen = torch.Generator(device=DEV); gen.manual_seed(3)
    drift = torch.randn(K, N, 18, dtype=torch.float64, device=DEV, generator=gen).mul_(2e-4).cumsum(0)
    traj = targets[None] + drift
    imu = q0[:, 3:7][None].repeat(K, 1, 1) + 1e-3 * torch.randn(K, N, 4, dtype=torch.float64, device=DEV, generator=gen)
    imu = imu / imu.norm(dim=2, keepdim=True)
    qh, vh, sh = robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18], imu_quat_traj=imu, record=True)
    mem_end = robot._mem.clone()
    # the same K ticks through host buffers
    robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0); robot._ref.copy_(ref0)
    nq, nv = robot.n_configuration_dimensions, robot.n_velocity_dimensions
    out = {"joint_targets": torch.empty(N, nq - 7, dtype=torch.float64).pin_memory(),
           "qdot": torch.empty(N, nv, dtype=torch.float64).pin_memory(),
           "status": torch.empty(N, dtype=torch.int32).pin_memory(), "iters": torch.empty(N, dtype=torch.int32).pin_memory()}
    for k in range(K):
        host_in = {"targets": traj[k].cpu().pin_memory(), "imu": imu[k].cpu().pin_memory()}
        for t in out.values():
            t.fill_(-7)
        h2d, d2h = robot.step_host(host_in, out, chunks=chunks, closed_loop=True)
        torch.cuda.synchronize()
        assert h2d == N * 8 * 22 and d2h == N * (8 * (nq - 7) + 8 * nv + 8)
        assert torch.equal(robot.current_joint_config, qh[k]), k
        assert torch.equal(out["joint_targets"], qh[k][:, 7:].cpu()) and torch.equal(out["qdot"], vh[k].cpu())
        assert torch.equal(out["status"], sh[k].cpu())
    assert torch.equal(robot._mem, mem_end)


def test_step_host_self_tuning_path_does_not_change_results():
    """`chunks=0` lets the model handle measure the zero-copy launch against 8 staged slices on its first four calls and keep
    the faster one (one GPU alone: zero-copy; eight GPUs on one host NUMA node: staged).  Whatever it picks and while it is
    still trying both, every tick lands where the device rollout lands; afterwards `host_path()` names the decision."""
    name, N, K = "a1_wx200", 8192, 8
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260067, 5e-4)
    q0, mem0 = robot.current_joint_config.clone(), robot._mem.clone()
    gen = torch.Generator(device=DEV); gen.manual_seed(11)
    traj = targets[None] + torch.randn(K, N, 18, dtype=torch.float64, device=DEV, generator=gen).mul_(2e-4).cumsum(0)
    imu = q0[:, 3:7][None].repeat(K, 1, 1)
    qh, vh, sh = robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18], imu_quat_traj=imu, record=True)
    robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0)
    nq = robot.n_configuration_dimensions
    out = {"joint_targets": torch.empty(N, nq - 7, dtype=torch.float64).pin_memory(), "status": torch.empty(N, dtype=torch.int32).pin_memory()}
    assert robot.host_path() == "undecided"
    for k in range(K):
        host_in = {"targets": traj[k].cpu().pin_memory(), "imu": imu[k].cpu().pin_memory()}
        robot.step_host(host_in, out, chunks=0, closed_loop=True)
        torch.cuda.synchronize()
        assert torch.equal(robot.current_joint_config, qh[k]) and torch.equal(out["joint_targets"], qh[k][:, 7:].cpu()), k
        assert torch.equal(out["status"], sh[k].cpu())
    assert robot.host_path() in ("zero_copy", "staged_8")
    # a forced path leaves the decision alone
    robot.step_host(host_in, out, chunks=-1, closed_loop=True)
    assert robot.host_path() in ("zero_copy", "staged_8")


@pytest.mark.parametrize("name,N,K", [("a1_wx200", 1, 4), ("a1_wx200", 37, 6), ("a1_wx200", 2368, 3), ("a1_wx200", 2369, 3),
                                      ("a1_px100_pin_ver", 5000, 4), ("a1_wx200", 6000, 1)])
def test_rollout_in_one_launch_equals_tick_by_tick(name, N, K):
    """`wbc_rollout` runs the whole closed-loop horizon as ONE persistent launch when the reduced-front instantiation applies
    (a robot stays with one warp for all K ticks).  It must land bit for bit where K separate fused launches land
    (`rollout(record=True)` steps tick by tick): batches smaller than one round (every prefetch of the next tick has to wait
    for the tail: the `late` path), exactly one round of 148 x 16 warps, one state more, several rounds with a ragged last
    one, and K = 1."""
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260059, 5e-4)
    q0, mem0, ref0 = robot.current_joint_config.clone(), robot._mem.clone(), robot._ref.clone()
    gen = torch.Generator(device=DEV); gen.manual_seed(5)
    traj = targets[None] + torch.randn(K, N, 18, dtype=torch.float64, device=DEV, generator=gen).mul_(2e-4).cumsum(0)
    imu = q0[:, 3:7][None].repeat(K, 1, 1) + 1e-3 * torch.randn(K, N, 4, dtype=torch.float64, device=DEV, generator=gen)
    imu = imu / imu.norm(dim=2, keepdim=True)
    ee, tr = traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18]
    for with_imu in (True, False):
        robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0); robot._ref.copy_(ref0)
        qh, vh, sh = robot.rollout(ee, tr, imu_quat_traj=imu if with_imu else None, record=True)
        mem_end, it_end, act_end = robot._mem.clone(), robot.last_iters.clone(), robot.last_active_set.clone()
        robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0); robot._ref.copy_(ref0)
        v = robot.rollout(ee, tr, imu_quat_traj=imu if with_imu else None)
        assert torch.equal(robot.current_joint_config, qh[-1]) and torch.equal(v, vh[-1])
        assert torch.equal(robot.last_status, sh[-1]) and torch.equal(robot._mem, mem_end)
        assert torch.equal(robot.last_iters, it_end) and torch.equal(robot.last_active_set, act_end)
        assert torch.equal(robot._ref, ref0)


@pytest.mark.parametrize("joint", ["HYBRID", "MANI"])
def test_rollout_in_one_launch_finite_difference_joint_task(joint):
    """The closed loop `sim3.py` actually runs (gripper task + "HYBRID" joint task + trunk / feet constraints, sim3.py:145-148
    inside the tick loop :287-327) as one persistent launch: bit for bit where the tick-by-tick launches land -- including the
    reference's perturbed-state quirk of the finite-difference modes (the configuration the tick integrates from is the
    perturbed one), which the next tick of the same launch has to see."""
    name, N, K = "a1_px100_pin_ver", 700, 5
    robot = _robot(name, N, P2_TASKS, P2_CONS, joint)
    q, targets = _load(robot, N, 20260073, 5e-4)
    q0, mem0 = robot.current_joint_config.clone(), robot._mem.clone()
    gen = torch.Generator(device=DEV); gen.manual_seed(13)
    traj = targets[None] + torch.randn(K, N, 18, dtype=torch.float64, device=DEV, generator=gen).mul_(2e-4).cumsum(0)
    ee, tr = traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18]
    qh, vh, sh = robot.rollout(ee, tr, record=True)
    mem_end, it_end = robot._mem.clone(), robot.last_iters.clone()
    robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0)
    v = robot.rollout(ee, tr)
    assert torch.equal(robot.current_joint_config, qh[-1]) and torch.equal(v, vh[-1])
    assert torch.equal(robot.last_status, sh[-1]) and torch.equal(robot._mem, mem_end) and torch.equal(robot.last_iters, it_end)


def test_fp32_host_io_mode_agrees_with_fp64():
    """The optional FP32 I/O mode (north_star: "an optional FP32 mode must agree within 1e-4"): float32 arrays on the
    host side, float64 arithmetic in the tick.  Closed loop over several ticks against the float64 call on the same data:
      * increment inputs (WBC_HOST_FLAG_DELTA_INPUTS, `HostDeltaEncoder`): qdot AND joint targets within 1e-4;
      * absolute float32 targets: the joint position targets runWBC returns within 1e-4 (measured ~1e-6); qdot carries
        the float32 rounding of a ~0.5 m position (3e-8) times 1 / dt = 500 times the leg Jacobian's inverse: 1e-5
        typical, a few 1e-3 in the worst state, so only a loose bound is asserted for it;
    and the open-loop call with every array travelling in float32."""
    import wbc_b200
    name, N, K = "a1_wx200", 3000, 4
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260059, 5e-4)
    q0, mem0, ref0 = robot.current_joint_config.clone(), robot._mem.clone(), robot._ref.clone()
    nq, nv = robot.n_configuration_dimensions, robot.n_velocity_dimensions
    gen = torch.Generator(device=DEV); gen.manual_seed(5)
    drift = torch.randn(K, N, 18, dtype=torch.float64, device=DEV, generator=gen).mul_(2e-4).cumsum(0)
    traj = (targets[None] + drift).cpu()
    imu = q0[:, 3:7][None].repeat(K, 1, 1) + 1e-3 * torch.randn(K, N, 4, dtype=torch.float64, device=DEV, generator=gen)
    imu = (imu / imu.norm(dim=2, keepdim=True)).cpu()

    def run(dtype, chunks, delta=False):
        robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0); robot._ref.copy_(ref0)
        out = {"joint_targets": torch.empty(N, nq - 7, dtype=dtype).pin_memory(), "qdot": torch.empty(N, nv, dtype=dtype).pin_memory(),
               "status": torch.empty(N, dtype=torch.int32).pin_memory(), "iters": torch.empty(N, dtype=torch.int32).pin_memory()}
        enc = wbc_b200.HostDeltaEncoder(robot) if delta else None
        res = []
        for k in range(K):
            if delta:
                t_in, i_in = enc.encode(traj[k], imu[k])
                t_in, i_in = t_in.pin_memory(), i_in.pin_memory()
            else:
                t_in, i_in = traj[k].to(dtype).pin_memory(), imu[k].to(dtype).pin_memory()
            h2d, d2h = robot.step_host({"targets": t_in, "imu": i_in}, out, chunks=chunks, closed_loop=True, delta_inputs=delta)
            torch.cuda.synchronize()
            res.append((out["joint_targets"].double().clone(), out["qdot"].double().clone(), out["status"].clone()))
        return res, h2d, d2h

    r64, h64, d64 = run(torch.float64, 0)
    for chunks in (0, 4):
        for delta in (True, False):
            r32, h32, d32 = run(torch.float32, chunks, delta)
            assert h32 * 2 == h64 and (d32 - 8 * N) * 2 == d64 - 8 * N
            for k in range(K):
                ok = (r64[k][2] == 0) & (r32[k][2] == 0)              # (a jittering IMU can push a few states into an
                assert ok.double().mean() > 0.98                      #  infeasible trunk box: compared where both solved)
                assert (r32[k][2] == r64[k][2]).double().mean() > 0.999
                dv = (r32[k][1] - r64[k][1])[ok].abs()
                dj = (r32[k][0] - r64[k][0])[ok].abs().max()
                assert dj < (1e-6 if delta else 1e-4), (k, chunks, delta, float(dj))
                if delta:
                    assert dv.max() < 1e-4, (k, chunks, float(dv.max()))
                else:
                    assert dv.max() < 2e-2 and dv.median() < 1e-4, (k, chunks, float(dv.max()), float(dv.median()))
    # open loop, everything travelling in float32 (q, targets, task memory, references)
    robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0); robot._ref.copy_(ref0)
    x64 = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False).clone()
    hin = {"q": q0.cpu().float().pin_memory(), "targets": targets.cpu().float().pin_memory(),
           "mem": mem0.cpu().float().pin_memory(), "ref": ref0.cpu().float().pin_memory()}
    hout = {"qdot": torch.empty(N, nv, dtype=torch.float32).pin_memory(), "status": torch.empty(N, dtype=torch.int32).pin_memory(),
            "iters": torch.empty(N, dtype=torch.int32).pin_memory()}
    for chunks in (0, 3):
        hout["qdot"].fill_(float("nan"))
        robot.step_host(hin, hout, chunks=chunks)
        torch.cuda.synchronize()
        ok = (hout["status"] == 0) & (robot.last_status.cpu() == 0)
        assert ok.double().mean() > 0.99
        dv = (hout["qdot"].double()[ok] - x64.cpu()[ok]).abs()
        assert dv.max() < 5e-2 and dv.median() < 1e-3, (chunks, float(dv.max()), float(dv.median()))


# ------------------------------------------------------------------------------------------------ the boundary from plain C
def test_c_abi_demo_matches_python_path(tmp_path):
    """examples/c_abi_demo.c -- a C11 program with no Python and no torch in it -- runs one batched runWBC tick through
    wbc_model_create + wbc_step on the arrays of a case file and must return, bit for bit, what the Python mirror returns for
    the same states: the drop-in boundary is the C ABI, not the Python on top of it."""
    import ctypes as C
    import struct
    import subprocess
    from tests.test_cabi_cpu import build_c_demo
    name, N = "a1_wx200", 3000
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260071, 5e-4)
    mem0, ref0 = robot._mem.clone(), robot._ref.clone()
    case, result = tmp_path / "case.bin", tmp_path / "result.bin"
    table, cfg = robot._table, robot._config()
    with open(case, "wb") as f:
        f.write(struct.pack("<qqd", 0x57424332, N, float(robot.dt)))
        f.write(struct.pack("<q", C.sizeof(table))); f.write(bytes(table))
        f.write(struct.pack("<q", C.sizeof(cfg))); f.write(bytes(cfg))
        for a in (robot.current_joint_config, targets, mem0, ref0):
            f.write(a.detach().cpu().contiguous().numpy().tobytes())
    exe = build_c_demo(tmp_path)
    out = subprocess.run([exe, str(case), str(result)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    print(out.stdout.strip())
    nq, nv = robot.n_configuration_dimensions, robot.n_velocity_dimensions
    raw = np.fromfile(result, dtype=np.uint8)
    o = 0
    qdot_c = raw[o:o + N * nv * 8].view(np.float64).reshape(N, nv); o += N * nv * 8
    qn_c = raw[o:o + N * nq * 8].view(np.float64).reshape(N, nq); o += N * nq * 8
    st_c = raw[o:o + N * 4].view(np.int32); o += N * 4
    it_c = raw[o:o + N * 4].view(np.int32)
    x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=True)
    assert (st_c == 0).all()
    assert np.array_equal(qdot_c, _np(x)) and np.array_equal(qn_c, _np(robot.current_joint_config))
    assert np.array_equal(st_c, _np(robot.last_status)) and np.array_equal(it_c, _np(robot.last_iters))
