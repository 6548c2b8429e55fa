"""-m gpu: a memcheck / racecheck of our own.  compute-sanitizer is CLOSED on this pool ("runs under it have left GPUs
needing a reset", profiles/r2_compute_sanitizer_closed.log), so out-of-bounds accesses and races are looked for with
the means the pool's notice suggests: guard zones, small cases and comparison with reference runs.

  * every array handed to the C ABI sits between two guard zones inside one big allocation: NaN guards next to float
    arrays, a bit pattern next to integer ones.  After the call the guards must be untouched (no out-of-bounds WRITE),
    and the results must be bit-identical to the same call on ordinary allocations (an out-of-bounds READ would pull NaN
    or a neighbour's data into the arithmetic);
  * the same launch repeated gives bit-identical results (a shared-memory race between the lanes of a warp -- the only
    kind possible here: every warp owns its workspace, the model block is read-only after its staging barrier -- would
    show up as run-to-run differences), across batch sizes that leave padding warps, for every kernel of the library.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.test_gpu_parity import _robot, _load, P1_TASKS, P2_TASKS, P2_CONS, NO_CONS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GUARD = 96          # elements either side


class Arena:
    """One allocation; `take` hands out views separated by guard zones and remembers them."""

    def __init__(self, nbytes=64 << 20):
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
        self.off = 0
        self.guards = []

    def take(self, shape, dtype, fill=None):
        esz = torch.empty(0, dtype=dtype).element_size()
        n = int(np.prod(shape)) if len(shape) else 1
        self.off = (self.off + 15) & ~15                                    # the library's 16-byte alignment contract for bulk copies
        total = (n + 2 * GUARD) * esz
        raw = self.buf[self.off:self.off + total].view(dtype)
        # make the payload itself 16-byte aligned as torch allocations are: guards are GUARD elements = multiples of 16 bytes
        assert (GUARD * esz) % 16 == 0
        self.off += total
        if dtype.is_floating_point:
            raw.fill_(float("nan"))
        else:
            raw.fill_(0x5A5A5A5A if esz == 4 else 0x5A5A5A5A5A5A5A5A)
        view = raw[GUARD:GUARD + n].view(*shape)
        if fill is not None:
            view.copy_(fill)
        self.guards.append((raw, n, dtype))
        return view

    def check(self):
        for raw, n, dtype in self.guards:
            for g in (raw[:GUARD], raw[GUARD + n:]):
                if dtype.is_floating_point:
                    assert torch.isnan(g).all(), "guard zone overwritten"
                else:
                    pat = 0x5A5A5A5A if g.element_size() == 4 else 0x5A5A5A5A5A5A5A5A
                    assert (g == pat).all(), "guard zone overwritten"


def _step_args(robot, arena, targets, imu, closed):
    """All arrays of one wbc_step call inside the arena."""
    from wbc_b200 import _cabi as cabi
    N, nq, nv = robot.N, robot.n_configuration_dimensions, robot.n_velocity_dimensions
    f64, i32, i64 = torch.float64, torch.int32, torch.int64
    a = {"q": arena.take((N, nq), f64, robot.current_joint_config), "targets": arena.take((N, 18), f64, targets),
         "mem": arena.take((N, 72), f64, robot._mem), "ref": arena.take((N, 24), f64, robot._ref),
         "qdot": arena.take((N, nv), f64), "status": arena.take((N,), i32), "iters": arena.take((N,), i32),
         "act": arena.take((N, 2), i64)}
    io = cabi.WbcStepIO()
    io.q, io.targets, io.mem_in, io.ref = (a[k].data_ptr() for k in ("q", "targets", "mem", "ref"))
    io.dt = float(robot.dt)
    io.qdot, io.status, io.iters, io.active_set = (a[k].data_ptr() for k in ("qdot", "status", "iters", "act"))
    if closed:
        a["imu"] = arena.take((N, 4), f64, imu)
        a["mem_out"] = arena.take((N, 72), f64)
        a["q_next"] = arena.take((N, nq), f64)
        a["joints"] = arena.take((N, nq - 7), f64)
        io.imu_quat, io.mem_out, io.q_next, io.joint_targets = (a[k].data_ptr() for k in ("imu", "mem_out", "q_next", "joints"))
    return a, io


CASES = [
    ("a1_wx200", P1_TASKS, P2_CONS, True, None),                       # reduced front (the bench instantiation)
    ("a1_px100_pin_ver", P1_TASKS, NO_CONS, True, None),               # bounds only, general front
    ("a1_px100_pin_ver", P2_TASKS, P2_CONS, "HYBRID", None),           # finite-difference joint task
    ("a1_wx200", P1_TASKS, dict(P2_CONS, CoM=True, Grip=True), "PREV", None),     # 21 rows: full-width solver, CoM rows
    ("a1_wx200", P1_TASKS, dict(P2_CONS, FR=False, FL=False, RR=False, RL=False), True, "config3"),   # extension rows
]


@pytest.mark.parametrize("name,tasks,cons,joint,extra", CASES)
@pytest.mark.parametrize("N", [1, 37, 148 * 16 + 5])
def test_fused_tick_inside_guard_zones(name, tasks, cons, joint, extra, N):
    from wbc_b200 import synthetic, _cabi as cabi
    if joint == "HYBRID" and N > 100:
        N = 148 * 12 + 5
    robot = _robot(name, N, tasks, cons, joint)
    if extra:
        robot.extra_rows = synthetic.config3_rows(robot.robot_model)
    q, targets = _load(robot, N, 20260071, 5e-3)
    imu = robot.current_joint_config[:, 3:7].clone()
    lib, cfg = cabi.load(), robot._config()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for closed in (False, True):
        # reference run on ordinary allocations, through the same entry point
        plain = {k: torch.zeros_like(v) for k, v in
                 _step_args(robot, Arena(24 << 20), targets, imu, closed)[0].items() if k in ("qdot", "status", "iters", "act", "mem_out", "q_next", "joints")}
        io0 = robot._io(targets=targets.contiguous(), qdot=plain["qdot"], status=plain["status"], iters=plain["iters"], active_set=plain["act"])
        if closed:
            io0.imu_quat, io0.mem_out, io0.q_next, io0.joint_targets = imu.data_ptr(), plain["mem_out"].data_ptr(), plain["q_next"].data_ptr(), plain["joints"].data_ptr()
        cabi.check(lib.wbc_step(robot._model, C.byref(cfg), C.byref(io0), N, stream))
        torch.cuda.synchronize()
        outs = ("qdot", "status", "iters", "act") + (("mem_out", "q_next", "joints") if closed else ())
        for rep in range(3):                                           # guarded, three times: bit-identical each time
            arena = Arena(8 << 20 if N < 100 else 48 << 20)
            a, io = _step_args(robot, arena, targets, imu, closed)
            cabi.check(lib.wbc_step(robot._model, C.byref(cfg), C.byref(io), N, stream))
            torch.cuda.synchronize()
            arena.check()
            for k in ("q", "targets", "mem", "ref"):                   # inputs are read-only (out of place here)
                src = {"q": robot.current_joint_config, "targets": targets, "mem": robot._mem, "ref": robot._ref}[k]
                assert torch.equal(a[k], src), k
            for k in outs:
                assert torch.equal(a[k], plain[k]), (k, closed, rep)


def test_accessor_kernels_and_qp_inside_guard_zones():
    """wbc_fk_jac, wbc_joint_jacobians, wbc_init_memory, wbc_integrate, wbc_base_estimate, wbc_assemble, wbc_qp_solve."""
    from wbc_b200 import _cabi as cabi
    name, N = "a1_wx200", 301
    robot = _robot(name, N, P1_TASKS, dict(P2_CONS, CoM=True), True)
    q, targets = _load(robot, N, 20260073, 5e-3)
    lib, stream = cabi.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nq, nv, nj = robot.n_configuration_dimensions, robot.n_velocity_dimensions, robot.robot_model.njoints
    f64, i32, i64 = torch.float64, torch.int32, torch.int64
    p = lambda t: C.c_void_p(t.data_ptr())
    ar = Arena(96 << 20)
    qg = ar.take((N, nq), f64, robot.current_joint_config)
    tg = ar.take((N, 18), f64, targets)
    # FK / Jacobians
    sel = (C.c_int32 * 6)(*range(6))
    oMf, J = ar.take((N, 6, 12), f64), ar.take((N, 6, 6, nv), f64)
    oMi, Jw = ar.take((N, nj, 12), f64), ar.take((N, 6, nv), f64)
    for rf in (0, 1, 2):
        cabi.check(lib.wbc_fk_jac(robot._model, p(qg), N, sel, 6, rf, p(oMf), p(J), stream))
        ref_oMf, ref_J = robot.frameJacobians(rf)
        torch.cuda.synchronize()
        assert torch.equal(oMf, ref_oMf) and torch.equal(J, ref_J)
    cabi.check(lib.wbc_joint_jacobians(robot._model, p(qg), N, p(oMi), p(Jw), stream))
    torch.cuda.synchronize()
    assert torch.equal(Jw, robot.J) and torch.equal(oMi, robot._oMi)
    # initialiseWBC snapshot, integrate, base estimate
    mem, ref = ar.take((N, 72), f64), ar.take((N, 24), f64)
    cabi.check(lib.wbc_init_memory(robot._model, p(qg), N, p(mem), p(ref), stream))
    m2, r2 = robot._log_previous_states()
    v = ar.take((N, nv), f64, torch.randn(N, nv, dtype=f64, device=DEV))
    qi = ar.take((N, nq), f64)
    cabi.check(lib.wbc_integrate(robot._model, p(qg), p(v), N, 0.002, p(qi), stream))
    qb, base = ar.take((N, nq), f64), ar.take((N, 3), f64)
    imu = ar.take((N, 4), f64, robot.current_joint_config[:, 3:7])
    cabi.check(lib.wbc_base_estimate(robot._model, p(qi), p(imu), p(tg), N, p(qb), p(base), stream))
    torch.cuda.synchronize()
    assert torch.equal(mem, m2) and torch.equal(ref, r2)
    assert torch.equal(qb[:, :3], base) and torch.equal(qb[:, 7:], qi[:, 7:]) and torch.isfinite(qb).all()
    # assembly accessor: every output requested
    cfg = robot._config()
    m, nc = robot._rows(cfg)
    out = cabi.WbcAssembleOut()
    got = {}
    for k, shp in (("A", (N, m, nv)), ("b", (N, m)), ("lb", (N, nv)), ("ub", (N, nv)), ("C", (N, nc, nv)), ("Clb", (N, nc)),
                   ("Cub", (N, nc)), ("H", (N, nv, nv)), ("g", (N, nv))):
        got[k] = ar.take(shp, f64)
        setattr(out, k, got[k].data_ptr())
    mem_a, ref_a = ar.take((N, 72), f64, robot._mem), ar.take((N, 24), f64, robot._ref)
    io = cabi.WbcStepIO()
    io.q, io.targets, io.mem_in, io.ref, io.dt = qg.data_ptr(), tg.data_ptr(), mem_a.data_ptr(), ref_a.data_ptr(), float(robot.dt)
    cabi.check(lib.wbc_assemble(robot._model, C.byref(cfg), C.byref(io), N, C.byref(out), stream))
    want = robot.assemble(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18])
    torch.cuda.synchronize()
    for k in got:
        assert torch.equal(got[k], want[k]), k
    # the batched QP drop-in on these matrices: register solver (nv = 26), then a run-time-size problem
    x, st, it, act = ar.take((N, nv), f64), ar.take((N,), i32), ar.take((N,), i32), ar.take((N, 2), i64)
    for rep in range(2):
        cabi.check(lib.wbc_qp_solve(N, nv, m, nc, p(got["A"]), p(got["b"]), None, None, p(got["lb"]), p(got["ub"]), p(got["C"]),
                                    p(got["Clb"]), p(got["Cub"]), 200, p(x), p(st), p(it), p(act), stream))
        torch.cuda.synchronize()
        if rep == 0:
            x0, it0 = x.clone(), it.clone()
    assert torch.equal(x, x0) and torch.equal(it, it0)
    n2, m2_, c2 = 11, 17, 5
    A2, b2 = ar.take((N, m2_, n2), f64, torch.randn(N, m2_, n2, dtype=f64, device=DEV)), ar.take((N, m2_), f64, torch.randn(N, m2_, dtype=f64, device=DEV))
    lb2, ub2 = ar.take((N, n2), f64, -torch.ones(N, n2, dtype=f64, device=DEV)), ar.take((N, n2), f64, torch.ones(N, n2, dtype=f64, device=DEV))
    C2 = ar.take((N, c2, n2), f64, torch.randn(N, c2, n2, dtype=f64, device=DEV))
    cl2, cu2 = ar.take((N, c2), f64, -torch.ones(N, c2, dtype=f64, device=DEV)), ar.take((N, c2), f64, torch.ones(N, c2, dtype=f64, device=DEV))
    x2, st2, it2 = ar.take((N, n2), f64), ar.take((N,), i32), ar.take((N,), i32)
    cabi.check(lib.wbc_qp_solve(N, n2, m2_, c2, p(A2), p(b2), None, None, p(lb2), p(ub2), p(C2), p(cl2), p(cu2), 200, p(x2), p(st2),
                                p(it2), None, stream))
    torch.cuda.synchronize()
    assert torch.isfinite(x2).all() and (st2 == 0).all()
    ar.check()


def test_rollout_in_place_inside_guard_zones_and_deterministic():
    """wbc_rollout advances q and the task memory IN PLACE while padding warps shadow the last state: guards intact, and
    two runs from the same start end on the same bits."""
    from wbc_b200 import _cabi as cabi
    name, N, K = "a1_wx200", 148 * 16 + 3, 6
    robot = _robot(name, N, P1_TASKS, P2_CONS, True)
    q, targets = _load(robot, N, 20260079, 5e-4)
    lib, stream = cabi.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nq, nv = robot.n_configuration_dimensions, robot.n_velocity_dimensions
    f64 = torch.float64
    ends = []
    for rep in range(2):
        ar = Arena(64 << 20)
        a, io = _step_args(robot, ar, targets, None, False)
        traj = ar.take((K, N, 18), f64, targets[None].repeat(K, 1, 1) + 1e-4 * torch.arange(K, device=DEV, dtype=f64)[:, None, None])
        imu = ar.take((K, N, 4), f64, robot.current_joint_config[:, 3:7][None].repeat(K, 1, 1))
        cabi.check(lib.wbc_rollout(robot._model, C.byref(robot._config()), C.byref(io), C.c_void_p(traj.data_ptr()),
                                   C.c_void_p(imu.data_ptr()), K, N, stream))
        torch.cuda.synchronize()
        ar.check()
        assert torch.isfinite(a["q"]).all() and torch.isfinite(a["mem"]).all()
        ends.append((a["q"].clone(), a["mem"].clone(), a["qdot"].clone(), a["iters"].clone()))
    for x, y in zip(*ends):
        assert torch.equal(x, y)
