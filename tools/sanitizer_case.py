"""Smallest case that walks every kernel / code path of libwbc_b200.so once, for compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck python tools/sanitizer_case.py`, one tool per gpurun call):
fused tick (reduced front, general front, full-width solver, FD joint task, extension rows, in-kernel fallback),
closed-loop tail in place, zero-copy host buffers (float64 and float32 increments), the staged host pipeline,
accessor kernels, the batched QP drop-in.  37 states: more than one CTA, padding warps that shadow the last state."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import wbc_b200
from wbc_b200 import synthetic

P2C = dict(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
dev = "cuda:0"
N = 37
for name, joint, cons, extra in (("a1_wx200", True, P2C, False),
                                 ("a1_px100_pin_ver", "HYBRID", dict(P2C, CoM=True), False),
                                 ("a1_wx200", True, dict(P2C, FR=False, FL=False, RR=False, RL=False), True),
                                 ("a1_wx200", "PREV", dict(P2C, Grip=True), False)):
    r = wbc_b200.RobotModel(name, batch=N, device=dev)
    r.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=joint)
    r.setConstraints(**cons)
    if extra:
        r.extra_rows = synthetic.config3_rows(r.robot_model)
    q = synthetic.sample_configurations(r.robot_model, N, 3)
    t = synthetic.load_batch(r, q, synthetic.sample_noise(N, 3, 5e-3))
    ee, tr = t[:, :15].reshape(N, 5, 3), t[:, 15:18]
    r.assemble(ee, tr)
    for k in range(3):
        r.step(ee, tr, imu_quat=r.current_joint_config[:, 3:7].clone(), advance=True)
    r.rollout(torch.stack([ee, ee]), torch.stack([tr, tr]))
    r.runWBC(r.current_joint_config[:, 3:7].clone(), ee, tr)
    r.trunkWorldPos(); r.findConstraints(); r.velDamperJointConstraints(); r.qpA(); r.qpb(ee, tr)
    r.jointVelocitiestoConfig(torch.zeros(N, r.n_velocity_dimensions, dtype=torch.float64, device=dev), update_model=True)
    r.frameJacobians(2); r.getFrameJacobian(r.trunk_frame_index, 1)
    nq, nv = r.n_configuration_dimensions, r.n_velocity_dimensions
    # host buffers: closed loop, zero-copy and staged, float64 and float32 increments; open loop with everything travelling
    for dtype, delta in ((torch.float64, False), (torch.float32, True)):
        out = {"joint_targets": torch.empty(N, nq - 7, dtype=dtype).pin_memory(), "qdot": torch.empty(N, nv, dtype=dtype).pin_memory(),
               "status": torch.empty(N, dtype=torch.int32).pin_memory(), "iters": torch.empty(N, dtype=torch.int32).pin_memory()}
        enc = wbc_b200.HostDeltaEncoder(r) if delta else None
        for chunks in (0, 3):
            if delta:
                a, b = enc.encode(t.cpu(), r.current_joint_config[:, 3:7].cpu())
                hin = {"targets": a.pin_memory(), "imu": b.pin_memory()}
            else:
                hin = {"targets": t.cpu().pin_memory(), "imu": r.current_joint_config[:, 3:7].cpu().pin_memory()}
            r.step_host(hin, out, chunks=chunks, closed_loop=True, delta_inputs=delta)
            torch.cuda.synchronize()
    hin = {"q": r.current_joint_config.cpu().pin_memory(), "targets": t.cpu().pin_memory(), "mem": r._mem.cpu().pin_memory(),
           "ref": r._ref.cpu().pin_memory()}
    out = {"qdot": torch.empty(N, nv, dtype=torch.float64).pin_memory(), "status": torch.empty(N, dtype=torch.int32).pin_memory(),
           "iters": torch.empty(N, dtype=torch.int32).pin_memory()}
    for chunks in (0, 2):
        r.step_host(hin, out, chunks=chunks)
        torch.cuda.synchronize()
    print(name, joint, "ok", int((r.last_status == 0).sum()), "/", N)
A = torch.randn(5, 40, 26, dtype=torch.float64, device=dev); b = torch.randn(5, 40, dtype=torch.float64, device=dev)
lb = -torch.ones(5, 26, dtype=torch.float64, device=dev); ub = -lb
Cm = torch.randn(5, 20, 26, dtype=torch.float64, device=dev); cl = -torch.ones(5, 20, dtype=torch.float64, device=dev)
x = wbc_b200.QP(A, b, lb, ub, Cm.transpose(1, 2), cl, -cl, n_of_velocity_dimensions=26).solveQP()
x = wbc_b200.QP(A, b, lb, ub, Cm[:, :9].transpose(1, 2), cl[:, :9], -cl[:, :9], n_of_velocity_dimensions=26).solveQP()
x = wbc_b200.QP(A[:, :, :7].contiguous(), b, lb[:, :7].contiguous(), ub[:, :7].contiguous(), n_of_velocity_dimensions=7).solveQP()
torch.cuda.synchronize()
print("sanitizer case done")
