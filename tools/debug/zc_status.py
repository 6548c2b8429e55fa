"""Zero-copy closed-loop host tick: what do the 4-byte status stores to host memory cost?  (131072 robots, standing sampler)"""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
import bench
from wbc_b200 import _cabi as cabi
class A: pass
args = A(); args.robot = "a1_wx200"; args.dt = 0.002; args.seed = 20260004; args.sigma = 5e-4
N = 131072
ctx = bench.Ctx()
robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma, standing=True)
q0, mem0 = robot.current_joint_config.clone(), robot._mem.clone()
nq = robot.n_configuration_dimensions
gen = torch.Generator(device="cpu"); gen.manual_seed(5)
walk = torch.randn(8, N, 18, dtype=torch.float64, generator=gen).mul_(1e-4).cumsum(0); walk[:, :, :12] = 0
ring = [{"targets": (targets.cpu() + walk[r]).pin_memory(), "imu": q0[:, 3:7].cpu().pin_memory()} for r in range(8)]
out = {"joint_targets": torch.empty(N, nq - 7, dtype=torch.float64).pin_memory(), "status": torch.empty(N, dtype=torch.int32).pin_memory()}
cfg = robot._config()
stage_imu = torch.empty(N, 4, dtype=torch.float64, device=ctx.dev); stage_j = torch.empty(N, nq - 7, dtype=torch.float64, device=ctx.dev)
def tick(k, with_status, with_joints=True):
    io = cabi.WbcStepIO()
    io.q = robot.current_joint_config.data_ptr(); io.targets = robot._targets.data_ptr(); io.mem_in = robot._mem.data_ptr()
    io.ref = robot._ref.data_ptr(); io.dt = args.dt; io.qdot = robot.qdot.data_ptr(); io.status = robot.last_status.data_ptr()
    io.iters = robot.last_iters.data_ptr(); io.q_next = io.q; io.mem_out = io.mem_in
    io.joint_targets = stage_j.data_ptr(); io.imu_quat = stage_imu.data_ptr()
    host = cabi.WbcHostIO(); host.dtype = cabi.HOST_F64; host.flags = 0
    host.targets = ring[k % 8]["targets"].data_ptr(); host.imu_quat = ring[k % 8]["imu"].data_ptr()
    if with_joints: host.joint_targets = out["joint_targets"].data_ptr()
    if with_status: host.status = out["status"].data_ptr()
    cabi.check(robot._lib.wbc_step_host(robot._model, C.byref(cfg), C.byref(io), C.byref(host), N, -1, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for label, ws, wj in (("joints + status", True, True), ("joints only", False, True), ("neither (inputs from host only)", False, False), ("joints + status", True, True)):
    robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0)
    for k in range(8): tick(k, ws, wj)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 300
    a.record()
    for k in range(K): tick(8 + k, ws, wj)
    b.record(); torch.cuda.synchronize()
    print(f"{label:34s} {N * K / (a.elapsed_time(b) * 1e-3) / 1e6:7.2f} M ticks/s")
# device-resident closed loop for comparison
robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0)
K = 40
traj = torch.stack([ring[k % 8]["targets"] for k in range(K)]).to(ctx.dev); imu = torch.stack([ring[k % 8]["imu"] for k in range(K)]).to(ctx.dev)
robot.rollout(traj[:4, :, :15].reshape(4, N, 5, 3), traj[:4, :, 15:18], imu_quat_traj=imu[:4], report_active_set=False)
robot.current_joint_config = q0.clone(); robot._mem.copy_(mem0); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18], imu_quat_traj=imu, report_active_set=False); b.record(); torch.cuda.synchronize()
print(f"{'device-resident rollout (one launch)':34s} {N * K / (a.elapsed_time(b) * 1e-3) / 1e6:7.2f} M ticks/s")
