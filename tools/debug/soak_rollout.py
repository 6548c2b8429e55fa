"""Soak: a 2000-tick closed-loop horizon (the reference's own bootstrap length, Robot_Wrapper4.py:161) as ONE persistent launch, 16384 robots."""
import sys, time, torch
sys.path.insert(0, '.')
import bench
class A: pass
args = A(); args.robot = "a1_wx200"; args.dt = 0.002; args.seed = 20260009; args.sigma = 5e-4
N, K = 16384, 2000
ctx = bench.Ctx()
robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma, standing=True)
t = torch.arange(1, K + 1, dtype=torch.float64, device=ctx.dev)[:, None] * args.dt
traj = targets[None].repeat(K, 1, 1)
traj[:, :, 12] += 0.05 * torch.sin(2 * 3.141592653589793 * 0.5 * t)      # gripper x: 5 cm, 0.5 Hz, 4 s horizon
traj[:, :, 16] += 0.01 * torch.sin(2 * 3.141592653589793 * 0.25 * t)     # trunk sway
imu = robot.current_joint_config[:, 3:7][None].repeat(K, 1, 1).contiguous()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18], imu_quat_traj=imu, report_active_set=False); b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(f"{K} ticks x {N} robots in one launch: {ms:.1f} ms, {N * K / ms / 1e3:.1f} M ticks/s, solved {float((robot.last_status == 0).double().mean()):.4f}, "
      f"finite {bool(torch.isfinite(robot.current_joint_config).all())}, mean iterations {float(robot.last_iters.double().mean()):.2f}")
