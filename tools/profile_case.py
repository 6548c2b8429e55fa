"""Launch sequence for ncu: 4 open-loop ticks (the bench workload: A1+WX200, P3, 131072 states) then 4 closed-loop ticks
(the runWBC tick: same kernel instantiation with the in-place tail: task memory, integrate, IMU feedback, base estimate).
    ncu --set full --clock-control none --import-source on -k regex:wbc_step_kernel -s 3 -c 1 ...   -> open-loop launch #4
    ncu ... -s 7 -c 1                                                                          -> closed-loop launch #4
"""
import sys
import torch
sys.path.insert(0, '.')
import bench

class A: pass
args = A(); args.robot = "a1_wx200"; args.dt = 0.002; args.seed = 20260003; args.sigma = 5e-4
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ctx = bench.Ctx()
robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma)
one = bench.stepper(robot, targets)
for _ in range(4):
    one()
torch.cuda.synchronize()
ee, tr = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
K = 4
robot.rollout(ee[None].repeat(K, 1, 1, 1), tr[None].repeat(K, 1, 1), imu_quat_traj=robot.current_joint_config[:, 3:7][None].repeat(K, 1, 1))
torch.cuda.synchronize()
print("ok", float((robot.last_status == 0).double().mean()))
