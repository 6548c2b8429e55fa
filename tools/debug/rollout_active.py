import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
class A: pass
args = A(); args.robot = "a1_wx200"; args.dt = 0.002; args.seed = 20260003; args.sigma = 5e-4
N = 16384
ctx = bench.Ctx()
plain = len(sys.argv) > 1 and sys.argv[1] == "plain"
robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma)
for k in range(3):
    lbub = robot.velDamperJointConstraints()
    robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=True, plain_integrate=plain)
    act = robot.last_active_set.cpu().numpy().astype(np.uint64)
    lo = np.array([[(int(a) >> (2 * c)) & 1 for c in range(26)] for a in act[:4096, 0]]).mean(0)
    up = np.array([[(int(a) >> (2 * c + 1)) & 1 for c in range(26)] for a in act[:4096, 0]]).mean(0)
    print("tick", k, "iters", float(robot.last_iters.double().mean()))
    print(" lower active:", np.round(lo, 2))
    print(" upper active:", np.round(up, 2))
    print(" lb mean:", np.round(lbub[0].mean(0).cpu().numpy(), 2))
    print(" ub mean:", np.round(lbub[1].mean(0).cpu().numpy(), 2))
    print(" |qdot| mean per dof:", np.round(robot.qdot.abs().median(0).values.cpu().numpy(), 3))
