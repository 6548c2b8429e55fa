"""Per-tick QP statistics of a closed-loop rollout (debug)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench

class A: pass
args = A(); args.robot = "a1_wx200"; args.dt = 0.002; args.seed = 20260003; args.sigma = 5e-4
N, K = 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "const"
ctx = bench.Ctx()
robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma,
                                  standing=(len(sys.argv) > 3 and sys.argv[3] == "standing"))
ee, tr = targets[:, :15].reshape(N, 5, 3), targets[:, 15:18]
gen = torch.Generator(device=ctx.dev); gen.manual_seed(5)
drift = torch.zeros(K, N, 18, dtype=torch.float64, device=ctx.dev)
if mode == "walk":
    drift[:, :, 12:18] = torch.randn(K, N, 6, dtype=torch.float64, device=ctx.dev, generator=gen).mul_(1e-4).cumsum(0)
traj = targets[None] + drift
for k in range(K):
    robot.step(traj[k, :, :15].reshape(N, 5, 3), traj[k, :, 15:18], advance=True)
    it = robot.last_iters.cpu().numpy(); st = robot.last_status.cpu().numpy()
    act = robot.last_active_set.cpu().numpy().astype(np.uint64)
    nbox = np.array([bin(int(a)).count("1") for a in act[:2048, 0]]).mean() - 3
    nrow = np.array([bin(int(a) & 0xFF).count("1") for a in act[:2048, 1]]).mean()
    print(f"tick {k:3d} iters mean {it.mean():6.2f} p50 {np.median(it):4.0f} p99 {np.percentile(it, 99):5.0f} max {it.max():4d}  status!=0 {np.mean(st != 0):.4f}"
          f"  active box {nbox:.2f} trunk rows {nrow:.2f}  |qdot| max {float(robot.qdot.abs().max()):.2f}")
