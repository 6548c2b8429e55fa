"""Prototype: does ordering the states of a closed-loop tick by their previous iteration count (so that the eight warps of a
phase-barrier group carry similar QPs) shorten the launch?  States are physically permuted between ticks (untimed)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
class A: pass
args = A(); args.robot = "a1_wx200"; args.dt = 0.002; args.seed = 20260003; args.sigma = 5e-4
N = 131072
ctx = bench.Ctx()
robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma, standing=True)
gen = torch.Generator(device=ctx.dev); gen.manual_seed(5)
K = 24
drift = torch.zeros(K, N, 18, dtype=torch.float64, device=ctx.dev)
drift[:, :, 12:18] = torch.randn(K, N, 6, dtype=torch.float64, device=ctx.dev, generator=gen).mul_(1e-4).cumsum(0)
traj = targets[None] + drift
def tick(k, tg):
    e0, e1 = ctx.event(), ctx.event()
    ee, tr = tg[:, :15].reshape(N, 5, 3).contiguous(), tg[:, 15:18].contiguous()
    robot._pack_targets(ee, tr)
    torch.cuda.synchronize()
    e0.record()
    robot.step(ee, tr, advance=True, report_active_set=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
perm_total = torch.arange(N, device=ctx.dev)
for mode in ("unsorted", "sorted"):
    robot2, _ = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma, standing=True)
    robot = robot2
    cur = traj.clone()
    ts = []
    for k in range(K):
        ms = tick(k, cur[k])
        it = robot.last_iters.clone()
        ts.append(ms)
        if mode == "sorted" and k % 4 == 3:
            p = torch.argsort(it, stable=True)
            robot.current_joint_config = robot.current_joint_config[p].contiguous()
            robot._mem.copy_(robot._mem[p]); robot._ref.copy_(robot._ref[p])
            cur = cur[:, p].contiguous()
        if k % 4 == 3:
            print(mode, "tick", k, "ms", round(ms, 4), "iters mean", float(it.double().mean()), "corr-free spread: std", float(it.double().std()))
    print(mode, "mean ms ticks 8..", np.mean(ts[8:]))
