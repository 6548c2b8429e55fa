"""Status histogram over a closed-loop rollout (bench workload), reduced vs general front."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import wbc_b200
from wbc_b200 import synthetic
N, K = 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = "cuda:0"
robot = wbc_b200.RobotModel("a1_wx200", batch=N, device=dev, dt=0.002)
robot.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
robot.setConstraints(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
q = synthetic.sample_configurations(robot.robot_model, N, 20260003)
noise = synthetic.sample_noise(N, 20260003, 5e-4)
targets = synthetic.load_batch(robot, q, noise)
gen = torch.Generator(device=dev); gen.manual_seed(20260008)
drift = torch.zeros(K, N, 18, dtype=torch.float64, device=dev)
drift[:, :, 12:18] = torch.randn(K, N, 6, dtype=torch.float64, device=dev, generator=gen).mul_(1e-4).cumsum(0)
traj = targets[None] + drift
qh, xh, sh = robot.rollout(traj[:, :, :15].reshape(K, N, 5, 3), traj[:, :, 15:18], record=True)
sh = sh.cpu().numpy()
for k in (0, 1, 2, 5, 10, 20, 50, K - 1):
    u, c = np.unique(sh[k], return_counts=True)
    print("tick", k, dict(zip(u.tolist(), c.tolist())), "max|qdot|", float(xh[k].abs().max()), "finite", bool(torch.isfinite(xh[k]).all()))
first_bad = (sh != 0).argmax(axis=0)
bad = (sh != 0).any(axis=0)
print("robots ever unsolved", int(bad.sum()), "first bad tick histogram", np.bincount(first_bad[bad])[:20])
