"""Compare the reduced QP front with the general one on the same batch (run twice with WBC_B200_NO_REDUCED)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tests.test_gpu_parity import _robot, _load, P1_TASKS, P2_CONS
name, N, seed, sigma = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
robot = _robot(name, N, P1_TASKS, P2_CONS, True)
q, targets = _load(robot, N, seed, sigma)
x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False)
st = robot.last_status.cpu().numpy(); it = robot.last_iters.cpu().numpy(); act = robot.last_active_set.cpu().numpy()
tag = "gen" if os.environ.get("WBC_B200_NO_REDUCED") == "1" else "red"
np.savez(f"gpurun_out/cmp_{tag}.npz", x=x.cpu().numpy(), st=st, it=it, act=act)
print(tag, "status counts", dict(zip(*np.unique(st, return_counts=True))), "mean iters", it.mean())
if tag == "gen" and os.path.exists("gpurun_out/cmp_red.npz"):
    r = np.load("gpurun_out/cmp_red.npz")
    dx = np.abs(r["x"] - x.cpu().numpy()).max(axis=1)
    bad = np.nonzero((r["st"] != st) | (r["it"] != it) | (dx > 1e-7))[0]
    print("differing states", len(bad), bad[:10])
    for s in bad[:6]:
        print(s, "st", r["st"][s], st[s], "it", r["it"][s], it[s], "dx", dx[s], "act", [hex(int(v)) for v in r["act"][s]], [hex(int(v)) for v in act[s]])
    print("max dx overall", dx.max())
