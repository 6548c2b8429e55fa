import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tests.test_gpu_parity import _robot, _load, P1_TASKS, P2_CONS
name, N, seed, sigma = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
robot = _robot(name, N, P1_TASKS, P2_CONS, True)
q, targets = _load(robot, N, seed, sigma)
asm = robot.assemble(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], want=("C", "Clb", "Cub", "lb", "ub", "H", "g"))
x = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False)
lb, ub = asm["lb"].cpu().numpy(), asm["ub"].cpu().numpy()
nq = lb.shape[1] - 3
eqs = np.nonzero((lb[:, :nq] == ub[:, :nq]).any(axis=1))[0]
print("states with lb == ub among free DoFs:", eqs)
st = robot.last_status.cpu().numpy()
print("bad status:", np.nonzero(st)[0])
Cx = torch.einsum("nrk,nk->nr", asm["C"], x).cpu().numpy()
for s in [360, 1680, 3912]:
    print(s, "Cx trunk", Cx[s, :4], "clb", asm["Clb"][s, :4].cpu().numpy(), "cub", asm["Cub"][s, :4].cpu().numpy())
    print("   C trunk rows leg cols max", np.abs(asm["C"][s, :4, 6:18].cpu().numpy()).max(), "x", x[s].cpu().numpy()[:8])
