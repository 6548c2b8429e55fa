import sys, numpy as np, torch
sys.path.insert(0, '.')
import wbc_b200
from wbc_b200 import synthetic
for name, joint, cons in (("a1_wx200", True, dict(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)),
                          ("a1_px100_pin_ver", "HYBRID", dict(CoM=True, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False))):
    N = 37
    r = wbc_b200.RobotModel(name, batch=N, device="cuda:0")
    r.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=joint)
    r.setConstraints(**cons)
    q = synthetic.sample_configurations(r.robot_model, N, 3)
    t = synthetic.load_batch(r, q, synthetic.sample_noise(N, 3, 5e-3))
    r.assemble(t[:, :15].reshape(N, 5, 3), t[:, 15:18])
    for k in range(3):
        r.step(t[:, :15].reshape(N, 5, 3), t[:, 15:18], advance=True)
    r.frameJacobians(2); r.getFrameJacobian(r.trunk_frame_index, 1)
    A = torch.randn(5, 40, 26, dtype=torch.float64, device="cuda:0"); b = torch.randn(5, 40, dtype=torch.float64, device="cuda:0")
    lb = -torch.ones(5, 26, dtype=torch.float64, device="cuda:0"); ub = -lb
    Cm = torch.randn(5, 20, 26, dtype=torch.float64, device="cuda:0"); cl = -torch.ones(5, 20, dtype=torch.float64, device="cuda:0")
    x = wbc_b200.QP(A, b, lb, ub, Cm.transpose(1, 2), cl, -cl, n_of_velocity_dimensions=26).solveQP()
    x = wbc_b200.QP(A[:, :, :7].contiguous(), b, lb[:, :7].contiguous(), ub[:, :7].contiguous(), n_of_velocity_dimensions=7).solveQP()
    torch.cuda.synchronize()
    print(name, "ok", int((r.last_status == 0).sum()), "/", N)
