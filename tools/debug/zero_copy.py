"""Experiment: the fused tick reading q / targets straight from pinned host memory and writing qdot / status / iters
straight into pinned host memory (no staging copies, one launch)."""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
import wbc_b200
from wbc_b200 import synthetic, _cabi as cabi
N = 131072
dev = "cuda:0"
robot = wbc_b200.RobotModel("a1_wx200", batch=N, device=dev, dt=0.002)
robot.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
robot.setConstraints(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
q = synthetic.sample_configurations(robot.robot_model, N, 20260003)
noise = synthetic.sample_noise(N, 20260003, 5e-4)
targets = synthetic.load_batch(robot, q, noise)
x_ref = robot.step(targets[:, :15].reshape(N, 5, 3), targets[:, 15:18], advance=False, report_active_set=False).clone()
hq = robot.current_joint_config.cpu().pin_memory(); ht = targets.cpu().pin_memory()
hx = torch.empty(N, 26, dtype=torch.float64).pin_memory(); hs = torch.empty(N, dtype=torch.int32).pin_memory(); hi = torch.empty(N, dtype=torch.int32).pin_memory()
lib = cabi.load(); cfg = robot._config()
def run(inp_host, out_host, steps=10):
    io = cabi.WbcStepIO()
    io.q = hq.data_ptr() if inp_host else robot.current_joint_config.data_ptr()
    io.targets = ht.data_ptr() if inp_host else targets.data_ptr()
    io.mem_in = robot._mem.data_ptr(); io.ref = robot._ref.data_ptr(); io.dt = 0.002
    io.qdot = hx.data_ptr() if out_host else robot.qdot.data_ptr()
    io.status = hs.data_ptr() if out_host else robot.last_status.data_ptr()
    io.iters = hi.data_ptr() if out_host else robot.last_iters.data_ptr()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3): cabi.check(lib.wbc_step(robot._model, C.byref(cfg), C.byref(io), N, sp))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): cabi.check(lib.wbc_step(robot._model, C.byref(cfg), C.byref(io), N, sp))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ok = bool(torch.equal(hx, x_ref.cpu())) if out_host else True
    print(f"inputs from host {inp_host}, outputs to host {out_host}: {N / ms / 1e3:.1f} M steps/s ({ms:.3f} ms), identical {ok}")
run(False, False); run(False, True); run(True, False); run(True, True)
