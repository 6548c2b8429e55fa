import sys, time, torch
sys.path.insert(0, '.')
import bench
class A: pass
args = A(); args.robot = "a1_px100_pin_ver"; args.dt = 0.002; args.seed = 20260001; args.sigma = 5e-3
ctx = bench.Ctx()
for N in (1, 4096, 16384):
    robot, targets = bench.make_robot(ctx, args.robot, N, args.dt, bench.ALL_TASKS, bench.P2_CONS, True, args.seed, args.sigma)
    one = bench.stepper(robot, targets)
    for _ in range(50): one()
    torch.cuda.synchronize()
    K = 2000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K): one()
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"N {N}: host issue {1e6*(t1-t0)/K:.1f} us/call, wall incl. drain {1e6*(t2-t0)/K:.1f} us/call, gpu events {1e3*e0.elapsed_time(e1)/K:.1f} us/call")
