"""CPU experiment: config 3 (friction pyramid rows) -- C oracle (factor updating) vs NumPy oracle (dense re-solves):
how often do the two take different pivoting paths?  (A proxy for the kernel-vs-oracle disagreement seen on the GPU.)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import helpers as H
from oracle import c_port
import bench

def config3_rows(mu=0.6, big=1e30):
    rows = []
    for foot in range(4):
        for cx, cy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            rows.append((foot, 2, [cx, cy, -mu, 0, 0, 0], -big, 0.0))
    return rows

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    name = "a1_wx200"
    q, targets, mem, ref = bench.cpu_inputs(name, n, 20260003, 5e-3)
    rm = H.make_oracle(name)
    rm.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
    rm.setConstraints(CoM=False, Trunk=True, FR=False, FL=False, RR=False, RL=False, Grip=False)
    from wbc_b200 import synthetic, TreeTable
    rm.extra_rows = synthetic.config3_rows(TreeTable.load(name))
    print(len(rm.extra_rows), "extra rows")
    ts, table = c_port.table_struct(name)
    cref = c_port.step(ts, c_port.config_struct(rm, table), q, targets, mem, ref, rm.dt)
    same_it = same_set = 0
    worst = 0.0
    bad = []
    for s in range(n):
        r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], solve=True, tail=False)
        wb, wr = H.act_to_bits(r["act"], 26)
        si = r["iters"] == cref["iters"][s]
        ss = (wb == int(cref["active_set"][s, 0])) and (wr == int(cref["active_set"][s, 1]))
        same_it += si; same_set += ss
        worst = max(worst, np.abs(r["qdot"] - cref["qdot"][s]).max())
        if not (si and ss):
            bad.append((s, r["iters"], int(cref["iters"][s]), hex(wr), hex(int(cref["active_set"][s, 1]))))
    print("same iters", same_it / n, "same set", same_set / n, "max dx", worst)
    for b_ in bad[:20]:
        print(b_)
