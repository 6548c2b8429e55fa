import sys, numpy as np
sys.path.insert(0,'.')
import bench
from tests import helpers as H
from tests.helpers import OracleQP
name="a1_wx200"
q, targets, mem, ref = bench.cpu_inputs(name, 48, 20260003, 5e-4)
rm = bench._p3_oracle(name, 0.002)
worstA = worstfk = 0.0; vals=[]
for s in range(48):
    r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], solve=True, tail=False)
    A, b, lb, ub, C, Clb, Cub = r["A"], r["b"], r["lb"], r["ub"], r["C"], r["Clb"], r["Cub"]
    x0 = r["qdot"]
    # (ii) A rounded to float32, everything else float64
    A32 = A.astype(np.float32).astype(np.float64)
    qp = OracleQP(A32, b, lb, ub, C.T, Clb, Cub, n_of_velocity_dimensions=A.shape[1]); x1 = np.array(qp.solveQP())
    # (i) b with the error a float32 FK position carries: 3e-8 m relative 6e-8 * 0.5 m, times gain/dt ~ 500
    rng = np.random.default_rng(s)
    db = np.zeros_like(b); db[:36] = rng.normal(0, 3e-8 * 500, size=36)
    qp = OracleQP(A, b + db, lb, ub, C.T, Clb, Cub, n_of_velocity_dimensions=A.shape[1]); x2 = np.array(qp.solveQP())
    vals.append((np.abs(x1-x0).max(), np.abs(x2-x0).max()))
v=np.array(vals)
print("A in float32: max |dqdot| median %.2e worst %.2e" % (np.median(v[:,0]), v[:,0].max()))
print("b with float32-FK error: max |dqdot| median %.2e worst %.2e" % (np.median(v[:,1]), v[:,1].max()))
