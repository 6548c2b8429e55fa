"""oracle.qp_wrapper -- CPU restatement of ``wrappers/QP_Wrapper.py`` (class ``QP``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference forms H = A^T A, g = -A^T b (QP_Wrapper.py:17-18, 66-67) and hands

    min 1/2 x^T H x + g^T x   s.t.  lb <= x <= ub,  Clb <= C x <= Cub

to qpOASES (``QProblemB.init`` when there is no C, ``SQProblem.init`` / ``.hotstart`` otherwise;
QP_Wrapper.py:23-73), reading back only the primal solution.  qpOASES is a third-party C++
library that is not vendored and not installable here (version unpinned) and the reference
records no QP inputs/outputs => **parity unpinned**.  Because H is positive definite on this
path (the joint-posture task adds (w/nv)^2 I), the minimiser is unique, so any exact active-set
method must return the same x and -- away from degenerate ties -- the same active set.  This
oracle is a textbook Goldfarb-Idnani dual active-set method written with *explicit dense
solves at every iteration* (no factor updating), so that it is numerically independent of the
CUDA solver, which follows the same pivoting rules but updates a J / R factorisation in place.
Every solution carries a KKT certificate (``kkt_residuals``).

Pivoting rules shared with the CUDA kernel (csrc/wbc_qp.cuh):
  * constraints are numbered c = 0..nv-1 (box on x_c) then nv..nv+nC-1 (rows of C);
  * equalities (lo == up) enter first, in index order;
  * the entering constraint is the most violated side, s = min(a.x - lo, up - a.x) < -FEAS_TOL;
    candidates within TIE_ABS + TIE_REL |s_min| of the minimum count as tied and the lowest index wins.  (Mirrored
    rows -- e.g. the +-x faces of a friction pyramid once their partners are active -- tie EXACTLY in exact
    arithmetic; without the window the winner is decided by rounding noise, which differs between an incrementally
    updated C x and a fresh product);
  * the leaving constraint is the active inequality with the smallest ratio u_k / r_k, r_k > 0,
    ties broken by the earliest position in the working set;
  * a candidate whose normal is dependent on the working set (z.n <= DEP_TOL * d.d) takes a
    pure dual step.
"""
import numpy as np

FEAS_TOL = 1e-10
TIE_REL = 1e-9
TIE_ABS = 1e-12
DEP_TOL = 1e-13
PIVOT_REL = 1e-14

STATUS_SOLVED = 0
STATUS_MAXITER = 1
STATUS_INFEASIBLE = 2
STATUS_NOT_PD = 4          # flag or-ed in: a Cholesky pivot was clamped

ACT_NONE, ACT_LOWER, ACT_UPPER, ACT_EQ = 0, 1, 2, 3


def _chol_clamped(H):
    """Cholesky with the pivot clamp the CUDA kernel uses; returns (L, clamped?)."""
    n = H.shape[0]
    L = np.zeros_like(H)
    piv_min = PIVOT_REL * max(float(np.max(np.diag(H))), 0.0)
    clamped = False
    for k in range(n):
        d = H[k, k] - L[k, :k] @ L[k, :k]
        if not (d > piv_min):
            d = piv_min if piv_min > 0 else 1.0
            clamped = True
        L[k, k] = np.sqrt(d)
        for i in range(k + 1, n):
            L[i, k] = (H[i, k] - L[i, :k] @ L[k, :k]) / L[k, k]
    return L, clamped


def solve_qp(H, g, lb, ub, C=None, Clb=None, Cub=None, max_iter=200):
    """Dual active-set solve.  Returns dict(x, status, iters, act, u, working_set)."""
    H = np.asarray(H, dtype=float)
    g = np.asarray(g, dtype=float).reshape(-1)
    n = H.shape[0]
    lb = np.asarray(lb, dtype=float).reshape(n)
    ub = np.asarray(ub, dtype=float).reshape(n)
    if C is None or Clb is None or Cub is None:
        C = np.zeros((0, n))
        Clb = np.zeros(0)
        Cub = np.zeros(0)
    C = np.asarray(C, dtype=float).reshape(-1, n)
    nC = C.shape[0]
    Aall = np.vstack([np.eye(n), C])                       # constraint normals, one per row
    lo = np.concatenate([lb, np.asarray(Clb, dtype=float).reshape(nC)])
    up = np.concatenate([ub, np.asarray(Cub, dtype=float).reshape(nC)])
    m = n + nC

    L, clamped = _chol_clamped(H)
    status = STATUS_NOT_PD if clamped else 0

    Linv = np.linalg.solve(L, np.eye(n))                              # dense, from scratch: no updating

    def Hinv(v):
        return Linv.T @ (Linv @ v)

    x = -Hinv(g)
    W = []          # working set: list of [c, side] with side in {-1 lower, +1 upper, 0 equality}
    u = []          # multipliers, same order
    iters = 0

    def normal(c, side):
        return -Aall[c] if side > 0 else Aall[c]

    def step_dirs(nrm):
        """z = projected H^-1 n, r = dual step, dd = d.d, zn = z.n (textbook GI quantities)."""
        hn = Hinv(nrm)
        dd = float(nrm @ hn)
        if not W:
            return hn, np.zeros(0), dd, dd
        N = np.stack([normal(c, s) for c, s in W], axis=1)          # n x iq
        HN = Hinv(N)
        G = N.T @ HN
        r = np.linalg.solve(G, N.T @ hn)
        z = hn - HN @ r
        return z, r, dd, float(z @ nrm)

    # ---- equalities first --------------------------------------------------------------
    for c in range(m):
        if lo[c] == up[c]:
            nrm = Aall[c]
            z, r, dd, zn = step_dirs(nrm)
            iters += 1
            if zn <= DEP_TOL * dd:
                if abs(float(nrm @ x) - lo[c]) > 1e-8:
                    status |= STATUS_INFEASIBLE
                continue                                              # redundant equality: skip
            t = (lo[c] - float(nrm @ x)) / zn
            x = x + t * z
            u = list(np.asarray(u) - t * r) if W else []
            W.append([c, 0])
            u.append(t)

    # ---- inequalities -------------------------------------------------------------------
    done = bool(status & STATUS_INFEASIBLE)
    while not done:
        ax = Aall @ x
        s_lo = ax - lo
        s_up = up - ax
        in_ws = np.zeros(m, dtype=bool)
        for c, _ in W:
            in_ws[c] = True
        viol = np.minimum(s_lo, s_up)
        viol[in_ws] = 0.0
        vmin = float(np.min(viol))
        if not (vmin < -FEAS_TOL):
            break
        ip = int(np.argmax(viol <= vmin + (TIE_ABS + TIE_REL * abs(vmin))))   # lowest index inside the tie window
        side = -1 if s_lo[ip] <= s_up[ip] else +1
        nrm = normal(ip, side)
        bnd = lo[ip] if side < 0 else -up[ip]                         # n.x >= bnd
        u_new = 0.0
        while True:
            if iters >= max_iter:
                status |= STATUS_MAXITER
                done = True
                break
            iters += 1
            z, r, dd, zn = step_dirs(nrm)
            s_ip = float(nrm @ x) - bnd
            # dual step length: smallest u_k / r_k over active inequalities with r_k > 0
            t1, l = np.inf, -1
            for k, (c, sd) in enumerate(W):
                if sd != 0 and r[k] > 0:
                    ratio = u[k] / r[k]
                    if ratio < t1:
                        t1, l = ratio, k
            dependent = zn <= DEP_TOL * dd
            t2 = np.inf if dependent else -s_ip / zn
            t = min(t1, t2)
            if not np.isfinite(t):
                status |= STATUS_INFEASIBLE
                done = True
                break
            if W:
                u = list(np.asarray(u) - t * r)
            u_new += t
            if dependent:                                             # pure dual step
                del W[l], u[l]
                continue
            x = x + t * z
            if t2 <= t1:                                              # full step: constraint enters
                W.append([ip, side])
                u.append(u_new)
                break
            del W[l], u[l]                                            # partial step: drop and retry

    act = np.zeros(m, dtype=np.int32)
    mult = np.zeros(m)
    for (c, sd), uk in zip(W, u):
        act[c] = ACT_EQ if sd == 0 else (ACT_LOWER if sd < 0 else ACT_UPPER)
        mult[c] = uk
    return {"x": x, "status": status, "iters": iters, "act": act, "u": mult,
            "working_set": [tuple(w) for w in W]}


def kkt_residuals(H, g, lb, ub, C, Clb, Cub, x, act_tol=1e-7):
    """KKT certificate of x for the QP: returns dict of residual norms.

    Multipliers are recovered by least squares on the constraints that are tight at x (within
    act_tol); stationarity H x + g = sum_k lam_k a_k with lam >= 0 on lower-tight sides, lam <= 0
    on upper-tight sides, free sign on equalities.
    """
    H = np.asarray(H, dtype=float)
    n = H.shape[0]
    if C is None:
        C = np.zeros((0, n))
        Clb = np.zeros(0)
        Cub = np.zeros(0)
    Aall = np.vstack([np.eye(n), np.asarray(C, dtype=float).reshape(-1, n)])
    lo = np.concatenate([np.reshape(lb, -1), np.reshape(Clb, -1)])
    up = np.concatenate([np.reshape(ub, -1), np.reshape(Cub, -1)])
    ax = Aall @ x
    primal = float(max(0.0, np.max(lo - ax), np.max(ax - up)))
    tight_lo = np.abs(ax - lo) <= act_tol
    tight_up = np.abs(ax - up) <= act_tol
    tight = np.where(tight_lo | tight_up)[0]
    grad = H @ x + g
    if tight.size:
        lam, *_ = np.linalg.lstsq(Aall[tight].T, grad, rcond=None)
        stat = float(np.max(np.abs(Aall[tight].T @ lam - grad)))
        dual = 0.0
        for k, c in enumerate(tight):
            if tight_lo[c] and tight_up[c]:
                continue
            if tight_lo[c]:
                dual = max(dual, -lam[k])
            else:
                dual = max(dual, lam[k])
    else:
        stat = float(np.max(np.abs(grad)))
        dual = 0.0
    return {"primal": primal, "stationarity": stat, "dual": float(dual)}


class QP:
    """Drop-in restatement of ``QP_Wrapper.QP`` (QP_Wrapper.py:9-73), explicit status/active set added."""

    def __init__(self, A, b, lb, ub, C=None, Clb=None, Cub=None, n_of_velocity_dimensions=None):
        self.lb = lb
        self.ub = ub
        self.Clb = Clb
        self.Cub = Cub
        self.C = C
        self.H = np.dot(A.T, A)                       # QP_Wrapper.py:17
        self.g = np.dot(-A.T, b)                      # QP_Wrapper.py:18
        self.no_solutions = n_of_velocity_dimensions
        self.qp = None
        self.result = None
        self.max_iter = 200        # working-set iteration cap (the reference hands qpOASES nWSR = 100000, QP_Wrapper.py:20)

    def _C_rows(self):
        # The reference passes C.T (Robot_Wrapper4.py:836) and sizes SQProblem with C.shape[1]
        # (QP_Wrapper.py:29); the Cython binding reads the raw row-major buffer, i.e. the
        # original nC x nv matrix.  Accept either orientation: rows = constraints.
        C = np.asarray(self.C, dtype=float)
        n = self.H.shape[0]
        if C.shape[0] == n and C.shape[1] != n:
            C = C.T
        elif C.shape[0] == n and C.shape[1] == n and len(np.reshape(self.Clb, -1)) == n:
            C = C.T
        return np.ascontiguousarray(C)

    def solveQP(self):
        if self.C is None or self.Clb is None or self.Cub is None:
            self.result = solve_qp(self.H, self.g, self.lb, self.ub, max_iter=self.max_iter)
        else:
            self.result = solve_qp(self.H, self.g, self.lb, self.ub, self._C_rows(), self.Clb, self.Cub,
                                   max_iter=self.max_iter)
        self.qp = True
        self.xOpt = self.result["x"]
        return self.xOpt

    def solveQPHotstart(self, A, b, lb, ub, C, Clb, Cub):
        if self.Clb is None or self.Cub is None:
            raise SystemExit("Error, cannot hotstart simply bounded QP")   # QP_Wrapper.py:57-59
        self.lb, self.ub, self.Clb, self.Cub, self.C = lb, ub, Clb, Cub, C
        self.H = np.dot(A.T, A)
        self.g = np.dot(-A.T, b)
        self.result = solve_qp(self.H, self.g, lb, ub, self._C_rows(), Clb, Cub, max_iter=self.max_iter)
        self.xOpt = self.result["x"]
        return self.xOpt
