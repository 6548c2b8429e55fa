"""Batched drop-in for ``wrappers/QP_Wrapper.py: class QP`` (reference :9-73).

``QP(A, b, lb, ub, C, Clb, Cub, n_of_velocity_dimensions).solveQP()`` keeps its arguments; every
array carries a leading batch dimension N (a single un-batched problem is accepted too and
treated as N = 1).  H = A^T A and g = -A^T b (:17-18) are formed inside the CUDA kernel; the
solve replaces qpOASES' ``QProblemB.init`` / ``SQProblem.init`` / ``.hotstart`` (:26-51, :70) by
the warp-per-problem dual active-set solver of ``csrc/wbc_qp.cuh``.  Unlike the reference, the
solver's report is kept: ``status``, ``iters``, ``active_set``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi as cabi


def _dev_tensor(x, device):
    if x is None:
        return None
    if not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x, dtype=np.float64))
    return x.to(device=device, dtype=torch.float64)


class QP:
    def __init__(self, A, b, lb, ub, C=None, Clb=None, Cub=None, n_of_velocity_dimensions=None, *, device=None,
                 max_iter=200):
        if not torch.cuda.is_available():
            raise cabi.WbcError("QP needs a CUDA device (B200): there is no CPU fallback")
        self._lib = cabi.load()
        if device is None:
            device = A.device if torch.is_tensor(A) and A.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.max_iter = int(max_iter)
        self.no_solutions = n_of_velocity_dimensions
        self.qp = None
        self.xOpt = None
        self.status = self.iters = self.active_set = None
        self._set(A, b, lb, ub, C, Clb, Cub)

    def _set(self, A, b, lb, ub, C_, Clb, Cub):
        dev = self.device
        A = _dev_tensor(A, dev)
        self._batched = A.dim() == 3
        if not self._batched:
            A = A.unsqueeze(0)
        N, m, nv = A.shape
        if self.no_solutions is None:
            self.no_solutions = nv
        self.A = A.contiguous()
        self.b = _dev_tensor(b, dev).reshape(N, m).contiguous()
        self.lb = _dev_tensor(lb, dev).reshape(N, nv).contiguous()
        self.ub = _dev_tensor(ub, dev).reshape(N, nv).contiguous()
        if C_ is None or Clb is None or Cub is None:
            self.C = self.Clb = self.Cub = None
            self.nC = 0
        else:
            self.Clb = _dev_tensor(Clb, dev).reshape(N, -1).contiguous()
            self.Cub = _dev_tensor(Cub, dev).reshape(N, -1).contiguous()
            nC = self.Clb.shape[1]
            Ct = _dev_tensor(C_, dev)
            if not self._batched:
                Ct = Ct.unsqueeze(0)
            # the reference hands over C.T, shape (nv, nC) (Robot_Wrapper4.py:836); rows-as-constraints is accepted too
            if Ct.shape[1] == nv and Ct.shape[2] == nC:
                Ct = Ct.transpose(1, 2)
            elif not (Ct.shape[1] == nC and Ct.shape[2] == nv):
                raise ValueError(f"C has shape {tuple(Ct.shape)}, expected (N, {nv}, {nC}) or (N, {nC}, {nv})")
            self.C = Ct.contiguous()
            self.nC = nC
        self.N, self.m, self.nv = N, m, nv

    def _solve(self):
        N, nv = self.N, self.nv
        x = torch.empty(N, nv, dtype=torch.float64, device=self.device)
        status = torch.empty(N, dtype=torch.int32, device=self.device)
        iters = torch.empty(N, dtype=torch.int32, device=self.device)
        act = torch.empty(N, 2, dtype=torch.int64, device=self.device)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)  # noqa: E731
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_qp_solve(N, nv, self.m, self.nC, p(self.A), p(self.b), None, None, p(self.lb),
                                              p(self.ub), p(self.C), p(self.Clb), p(self.Cub), self.max_iter, p(x),
                                              p(status), p(iters), p(act),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.status, self.iters, self.active_set = status, iters, act
        self.xOpt = x if self._batched else x[0]
        self.qp = True
        return self.xOpt

    @property
    def H(self):
        return torch.matmul(self.A.transpose(1, 2), self.A)

    @property
    def g(self):
        return -torch.matmul(self.A.transpose(1, 2), self.b.unsqueeze(-1)).squeeze(-1)

    def solveQP(self):
        return self._solve()

    def solveQPHotstart(self, A, b, lb, ub, C, Clb, Cub):
        if self.Clb is None or self.Cub is None:
            # the reference prints and calls exit() here (QP_Wrapper.py:57-59)
            raise SystemExit("Error, cannot hotstart simply bounded QP")
        self._set(A, b, lb, ub, C, Clb, Cub)
        return self._solve()
