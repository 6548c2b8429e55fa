// wbc_qp.cuh -- one-warp-per-problem dense QP solver (Goldfarb-Idnani dual active set).
//
//   min 1/2 x^T H x + g^T x   s.t.  lb <= x <= ub,  Clb <= C x <= Cub,      H = A^T A > 0
//
// Replaces QP.solveQP / QP.solveQPHotstart of the reference (wrappers/QP_Wrapper.py:23-73), which
// hand the same problem to qpOASES (QProblemB.init / SQProblem.init / .hotstart).  qpOASES is an
// active-set method that is exact up to its termination tolerance; H is positive definite on this
// path, so the minimiser is unique and any exact active-set method returns it.  The pivoting rules
// are the ones documented in oracle/qp_wrapper.py (the CPU oracle uses explicit dense solves at
// every iteration; this kernel updates the factorisation J = L^-T Q, R = Q^T L^-1 N in place).
//
// Data placement (n <= 32 variables, nC <= 32 rows, one warp):
//   lane k    : x_k, g_k, lb_k, ub_k, status of box k      |  lane r : Clb_r, Cub_r, status of row r
//   lane pos  : working-set entry `pos` (constraint id, side, multiplier, R column slot, 1/R_pos,pos)
//   shared    : M0 [n][LD]  H -> Cholesky L -> R (upper triangular, columns addressed through slots)
//               J  [n][LD]  L^-T, updated by Householder reflections (add) / Givens rotations (drop)
//               C  [nC][LD] constraint rows
//               vx, vd, vg [32] broadcast copies of x, d, g
// LD is odd so that both row-wise (lane = row) and column-wise (lane = column) sweeps are free of
// shared-memory bank conflicts beyond the 2-wavefront minimum of 64-bit accesses.
//
// This file holds the run-time-size body (1 <= n <= 32, everything in shared memory) behind the generic
// QP(A, b, ...) drop-in; the robot sizes (nv = 25 / 26) use the register-resident solver of
// wbc_qp_reg.cuh, which follows the same algorithm and the same pivoting decisions.
#pragma once
#include "wbc_device.cuh"

#define WBC_QP_FEAS_TOL 1e-10
// entering constraint: candidates within TIE_ABS + TIE_REL |min| of the most violated one are tied, the lowest index
// wins (oracle/qp_wrapper.py).  Mirrored rows (the +-x faces of a friction pyramid once their partners are active) tie
// exactly in exact arithmetic; without the window rounding noise decides, and the incrementally updated C x of the
// register solver rounds differently from the fresh products of the oracles.
#define WBC_QP_TIE_REL 1e-9
#define WBC_QP_TIE_ABS 1e-12
#define WBC_QP_DEP_TOL 1e-13
#define WBC_QP_PIVOT_REL 1e-14

#define WBC_REP32_DESC(X) X(31) X(30) X(29) X(28) X(27) X(26) X(25) X(24) X(23) X(22) X(21) X(20) X(19) X(18) X(17) X(16) \
                          X(15) X(14) X(13) X(12) X(11) X(10) X(9) X(8) X(7) X(6) X(5) X(4) X(3) X(2) X(1) X(0)
#define WBC_REP32_ASC(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) \
                         X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31)

struct QpShared {
  double* M0;
  double* J;
  double* C;
  double* vx;
  double* vd;
  double* vg;
};

struct QpResult {
  int status;
  int iters;
  unsigned long long act_box;
  unsigned long long act_rows;
};

// argmin over the warp of (val, idx); ties -> smallest idx.  Result uniform across lanes.
__device__ __forceinline__ void warp_argmin(double& val, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(WBC_FULL_MASK, val, o);
    const int oi = __shfl_xor_sync(WBC_FULL_MASK, idx, o);
    if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
}

// spread the 32 bits of v to the even bit positions of a 64-bit word
__device__ __forceinline__ unsigned long long spread_bits(unsigned v) {
  unsigned long long x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}

__device__ __forceinline__ void pack_active_sets(int lane, int n, int nC, int bstat, int cstat, QpResult& res) {
  const unsigned lo_b = __ballot_sync(WBC_FULL_MASK, lane < n && (bstat & 1));
  const unsigned up_b = __ballot_sync(WBC_FULL_MASK, lane < n && (bstat & 2));
  const unsigned lo_r = __ballot_sync(WBC_FULL_MASK, lane < nC && (cstat & 1));
  const unsigned up_r = __ballot_sync(WBC_FULL_MASK, lane < nC && (cstat & 2));
  res.act_box = spread_bits(lo_b) | (spread_bits(up_b) << 1);
  res.act_rows = spread_bits(lo_r) | (spread_bits(up_r) << 1);
}

// =================================================================================================
// run-time size (generic QP drop-in)
// =================================================================================================
__device__ __forceinline__ QpResult warp_qp_solve_rt(const QpShared S, const int n, const int LD, const int nC,
                                                     const double g, const double lb, const double ub, const double clb,
                                                     const double cub, const int max_iter, double& x_out) {
  const int lane = threadIdx.x & 31;
  double* __restrict__ M0 = S.M0;
  double* __restrict__ J = S.J;
  const double* __restrict__ C = S.C;
  QpResult res;
  res.status = 0;
  res.iters = 0;

  double hd = (lane < n) ? M0[lane * LD + lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hd = fmax(hd, __shfl_xor_sync(WBC_FULL_MASK, hd, o));
  const double piv_min = WBC_QP_PIVOT_REL * fmax(hd, 0.0);
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    double s = 0.0;
    if (lane >= k && lane < n) {
      s = M0[lane * LD + k];
      const double* rowi = M0 + lane * LD;
      const double* rowk = M0 + k * LD;
#pragma unroll 2
      for (int j = 0; j < k; ++j) s -= rowi[j] * rowk[j];
    }
    double dk = __shfl_sync(WBC_FULL_MASK, s, k);
    if (!(dk > piv_min)) {
      dk = piv_min > 0.0 ? piv_min : 1.0;
      res.status |= WBC_QP_NOT_PD;
    }
    const double r = rsqrt(dk);
    __syncwarp();
    if (lane == k) {
      M0[k * LD + k] = dk * r;
      S.vd[k] = r;
    } else if (lane > k && lane < n) {
      M0[lane * LD + k] = s * r;
    }
    __syncwarp();
  }
  if (lane < n) {
    double* rowc = J + lane * LD;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      double acc = (i == lane) ? 1.0 : 0.0;
      const double* Li = M0 + i * LD;
#pragma unroll 2
      for (int j = 0; j < i; ++j) acc -= Li[j] * rowc[j];
      rowc[i] = acc * S.vd[i];
    }
  }
  S.vg[lane] = (lane < n) ? g : 0.0;
  __syncwarp();
  double x = 0.0;
  {
    double w = 0.0;
    if (lane < n)
      for (int i = 0; i <= lane; ++i) w += J[i * LD + lane] * S.vg[i];
    __syncwarp();
    S.vd[lane] = w;
    __syncwarp();
    if (lane < n) {
      const double* rowi = J + lane * LD;
      for (int j = lane; j < n; ++j) x -= rowi[j] * S.vd[j];
    }
  }
  __syncwarp();

  int iq = 0, p_eq = 0;
  int ws_c = -1, slot = lane;
  double u = 0.0, rinv = 0.0;
  int bstat = 0, cstat = 0;
  unsigned eq_mask_box = __ballot_sync(WBC_FULL_MASK, lane < n && lb == ub);
  unsigned eq_mask_row = __ballot_sync(WBC_FULL_MASK, lane < nC && clb == cub);
  bool done = false;
#pragma unroll 1
  while (!done) {
    int ip, side;
    bool is_eq = false;
    if (eq_mask_box) {
      ip = __ffs(eq_mask_box) - 1;
      eq_mask_box &= eq_mask_box - 1;
      side = -1; is_eq = true;
    } else if (eq_mask_row) {
      ip = n + __ffs(eq_mask_row) - 1;
      eq_mask_row &= eq_mask_row - 1;
      side = -1; is_eq = true;
    } else {
      S.vx[lane] = x;
      __syncwarp();
      double vb = INFINITY, vc = INFINITY;
      int myside_b = -1, myside_c = -1;
      if (lane < n && bstat == 0) {
        const double slo = x - lb, sup = ub - x;
        vb = fmin(slo, sup);
        myside_b = (slo <= sup) ? -1 : +1;
      }
      if (lane < nC && cstat == 0) {
        const double* Cr = C + lane * LD;
        double ax = 0.0;
        for (int j = 0; j < n; ++j) ax += Cr[j] * S.vx[j];
        const double slo = ax - clb, sup = cub - ax;
        vc = fmin(slo, sup);
        myside_c = (slo <= sup) ? -1 : +1;
      }
      double best = fmin(vb, vc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(WBC_FULL_MASK, best, o));
      if (!(best < -WBC_QP_FEAS_TOL)) break;
      const double thr = best + (WBC_QP_TIE_ABS + WBC_QP_TIE_REL * fabs(best));   // lowest index inside the tie window
      const unsigned wb = __ballot_sync(WBC_FULL_MASK, vb <= thr), wc = __ballot_sync(WBC_FULL_MASK, vc <= thr);
      ip = wb ? __ffs(wb) - 1 : n + __ffs(wc) - 1;
      const int src = (ip < n) ? ip : ip - n;
      side = __shfl_sync(WBC_FULL_MASK, (ip < n) ? myside_b : myside_c, src);
    }
    const double sgn = (side > 0) ? -1.0 : 1.0;
    double u_new = 0.0;
#pragma unroll 1
    while (true) {
      if (!is_eq && res.iters >= max_iter) { res.status |= WBC_QP_MAXITER; done = true; break; }
      res.iters++;
      double d = 0.0;
      if (lane < n) {
        if (ip < n) {
          d = sgn * J[ip * LD + lane];
        } else {
          const double* Cr = C + (ip - n) * LD;
          for (int i = 0; i < n; ++i) d += J[i * LD + lane] * Cr[i];
          d *= sgn;
        }
      }
      S.vd[lane] = d;
      double dd = d * d, dd2 = (lane >= iq) ? d * d : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dd += __shfl_xor_sync(WBC_FULL_MASK, dd, o);
        dd2 += __shfl_xor_sync(WBC_FULL_MASK, dd2, o);
      }
      __syncwarp();
      double z = 0.0;
      if (lane < n) {
        const double* rowi = J + lane * LD;
        for (int j = iq; j < n; ++j) z += rowi[j] * S.vd[j];
      }
      double rr = (lane < iq) ? d : 0.0;
#pragma unroll 1
      for (int k = iq - 1; k >= p_eq; --k) {
        const int slot_k = __shfl_sync(WBC_FULL_MASK, slot, k);
        const double rk = __shfl_sync(WBC_FULL_MASK, rr * rinv, k);
        if (lane == k) rr = rk;
        else if (lane < k && lane >= p_eq) rr -= M0[lane * LD + slot_k] * rk;
      }
      double s_ip;
      if (ip < n) {
        const double xi = __shfl_sync(WBC_FULL_MASK, x, ip);
        const double lo_i = __shfl_sync(WBC_FULL_MASK, lb, ip), up_i = __shfl_sync(WBC_FULL_MASK, ub, ip);
        s_ip = (side > 0) ? (up_i - xi) : (xi - lo_i);
      } else {
        const double* Cr = C + (ip - n) * LD;
        const double ax = warp_sum((lane < n) ? Cr[lane] * x : 0.0);
        const double lo_i = __shfl_sync(WBC_FULL_MASK, clb, ip - n), up_i = __shfl_sync(WBC_FULL_MASK, cub, ip - n);
        s_ip = (side > 0) ? (up_i - ax) : (ax - lo_i);
      }
      const bool dependent = dd2 <= WBC_QP_DEP_TOL * dd;
      if (is_eq) {
        if (dependent) {
          if (fabs(s_ip) > 1e-8) res.status |= WBC_QP_INFEASIBLE;
          break;
        }
        const double t = -s_ip / dd2;
        x += t * z;
        u_new = t;
      } else {
        double t1 = INFINITY;
        int l = 0x7fffffff;
        if (lane >= p_eq && lane < iq && rr > 0.0) { t1 = u / rr; l = lane; }
        warp_argmin(t1, l);
        const double t2 = dependent ? INFINITY : -s_ip / dd2;
        const double t = fmin(t1, t2);
        if (!(t < INFINITY)) { res.status |= WBC_QP_INFEASIBLE; done = true; break; }
        if (lane >= p_eq && lane < iq) u -= t * rr;
        u_new += t;
        if (!dependent) x += t * z;
        if (dependent || !(t2 <= t1)) {
          const int c_drop = __shfl_sync(WBC_FULL_MASK, ws_c, l);
#pragma unroll 1
          for (int k = l; k < iq - 1; ++k) {
            const int slot_k1 = __shfl_sync(WBC_FULL_MASK, slot, k + 1);
            const double a = M0[k * LD + slot_k1], b = M0[(k + 1) * LD + slot_k1];
            const double rho = sqrt(a * a + b * b);
            const double cg = (rho > 0.0) ? a / rho : 1.0, sg = (rho > 0.0) ? b / rho : 0.0;
            __syncwarp();
            if (lane > k && lane < iq) {
              const double r0 = M0[k * LD + slot], r1 = M0[(k + 1) * LD + slot];
              M0[k * LD + slot] = cg * r0 + sg * r1;
              M0[(k + 1) * LD + slot] = -sg * r0 + cg * r1;
            }
            if (lane < n) {
              const double j0 = J[lane * LD + k], j1 = J[lane * LD + k + 1];
              J[lane * LD + k] = cg * j0 + sg * j1;
              J[lane * LD + k + 1] = -sg * j0 + cg * j1;
            }
            __syncwarp();
          }
          {
            const int dropped_slot = __shfl_sync(WBC_FULL_MASK, slot, l);
            const int nc_ = __shfl_down_sync(WBC_FULL_MASK, ws_c, 1);
            const int nslot = __shfl_down_sync(WBC_FULL_MASK, slot, 1);
            const double nu = __shfl_down_sync(WBC_FULL_MASK, u, 1);
            if (lane >= l && lane < iq - 1) { ws_c = nc_; slot = nslot; u = nu; }
            if (lane == iq - 1) { slot = dropped_slot; ws_c = -1; u = 0.0; }
            if (lane >= l && lane < iq - 1) rinv = 1.0 / M0[lane * LD + slot];
            if (c_drop < n) { if (lane == c_drop) bstat = 0; }
            else if (lane == c_drop - n) cstat = 0;
            iq--;
          }
          __syncwarp();
          continue;
        }
      }
      {
        const double d_iq = __shfl_sync(WBC_FULL_MASK, d, iq);
        const double nrm = sqrt(dd2);
        const double sigma = (d_iq >= 0.0) ? nrm : -nrm;
        const double v_iq = d_iq + sigma;
        const double beta = 1.0 / (sigma * v_iq);
        if (lane == iq) S.vd[iq] = v_iq;
        __syncwarp();
        if (lane < n) {
          double* rowi = J + lane * LD;
          const double bw = beta * (z + sigma * rowi[iq]);
          for (int j = iq; j < n; ++j) rowi[j] -= bw * S.vd[j];
        }
        const int slot_new = __shfl_sync(WBC_FULL_MASK, slot, iq);
        if (lane < iq) M0[lane * LD + slot_new] = d;
        if (lane == iq) {
          M0[iq * LD + slot_new] = -sigma;
          rinv = -1.0 / sigma;
          ws_c = ip;
          u = u_new;
        }
        const int st = is_eq ? 3 : (side > 0 ? 2 : 1);
        if (ip < n) { if (lane == ip) bstat = st; }
        else if (lane == ip - n) cstat = st;
        iq++;
        if (is_eq) p_eq = iq;
        __syncwarp();
      }
      break;
    }
  }
  x_out = x;
  pack_active_sets(lane, n, nC, bstat, cstat, res);
  return res;
}
