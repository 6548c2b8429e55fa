// wbc_qp_reg.cuh -- register-resident dual active-set QP solver, one warp per problem (compile-time size).
//
//   min 1/2 x^T H x + g^T x   s.t.  lb <= x <= ub,  Clb <= C x <= Cub,      H = A^T A > 0
//
// Same method and the same pivoting rules as warp_qp_solve_rt (wbc_qp.cuh) and oracle/qp_wrapper.py
// (Goldfarb-Idnani; replaces QP.solveQP / solveQPHotstart, wrappers/QP_Wrapper.py:23-73), but the
// factor J = L^-T Q never lives in shared memory.  A matrix-vector product whose matrix sits in shared
// memory is capped by the LDS pipe at 16 FMA/clk/SM (a quarter of the FP64 pipe); here every matrix
// element is a register and shared memory only carries broadcast vectors (one wavefront per load).
//
// Data placement:
//   lane i < NV   Jr[0..NV)  = row i of J            = J^T e_i   -- the "d vector" of box constraint i
//   lane NV       Jr[0..NV)  = L^-1 g                            -- rides along in the spare lane for x0
//   lane c < nC   Dr[0..NV)  = J^T C[c]^T                        -- the "d vector" of row constraint c
//                 (SPLIT, nC <= 16: lane c holds elements [0, HALF) and lane c + 16 elements [HALF, NV) of it,
//                  which frees 2 * (NV - HALF) registers per thread and a quarter of the FMAs of every pass)
//   lane i        x_i, lb_i, ub_i, box status     |  lane c: (C x)_c, Clb_c, Cub_c, row status
//   lane pos      working-set entry `pos` (constraint id, multiplier, R column slot, 1/R_pos,pos)
//   shared        R [NV][LD] triangular factor of the inequality block, vd[32] broadcast vector,
//                 (the same block holds the columns of L while H is factorised), col[32] = 1 / L_kk
//
// Consequences:
//   * H = L L^T is right-looking with broadcast columns; the forward substitutions L^-1 [I | g | C^T] then
//     run for all right-hand sides at once (one broadcast load feeds two FMAs per lane);
//   * d = J^T n is never computed: every constraint carries its own d, updated by the same reflector as J;
//   * the reflector (norm, sigma, beta) is computed redundantly by all lanes from the broadcast d: no warp
//     reductions in the iteration;
//   * C x is tracked incrementally (C z = Dr . d2 falls out of the reflector application).
//   * variables with lb == ub are eliminated before the factorisation (row/column of H replaced by the
//     identity, g and the row bounds shifted); they are reported as equality-active and counted as one
//     working-set change each, as the oracle counts them.
//
// Partial sums "over j >= iq" are full-length static loops over a broadcast vector whose first iq entries
// were zeroed in shared memory (one predicated store), so every register index is a compile-time constant and
// the passes are straight-line code.  Shared memory is addressed as [32-bit base + immediate] (wbc_device.cuh).
#pragma once
#include "wbc_qp.cuh"

#define WBC_LDT 38    // doubles per column of the transposed task rows: 304 B = 19 x 16 B, conflict-free 128-bit accesses

struct QpRegShared {   // 32-bit shared-window addresses (smem_addr), all 16-byte aligned
  uint32_t R;        // [NV][NV + 2]: columns of L during the factorisation, then R with LD = NV | 1
  uint32_t col;      // [64]: 1 / L_kk
  uint32_t vd;       // [32]: broadcast vector
  uint32_t C;        // [nC][LD] constraint rows (read once; up to two doubles past the end are touched)
  uint32_t clb, cub; // [32] each: row bounds (input)
  uint32_t dd;       // [64]: |d|^2 of the box constraints [0, 32) and of the rows [32, 64)
  // reduced (null-space) front, RED instantiation only: first row of C of the foot whose limb columns are
  // [6 + 3 j, 9 + 3 j), one byte per j; bit mask of the twelve foot rows
  uint32_t red_rows, feet_mask;
  uint32_t red_blk;  // 2 bits per foot task t: j of its limb columns [6 + 3 j, 9 + 3 j)
  uint32_t red_other; // one byte per spare lane NV + m: the m-th row of C that is not a foot row (at most four)
  uint32_t b;        // [36] targets of the Cartesian task rows
  uint32_t skip_act; // non-zero: the caller does not want the active-set bit masks (QpResult::act_box / act_rows stay 0)
};

#ifndef WBC_NEWTON_POLISH
#define WBC_NEWTON_POLISH 0    // second (polishing) Newton step of fast_rsqrt / fast_rcp: off.  The cubic first step already
                               // takes the ~2^-22 hardware seed below 2^-60; the polish only trims the last 2-3 ulp of
                               // rounding noise, and it sits on the critical path of every Cholesky column (measured: +2 %
                               // throughput without it, all parity tests unchanged incl. iteration counts / active sets)
#endif
// 1/sqrt(x) and 1/x for normal positive / non-zero finite x: hardware seed + Newton steps, branch-free
// (the library rsqrt()/division carry a slow path for denormals that costs convergence barriers and
// ~30 dependent instructions in the middle of every iteration).  Relative error a few ulp (<= ~2 ulp with the polish).
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);                      // 1 - x y^2
  y = fma(y * fma(e, 0.375, 0.5), e, y);               // y (1 + e/2 + 3 e^2/8)
#if WBC_NEWTON_POLISH
  e = fma(-x * y, y, 1.0);
  y = fma(y * 0.5, e, y);
#endif
  return y;
}
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, fma(e, e, e), y);                         // y (1 + e + e^2)
#if WBC_NEWTON_POLISH
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
#endif
  return y;
}

#define WBC_IX(j) ((j) < NQ ? (j) : 0)
#define WBC_DX(j) (((j) >= 0 && (j) < ND) ? (j) : 0)

// the calling lane publishes N register values as a[0..N) at shared address `a0` (zero padded to an even count)
template <int V> struct QpIntC { static constexpr int value = V; };
template <int N, int START = 0>
__device__ __forceinline__ void publish_row(uint32_t a0, const double (&a)[N]) {
#pragma unroll
  for (int p = START / 2; p < (N + 1) / 2; ++p)
    sts_f64x2(a0 + 16 * p, a[2 * p], (2 * p + 1 < N) ? a[(2 * p + 1 < N) ? 2 * p + 1 : 0] : 0.0);
}

// SYNC: the caller's warps run in lockstep groups (fused kernel): re-align between the straight-line phases.
// NF:   the last NF variables are known at compile time to be fixed (lb == ub, e.g. the locked gripper DoFs of
//       velDamperJointConstraints, Robot_Wrapper4.py:627-631): they never enter the factorisation, the solver runs
//       on NQ = NV - NF variables.  Contract: h[k] = H[lane][k] for lane < NQ and 0 for lane >= NQ; lb_in == ub_in
//       on lanes [NQ, NV).  They are reported like the run-time fixed variables (equality-active, one working-set
//       change each).
// Out-of-line copy of the general solver: the per-state fallback of the reduced front (wbc_qp_red.inc).  Kept behind a
// call so that the two fronts do not share one register allocation (inlined side by side they spill).
template <int NV, bool SPLIT, int NF>
__device__ __noinline__ QpResult warp_qp_solve_reg_cold(const QpRegShared S, const double* a_in, const double aj,
                                                        const double bj, const int nC, const double lb_in,
                                                        const double ub_in, const int max_iter, double* x_out);

#ifndef WBC_QP_KEQ
#define WBC_QP_KEQ 12      // equality rows with a dedicated straight-line block each (static position in the working set)
#endif
// RED: the reduced front of wbc_qp_red.inc; h / hdiag / g are not read then, the problem arrives as the task-row columns
// `ta` (lane l: column l of the 36 weighted Cartesian rows), S.b and the joint task (aj, bj).
template <int NV, bool SPLIT, bool SYNC, int NF, bool RED>
__device__ __forceinline__ QpResult warp_qp_solve_reg_impl(const QpRegShared S, double (&h)[NV], const double hdiag,
                                                           const int nC, double g, const double lb_in, const double ub_in,
                                                           const int max_iter, double& x_out, const double (&ta)[36],
                                                           const double aj, const double bj) {
  constexpr int NQ = NV - NF;                // variables of the factorisation
  constexpr int n = NQ;
  constexpr int LD = NV | 1;                 // leading dimension of C (caller's layout) and of R
  constexpr int LC = (NQ + 1) & ~1;          // column stride of the stored L (even: 128-bit broadcast loads)
  constexpr int NP = (NQ + 1) / 2;           // pairs per vector
  constexpr int HALF = SPLIT ? ((((NQ + 1) / 2) + 1) & ~1) : NQ;   // elements of a row's d vector per lane
  constexpr int ND = HALF;
  constexpr int KEQ = (WBC_QP_KEQ < HALF ? WBC_QP_KEQ : HALF) < NQ ? (WBC_QP_KEQ < HALF ? WBC_QP_KEQ : HALF) : NQ - 1;
  static_assert(NQ < 32 && NF >= 0 && NQ >= 2, "lane NQ carries L^-1 g");
  static_assert(!SPLIT || (2 * HALF <= 32 && NQ - HALF <= HALF), "split layout");
  const int lane = threadIdx.x & 31;
  const bool act = lane < NQ;
  const bool sfix = NF > 0 && lane >= NQ && lane < NV;   // statically fixed variable
  const bool upper = SPLIT && lane >= 16;    // this lane holds the second half of its row's d vector
  const int crow = SPLIT ? (lane & 15) : lane;
  const bool has_row = crow < nC;
  const uint32_t R_a = S.R;                  // also the columns of L
  const uint32_t vd_a = S.vd;
  const uint32_t rk_a = S.col;
  const uint32_t doff = upper ? 8u * HALF : 0u;   // byte offset of this lane's segment inside a broadcast vector
  QpResult res;
  res.status = 0;
  res.iters = NF;

  double hd = (lane < NV) ? hdiag : 0.0;
  if (!RED) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hd = fmax(hd, __shfl_xor_sync(WBC_FULL_MASK, hd, o));
  }
  double piv_min = WBC_QP_PIVOT_REL * fmax(hd, 0.0);

  // state handed from the front (factorisation + equality rows) to the inequality loop
  double Jr[NQ], Dr[ND];
  const uint32_t lo_a = rk_a, up_a = rk_a + 8 * 32;     // per-lane bounds (once 1 / L_kk is dead)
  double x = 0.0, ax = 0.0;
  int iq = 0, p_eq = 0;
  int ws_c = -1, slot = lane;                // per working-set position (lane = position)
  double u = 0.0, rinv = 0.0;
  int bstat = 0, cstat = 0;                  // 0 none 1 lower 2 upper 3 eq
  bool fixed = false;
  // RED: the (at most four) rows of C that are not foot rows keep their d vectors as extra "rows of J" in the spare lanes
  // [NV, NV + 4): row c sits in lane hl_mine of lane c (-1: none), so the second array Dr and its passes vanish
  int hl_mine = -1;
  unsigned other_rows = 0;
#include "wbc_qp_red.inc"
  if (!RED) {
  if (NF > 0) {                              // g_i += sum_k H[i][k] x_k over the statically fixed variables
#pragma unroll
    for (int k = NQ; k < NV; ++k) {
      const double xk = __shfl_sync(WBC_FULL_MASK, lb_in, k);
      g = fma(h[k], xk, g);
    }
  }
  // ---- fixed variables (lb == ub) are eliminated: row/column of H -> identity, g shifted -------------
  // Per-lane bounds live in shared memory during the iterations (lo[0..32) | up[32..64) in `col` once the
  // factorisation is done; row bounds in S.clb / S.cub): registers are the scarce resource of this kernel.
  const unsigned eqb = __ballot_sync(WBC_FULL_MASK, act && lb_in == ub_in);
  fixed = (eqb >> lane) & 1u;
  if (eqb) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      if ((eqb >> k) & 1u) {                            // warp-uniform
        const double xk = __shfl_sync(WBC_FULL_MASK, lb_in, k);
        g = fma(h[k], xk, g);
        h[k] = (lane == k) ? 1.0 : 0.0;
      } else if (fixed) {
        h[k] = 0.0;
      }
    }
    if (fixed) g = -lb_in;
    res.iters += __popc(eqb);
  }
  __syncwarp();
  sts_f64(vd_a + 8 * lane, act ? g : 0.0);              // also zeroes the padding vd[NQ..32)

  // ---- phase A: H = L L^T, right-looking; column k of L (lane i holds L[i][k]) is stored to shared memory
  //      (Lc[k][i], 16-byte aligned columns) and broadcast back for the trailing update H[i][j] -= L[i][k] L[j][k]
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    double dk = __shfl_sync(WBC_FULL_MASK, h[k], k);
    if (!(dk > piv_min)) {
      dk = piv_min > 0.0 ? piv_min : 1.0;
      res.status |= WBC_QP_NOT_PD;
    }
    const double r = fast_rsqrt(dk);
    const double lik = h[k] * r;
    // unpredicated stores: lanes >= LC write zeros (h is 0 there) into the head of column k + 1, which is written
    // after this one; r is uniform
    sts_f64(R_a + 8 * (k * LC + lane), lik);
    sts_f64(rk_a + 8 * k, r);
    __syncwarp();
#pragma unroll
    for (int p = (k + 1) / 2; p < NP; ++p) {
      const double2 l2 = lds_f64x2(R_a + 8 * (k * LC + 2 * p));
      if (2 * p > k) h[WBC_IX(2 * p)] = fma(-lik, l2.x, h[WBC_IX(2 * p)]);
      if (2 * p + 1 < NQ) h[WBC_IX(2 * p + 1)] = fma(-lik, l2.y, h[WBC_IX(2 * p + 1)]);
    }
  }

  if (SYNC) phase_sync<true>();
  // ---- phase B: forward substitutions L^-1 [I | g | C^T], column-oriented, all right-hand sides at once:
  //      lane i < NQ: e_i (-> row i of J = L^-T), lane NQ: g, second array: the rows of C with the fixed
  //      variables' coefficients moved into `shift`
  double shift = 0.0;                                   // sum_k C[c][k] x_k over the fixed variables
  {
    const uint32_t crow_a = S.C + 8 * ((has_row ? crow : 0) * LD) + doff;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const double gj = lds_f64(vd_a + 8 * j);
      Jr[j] = (lane == NQ) ? gj : ((lane == j) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int jl = 0; jl < ND; ++jl) {
      // (a lower lane reads C[c][jl], an upper lane C[c][HALF + jl]; reads past the row are discarded)
      const double cj = lds_f64(crow_a + 8 * jl);
      Dr[jl] = (has_row && (!upper || jl + HALF < NQ)) ? cj : 0.0;
    }
    if (NF > 0) {                                       // statically fixed variables: only the lower lane adds
#pragma unroll
      for (int k = NQ; k < NQ + NF; ++k) {
        const double xk = __shfl_sync(WBC_FULL_MASK, lb_in, k);
        const double ck = lds_f64(S.C + 8 * ((has_row ? crow : 0) * LD + k));
        if (has_row && !upper) shift = fma(ck, xk, shift);
      }
    }
    if (eqb) {
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        if ((eqb >> k) & 1u) {
          const double xk = __shfl_sync(WBC_FULL_MASK, lb_in, k);
          const bool holder = !SPLIT || (upper == (k >= HALF));
          const int kl = (SPLIT && k >= HALF) ? k - HALF : k;
          if (holder) {
            shift = fma(Dr[WBC_DX(kl)], xk, shift);
            Dr[WBC_DX(kl)] = 0.0;
          }
        }
      }
    }
    if (SPLIT && (NF > 0 || eqb)) shift += __shfl_xor_sync(WBC_FULL_MASK, shift, 16);
  }
  {
    const uint32_t Lb = R_a + doff;                     // this lane's segment of a column of L
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const double r = lds_f64(rk_a + 8 * k);
      const double yJ = Jr[k] * r;
      Jr[k] = yJ;
      const int kl = (SPLIT && k >= HALF) ? k - HALF : k;
      double yD = Dr[WBC_DX(kl)] * r;
      if (SPLIT) {
        const bool holder = upper == (k >= HALF);
        if (holder) Dr[WBC_DX(kl)] = yD;
        yD = __shfl_sync(WBC_FULL_MASK, yD, crow + (k >= HALF ? 16 : 0));
      } else {
        Dr[WBC_DX(kl)] = yD;
      }
#pragma unroll
      for (int p = (k + 1) / 2; p < NP; ++p) {
        const double2 l2 = lds_f64x2(R_a + 8 * (k * LC + 2 * p));
        if (2 * p > k) Jr[WBC_IX(2 * p)] = fma(-l2.x, yJ, Jr[WBC_IX(2 * p)]);
        if (2 * p + 1 < NQ) Jr[WBC_IX(2 * p + 1)] = fma(-l2.y, yJ, Jr[WBC_IX(2 * p + 1)]);
      }
      // second array: element jl of a lower lane is column jl, of an upper lane column HALF + jl; elements that
      // only one half updates use a multiplier that is zero in the other half (no per-element selects)
      const double yD_lo = upper ? 0.0 : yD, yD_up = upper ? yD : 0.0;
#pragma unroll
      for (int pl = 0; pl < ND / 2 + (ND & 1); ++pl) {
        const int j0 = 2 * pl, j1 = 2 * pl + 1;
        const bool lv0 = j0 > k, lv1 = j1 > k && j1 < ND;
        const bool uv0 = SPLIT && (j0 + HALF > k) && (j0 + HALF < NQ), uv1 = SPLIT && j1 < ND && (j1 + HALF > k) && (j1 + HALF < NQ);
        if (!(lv0 || lv1 || uv0 || uv1)) continue;
        const double2 l2 = lds_f64x2(Lb + 8 * (k * LC + 2 * pl));
        if (lv0 || uv0) Dr[WBC_DX(j0)] = fma(-l2.x, (lv0 && (uv0 || !SPLIT)) ? yD : (lv0 ? yD_lo : yD_up), Dr[WBC_DX(j0)]);
        if (lv1 || uv1) Dr[WBC_DX(j1)] = fma(-l2.y, (lv1 && (uv1 || !SPLIT)) ? yD : (lv1 ? yD_lo : yD_up), Dr[WBC_DX(j1)]);
      }
    }
  }

  if (SYNC) phase_sync<true>();
  // ---- |d|^2 of every constraint (invariant under the orthogonal updates), x0 = -J (L^-1 g), C x0 -----
  {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (j & 1) a1 = fma(Jr[j], Jr[j], a1);
      else a0 = fma(Jr[j], Jr[j], a0);
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      if (j & 1) b1 = fma(Dr[j], Dr[j], b1);
      else b0 = fma(Dr[j], Dr[j], b0);
    }
    double ddD = b0 + b1;
    if (SPLIT) ddD += __shfl_xor_sync(WBC_FULL_MASK, ddD, 16);
    __syncwarp();
    sts_f64(S.dd + 8 * lane, a0 + a1);
    sts_f64_if(!SPLIT || lane < 16, S.dd + 8 * (32 + lane), ddD);
    sts_f64(lo_a + 8 * lane, lb_in);
    sts_f64(up_a + 8 * lane, ub_in);
    if (lane == NQ) publish_row<NQ>(vd_a, Jr);
    __syncwarp();
    double x0 = 0.0, x1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const double2 w2 = lds_f64x2(vd_a + 16 * p);
      x0 = fma(-Jr[2 * p], w2.x, x0);
      if (2 * p + 1 < NQ) x1 = fma(-Jr[WBC_IX(2 * p + 1)], w2.y, x1);
    }
#pragma unroll
    for (int p = 0; p < (ND + 1) / 2; ++p) {
      const double2 w2 = lds_f64x2(vd_a + doff + 16 * p);
      c0 = fma(-Dr[2 * p], w2.x, c0);
      if (2 * p + 1 < ND) c1 = fma(-Dr[WBC_DX(2 * p + 1)], w2.y, c1);
    }
    x = (fixed || sfix) ? lb_in : (x0 + x1);
    double cx = c0 + c1;
    if (SPLIT) cx += __shfl_xor_sync(WBC_FULL_MASK, cx, 16);
    ax = cx + shift;
    __syncwarp();
  }

  // ---- working set ----------------------------------------------------------------------------------
  bstat = (fixed || sfix) ? 3 : 0;
  unsigned eq_mask_row = __ballot_sync(WBC_FULL_MASK, lane < nC && lds_f64(S.clb + 8 * lane) == lds_f64(S.cub + 8 * lane));

  // ---- equality rows first, in index order: no step-length logic, no multipliers, no column of R, no d1 ----
  // The first KEQ accepted rows each have a straight-line block of their own: the position K in the working set is a
  // compile-time constant, so the passes run over the live entries [K, NQ) only (no zero padding), column K of J
  // and of the rows' d vectors is a plain register (no switch), and column K is not updated at all: columns
  // [0, p_eq) are never read again once their reflector has been applied (the dual step and the Givens
  // rotations only touch the inequality block).
  {
    bool eq_more = eq_mask_row != 0;
#pragma unroll
    for (int K = 0; K < KEQ; ++K) {
      if (eq_more) {                                                  // warp-uniform
        bool placed = false;
#pragma unroll 1
        while (!placed && eq_mask_row) {
          const int c = __ffs(eq_mask_row) - 1;
          eq_mask_row &= eq_mask_row - 1;
          res.iters++;
          if (crow == c) {
            publish_row<ND>(vd_a + doff, Dr);                          // SPLIT: both halves write their segment
            if (SPLIT && !upper) {                                     // dead entries [0, K) read as zeros by the lower lanes
#pragma unroll
              for (int p = 0; p < K / 2; ++p) sts_f64x2(vd_a + 16 * p, 0.0, 0.0);
              if (K & 1) sts_f64(vd_a + 8 * (K - 1), 0.0);
            }
          }
          __syncwarp();
          double z0 = 0.0, z1 = 0.0, w0 = 0.0, w1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
          for (int p = K / 2; p < NP; ++p) {
            const double2 d2 = lds_f64x2(vd_a + 16 * p);
            if (2 * p >= K) {
              z0 = fma(Jr[2 * p], d2.x, z0);
              e0 = fma(d2.x, d2.x, e0);
              if (!SPLIT) w0 = fma(Dr[WBC_DX(2 * p)], d2.x, w0);
            }
            if (2 * p + 1 < NQ) {
              z1 = fma(Jr[WBC_IX(2 * p + 1)], d2.y, z1);
              e1 = fma(d2.y, d2.y, e1);
              if (!SPLIT) w1 = fma(Dr[WBC_DX(2 * p + 1)], d2.y, w1);
            }
          }
          if (SPLIT) {                                                // K <= HALF: the upper lanes use all their entries
#pragma unroll
            for (int p = 0; p < ND / 2; ++p) {
              const double2 d2 = lds_f64x2(vd_a + doff + 16 * p);
              w0 = fma(Dr[WBC_DX(2 * p)], d2.x, w0);
              w1 = fma(Dr[WBC_DX(2 * p + 1)], d2.y, w1);
            }
          }
          double w = w0 + w1;
          if (SPLIT) w += __shfl_xor_sync(WBC_FULL_MASK, w, 16);
          const double z = z0 + z1, dd2 = e0 + e1;
          const double s_c = __shfl_sync(WBC_FULL_MASK, ax, c) - lds_f64(S.clb + 8 * c);     // C_c x - bound
          if (dd2 <= WBC_QP_DEP_TOL * lds_f64(S.dd + 8 * (32 + c))) {      // redundant (or inconsistent) equality
            if (fabs(s_c) > 1e-8) res.status |= WBC_QP_INFEASIBLE;
            __syncwarp();
            continue;
          }
          const double rs = fast_rsqrt(dd2);
          const double t = -s_c * (rs * rs);
          x = fma(t, z, x);
          ax = fma(t, w, ax);
          const double d_iq = lds_f64(vd_a + 8 * K);
          const double nrm = dd2 * rs;
          const double sigma = (d_iq >= 0.0) ? nrm : -nrm;
          const double isig = (d_iq >= 0.0) ? rs : -rs;
          const double beta = isig * fast_rcp(d_iq + sigma);
          double diq;
          if (!SPLIT) diq = Dr[WBC_DX(K)];
          else {                                                      // K < HALF: column K sits in the lower lanes
            diq = upper ? 0.0 : Dr[WBC_DX(K)];
            diq += __shfl_xor_sync(WBC_FULL_MASK, diq, 16);
          }
          const double nbJ = -beta * fma(sigma, Jr[K], z);
          const double nbD = -beta * fma(sigma, diq, w);
#pragma unroll
          for (int p = (K + 1) / 2; p < NP; ++p) {                    // v_j = d_j for j > K
            const double2 v2 = lds_f64x2(vd_a + 16 * p);
            if (2 * p > K) {
              Jr[2 * p] = fma(nbJ, v2.x, Jr[2 * p]);
              if (!SPLIT) Dr[WBC_DX(2 * p)] = fma(nbD, v2.x, Dr[WBC_DX(2 * p)]);
            }
            if (2 * p + 1 < NQ) {
              Jr[WBC_IX(2 * p + 1)] = fma(nbJ, v2.y, Jr[WBC_IX(2 * p + 1)]);
              if (!SPLIT) Dr[WBC_DX(2 * p + 1)] = fma(nbD, v2.y, Dr[WBC_DX(2 * p + 1)]);
            }
          }
          if (SPLIT) {
#pragma unroll
            for (int p = 0; p < ND / 2; ++p) {
              const double2 v2 = lds_f64x2(vd_a + doff + 16 * p);
              // (entry K of a lower lane picks up nbD * d_K: column K is dead from here on)
              Dr[WBC_DX(2 * p)] = fma(nbD, v2.x, Dr[WBC_DX(2 * p)]);
              Dr[WBC_DX(2 * p + 1)] = fma(nbD, v2.y, Dr[WBC_DX(2 * p + 1)]);
            }
          }
          if (lane == c) cstat = 3;          // (no multiplier, 1 / R_KK or constraint id: equality positions are never dropped)
          placed = true;
          __syncwarp();
        }
        if (placed) iq = K + 1;
        else eq_more = false;
      }
    }
    p_eq = iq;
  }
  // any further equality rows: one rolled loop, run-time position
#pragma unroll 1
  while (eq_mask_row) {
    const int c = __ffs(eq_mask_row) - 1;
    eq_mask_row &= eq_mask_row - 1;
    res.iters++;
    if (crow == c) publish_row<ND>(vd_a + doff, Dr);                 // SPLIT: both halves write their segment
    __syncwarp();
    sts_f64_if(lane < iq, vd_a + 8 * lane, 0.0);
    __syncwarp();
    double z0 = 0.0, z1 = 0.0, w0 = 0.0, w1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const double2 d2 = lds_f64x2(vd_a + 16 * p);
      z0 = fma(Jr[2 * p], d2.x, z0);
      e0 = fma(d2.x, d2.x, e0);
      if (!SPLIT) w0 = fma(Dr[WBC_DX(2 * p)], d2.x, w0);
      if (2 * p + 1 < NQ) {
        z1 = fma(Jr[WBC_IX(2 * p + 1)], d2.y, z1);
        e1 = fma(d2.y, d2.y, e1);
        if (!SPLIT) w1 = fma(Dr[WBC_DX(2 * p + 1)], d2.y, w1);
      }
    }
    if (SPLIT) {
#pragma unroll
      for (int p = 0; p < ND / 2; ++p) {
        const double2 d2 = lds_f64x2(vd_a + doff + 16 * p);
        w0 = fma(Dr[WBC_DX(2 * p)], d2.x, w0);
        w1 = fma(Dr[WBC_DX(2 * p + 1)], d2.y, w1);
      }
    }
    double w = w0 + w1;
    if (SPLIT) w += __shfl_xor_sync(WBC_FULL_MASK, w, 16);
    const double z = z0 + z1, dd2 = e0 + e1;
    const double s_c = __shfl_sync(WBC_FULL_MASK, ax, c) - lds_f64(S.clb + 8 * c);     // C_c x - bound
    if (dd2 <= WBC_QP_DEP_TOL * lds_f64(S.dd + 8 * (32 + c))) {      // redundant (or inconsistent) equality
      if (fabs(s_c) > 1e-8) res.status |= WBC_QP_INFEASIBLE;
      __syncwarp();
      continue;
    }
    const double rs = fast_rsqrt(dd2);
    const double t = -s_c * (rs * rs);
    x = fma(t, z, x);
    ax = fma(t, w, ax);
    const double d_iq = lds_f64(vd_a + 8 * iq);
    const double nrm = dd2 * rs;
    const double sigma = (d_iq >= 0.0) ? nrm : -nrm;
    const double isig = (d_iq >= 0.0) ? rs : -rs;
    const double v_iq = d_iq + sigma;
    const double beta = isig * fast_rcp(v_iq);
    double jiq = 0.0, diq = 0.0;
    switch (iq) {
#define WBC_PK(j) case (j): if ((j) < NQ) { jiq = Jr[WBC_IX(j)]; \
        if (!SPLIT) diq = Dr[WBC_DX(j)]; \
        else if ((j) < HALF) diq = upper ? 0.0 : Dr[WBC_DX(j)]; \
        else diq = upper ? Dr[WBC_DX((j) - HALF)] : 0.0; } break;
      WBC_REP32_ASC(WBC_PK)
#undef WBC_PK
      default: break;
    }
    if (SPLIT) diq += __shfl_xor_sync(WBC_FULL_MASK, diq, 16);
    const double nbJ = -beta * fma(sigma, jiq, z);
    const double nbD = -beta * fma(sigma, diq, w);
    sts_f64_if(lane == 0, vd_a + 8 * iq, v_iq);
    __syncwarp();
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const double2 v2 = lds_f64x2(vd_a + 16 * p);
      Jr[2 * p] = fma(nbJ, v2.x, Jr[2 * p]);
      if (!SPLIT) Dr[WBC_DX(2 * p)] = fma(nbD, v2.x, Dr[WBC_DX(2 * p)]);
      if (2 * p + 1 < NQ) {
        Jr[WBC_IX(2 * p + 1)] = fma(nbJ, v2.y, Jr[WBC_IX(2 * p + 1)]);
        if (!SPLIT) Dr[WBC_DX(2 * p + 1)] = fma(nbD, v2.y, Dr[WBC_DX(2 * p + 1)]);
      }
    }
    if (SPLIT) {
#pragma unroll
      for (int p = 0; p < ND / 2; ++p) {
        const double2 v2 = lds_f64x2(vd_a + doff + 16 * p);
        Dr[WBC_DX(2 * p)] = fma(nbD, v2.x, Dr[WBC_DX(2 * p)]);
        Dr[WBC_DX(2 * p + 1)] = fma(nbD, v2.y, Dr[WBC_DX(2 * p + 1)]);
      }
    }
    if (lane == iq) { rinv = -isig; ws_c = n + c; u = t; }
    if (lane == c) cstat = 3;
    iq++;
    p_eq = iq;
    __syncwarp();
  }
  }   // !RED

  // One flat loop: every pass is one step of the method for the current candidate (pick one if there is none).
  // (inequalities only: the equalities are in the working set by now, so is_eq is a compile-time false here)
  constexpr bool is_eq = false;
  bool have = false, is_box = false;
  int ip = 0, side = -1, owner = 0;
  double sgn = 1.0, dd = 0.0, u_new = 0.0;
  // The loop body is instantiated twice: P0 is a compile-time lower bound of the live columns of J and of the rows' d
  // vectors.  With the usual KEQ equality rows in the working set the passes run over columns [KEQ, NQ) only.
  auto ineq_loop = [&](auto P0c) {
  constexpr int P0 = decltype(P0c)::value;
#pragma unroll 1
  for (;;) {
    if (!have) {
      // ------------------------------------------------------------ pick the entering constraint
      {
        // most violated side over the boxes (index lane) and the rows (index n + lane), ties (within the window of
        // wbc_qp.cuh) to the lowest index: a min-reduction of the value, then two votes (every box index is below every
        // row index)
        double vb = INFINITY, vc = INFINITY;
        int myside_b = -1, myside_c = -1;
        if (act && bstat == 0) {
          const double slo = x - lds_f64(lo_a + 8 * lane), sup = lds_f64(up_a + 8 * lane) - x;
          vb = fmin(slo, sup);
          myside_b = (slo <= sup) ? -1 : +1;
        }
        if (lane < nC && cstat == 0) {
          const double slo = ax - lds_f64(S.clb + 8 * lane), sup = lds_f64(S.cub + 8 * lane) - ax;
          vc = fmin(slo, sup);
          myside_c = (slo <= sup) ? -1 : +1;
        }
        double best = fmin(vb, vc);
        if (!__any_sync(WBC_FULL_MASK, best < -WBC_QP_FEAS_TOL)) break;   // primal feasible: optimal (the usual exit)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(WBC_FULL_MASK, best, o));
        const double thr = best + (WBC_QP_TIE_ABS + WBC_QP_TIE_REL * fabs(best));     // tie window (wbc_qp.cuh)
        const unsigned wb = __ballot_sync(WBC_FULL_MASK, vb <= thr), wc = __ballot_sync(WBC_FULL_MASK, vc <= thr);
        ip = wb ? __ffs(wb) - 1 : n + __ffs(wc) - 1;
        const int src = (ip < n) ? ip : ip - n;
        side = __shfl_sync(WBC_FULL_MASK, (ip < n) ? myside_b : myside_c, src);
      }
      sgn = (side > 0) ? -1.0 : 1.0;                               // normal = sgn * a_ip
      is_box = ip < n;
      owner = is_box ? ip : ip - n;
      dd = lds_f64(S.dd + 8 * (is_box ? owner : 32 + owner));
      u_new = 0.0;
      have = true;
    }
    if (res.iters >= max_iter) { res.status |= WBC_QP_MAXITER; break; }
    res.iters++;
    // broadcast the (unsigned) d vector of the entering constraint; keep d_lane, then zero the first iq entries
    if (is_box) {
      if (lane == owner) publish_row<NQ, P0>(vd_a, Jr);
    } else if (RED) {
      if (lane == NV + __popc(other_rows & ((1u << owner) - 1u))) publish_row<NQ, P0>(vd_a, Jr);
    } else {
      if (crow == owner) publish_row<ND>(vd_a + doff, Dr);         // SPLIT: both halves write their segment
    }
    __syncwarp();
    const double d_own = lds_f64(vd_a + 8 * lane);
    __syncwarp();
    sts_f64_if(lane < iq, vd_a + 8 * lane, 0.0);
    __syncwarp();
    // z = J2 d2, w = D2 d2 (= C z), dd2 = |d2|^2
    double z0 = 0.0, z1 = 0.0, w0 = 0.0, w1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int p = P0 / 2; p < NP; ++p) {
      const double2 d2 = lds_f64x2(vd_a + 16 * p);
      if (2 * p >= P0) {
        z0 = fma(Jr[2 * p], d2.x, z0);
        e0 = fma(d2.x, d2.x, e0);
        if (!SPLIT) w0 = fma(Dr[WBC_DX(2 * p)], d2.x, w0);
      }
      if (2 * p + 1 < NQ) {
        z1 = fma(Jr[WBC_IX(2 * p + 1)], d2.y, z1);
        e1 = fma(d2.y, d2.y, e1);
        if (!SPLIT) w1 = fma(Dr[WBC_DX(2 * p + 1)], d2.y, w1);
      }
    }
    if (SPLIT && !RED) {
#pragma unroll
      for (int p = 0; p < ND / 2; ++p) {
        const double2 d2 = lds_f64x2(vd_a + doff + 16 * p);
        w0 = fma(Dr[WBC_DX(2 * p)], d2.x, w0);
        w1 = fma(Dr[WBC_DX(2 * p + 1)], d2.y, w1);
      }
    }
    double w = w0 + w1;
    if (SPLIT && !RED) w += __shfl_xor_sync(WBC_FULL_MASK, w, 16);
    const double z = z0 + z1, dd2 = e0 + e1;
    if (RED) {                                                     // C_c z is the z of the lane that carries row c
      const double zr = __shfl_sync(WBC_FULL_MASK, z, hl_mine >= 0 ? hl_mine : 0);
      w = (hl_mine >= 0) ? zr : 0.0;
    }
    // r = R^-1 d1 on the inequality block [p_eq, iq)
    double rr = (lane < iq) ? sgn * d_own : 0.0;
#pragma unroll 1
    for (int k = iq - 1; k >= p_eq; --k) {
      const int slot_k = __shfl_sync(WBC_FULL_MASK, slot, k);
      const double rk = __shfl_sync(WBC_FULL_MASK, rr * rinv, k);
      if (lane == k) rr = rk;
      else if (lane < k && lane >= p_eq) rr -= lds_f64(R_a + 8 * (lane * LD + slot_k)) * rk;
    }
    // constraint value at x:  s = n.x - bnd  (negative when violated)
    double s_ip;
    {
      const double v_i = __shfl_sync(WBC_FULL_MASK, is_box ? x : ax, owner);
      const double lo_i = lds_f64((is_box ? lo_a : S.clb) + 8 * owner);
      const double up_i = lds_f64((is_box ? up_a : S.cub) + 8 * owner);
      s_ip = (side > 0) ? (up_i - v_i) : (v_i - lo_i);
    }
    const bool dependent = dd2 <= WBC_QP_DEP_TOL * dd;
    // 1 / |d2|, shared by the step length and the reflector (dd2 > 0 unless dependent: then unused)
    const double rs = fast_rsqrt(fmax(dd2, 1e-300));
    const double t2 = dependent ? INFINITY : -s_ip * (rs * rs);
    bool add = false;

    if (is_eq) {
      if (dependent) {                                               // redundant (or inconsistent) equality
        if (fabs(s_ip) > 1e-8) res.status |= WBC_QP_INFEASIBLE;
        have = false;
      } else {
        x = fma(t2 * sgn, z, x);
        ax = fma(t2 * sgn, w, ax);
        u_new = t2;
        add = true;
      }
    } else {
      // dual step length over active inequalities
      double t1 = INFINITY;
      int l = 0x7fffffff;
      if (iq > p_eq) {                                               // (no active inequality: nothing can block the step)
        const bool blk = lane >= p_eq && lane < iq && rr > 0.0;
        double mine = blk ? u / rr : INFINITY;
        t1 = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t1 = fmin(t1, __shfl_xor_sync(WBC_FULL_MASK, t1, o));
        const unsigned wl = __ballot_sync(WBC_FULL_MASK, blk && mine == t1);        // ties to the earliest position
        l = wl ? __ffs(wl) - 1 : 0x7fffffff;
      }
      const double t = fmin(t1, t2);
      if (!(t < INFINITY)) { res.status |= WBC_QP_INFEASIBLE; break; }
      if (lane >= p_eq && lane < iq) u -= t * rr;
      u_new += t;
      if (!dependent) {
        x = fma(t * sgn, z, x);
        ax = fma(t * sgn, w, ax);
      }
      if (dependent || !(t2 <= t1)) {
        // ------------------------------------------------------ drop working-set position l, keep the candidate
        const int c_drop = __shfl_sync(WBC_FULL_MASK, ws_c, l);
#pragma unroll 1
        for (int k = l; k < iq - 1; ++k) {
          const int slot_k1 = __shfl_sync(WBC_FULL_MASK, slot, k + 1);
          const double a = lds_f64(R_a + 8 * (k * LD + slot_k1)), b = lds_f64(R_a + 8 * ((k + 1) * LD + slot_k1));
          const double rho = sqrt(a * a + b * b);
          const double cg = (rho > 0.0) ? a / rho : 1.0, sg = (rho > 0.0) ? b / rho : 0.0;
          __syncwarp();                                               // everyone has read a, b before rows k, k+1 change
          if (lane > k && lane < iq) {
            const uint32_t a0 = R_a + 8 * (k * LD + slot), a1 = R_a + 8 * ((k + 1) * LD + slot);
            const double r0 = lds_f64(a0), r1 = lds_f64(a1);
            sts_f64(a0, cg * r0 + sg * r1);
            sts_f64(a1, -sg * r0 + cg * r1);
          }
          // rotate columns k, k + 1 of J and of the rows' d vectors
          switch (k) {
#define WBC_GV(j) case (j): if ((j) >= P0 && (j) + 1 < NQ) { \
              const double j0 = Jr[WBC_IX(j)], j1 = Jr[WBC_IX((j) + 1)]; \
              Jr[WBC_IX(j)] = cg * j0 + sg * j1; Jr[WBC_IX((j) + 1)] = -sg * j0 + cg * j1; \
              if (RED) { \
              } else if (!SPLIT || (j) + 1 < HALF) { \
                if (!upper) { const double d0 = Dr[WBC_DX(j)], d1 = Dr[WBC_DX((j) + 1)]; \
                  Dr[WBC_DX(j)] = cg * d0 + sg * d1; Dr[WBC_DX((j) + 1)] = -sg * d0 + cg * d1; } \
              } else if ((j) >= HALF) { \
                if (upper) { const double d0 = Dr[WBC_DX((j) - HALF)], d1 = Dr[WBC_DX((j) + 1 - HALF)]; \
                  Dr[WBC_DX((j) - HALF)] = cg * d0 + sg * d1; Dr[WBC_DX((j) + 1 - HALF)] = -sg * d0 + cg * d1; } \
              } else {   /* j == HALF - 1: column j sits in the lower lane, column j + 1 in the upper lane */ \
                const double mine = upper ? Dr[0] : Dr[WBC_DX(HALF - 1)]; \
                const double other = __shfl_xor_sync(WBC_FULL_MASK, mine, 16); \
                if (upper) Dr[0] = -sg * other + cg * mine; else Dr[WBC_DX(HALF - 1)] = cg * mine + sg * other; \
              } } break;
            WBC_REP32_ASC(WBC_GV)
#undef WBC_GV
            default: break;
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
          if (lane >= l && lane < iq - 1) rinv = 1.0 / lds_f64(R_a + 8 * (lane * LD + slot));
          if (c_drop < n) { if (lane == c_drop) bstat = 0; }
          else if (lane == c_drop - n) cstat = 0;
          iq--;
        }
      } else {
        add = true;
      }
    }
    if (add) {
      // -------------------------------------------------------- full step: constraint ip enters at position iq
      const double d_iq = lds_f64(vd_a + 8 * iq);                     // unsigned (only entries below iq were zeroed)
      const double nrm = dd2 * rs;
      const double sigma = (d_iq >= 0.0) ? nrm : -nrm;
      const double isig = (d_iq >= 0.0) ? rs : -rs;                   // 1 / sigma
      const double v_iq = d_iq + sigma;
      const double beta = isig * fast_rcp(v_iq);                      // 1 / (sigma v_iq)
      double jiq = 0.0, diq = 0.0;
      switch (iq) {                                                   // column iq of J and of the rows' d vectors
#define WBC_PK(j) case (j): if ((j) >= P0 && (j) < NQ) { jiq = Jr[WBC_IX(j)]; \
          if (RED) { } \
          else if (!SPLIT) diq = Dr[WBC_DX(j)]; \
          else if ((j) < HALF) diq = upper ? 0.0 : Dr[WBC_DX(j)]; \
          else diq = upper ? Dr[WBC_DX((j) - HALF)] : 0.0; } break;
        WBC_REP32_ASC(WBC_PK)
#undef WBC_PK
        default: break;
      }
      if (SPLIT && !RED) diq += __shfl_xor_sync(WBC_FULL_MASK, diq, 16);
      const double nbJ = -beta * fma(sigma, jiq, z);
      const double nbD = -beta * fma(sigma, diq, w);
      sts_f64_if(lane == 0, vd_a + 8 * iq, v_iq);                    // vd = v = d2 + sigma e_iq (zeros below iq)
      __syncwarp();
#pragma unroll
      for (int p = P0 / 2; p < NP; ++p) {
        const double2 v2 = lds_f64x2(vd_a + 16 * p);
        if (2 * p >= P0) {
          Jr[2 * p] = fma(nbJ, v2.x, Jr[2 * p]);
          if (!SPLIT) Dr[WBC_DX(2 * p)] = fma(nbD, v2.x, Dr[WBC_DX(2 * p)]);
        }
        if (2 * p + 1 < NQ) {
          Jr[WBC_IX(2 * p + 1)] = fma(nbJ, v2.y, Jr[WBC_IX(2 * p + 1)]);
          if (!SPLIT) Dr[WBC_DX(2 * p + 1)] = fma(nbD, v2.y, Dr[WBC_DX(2 * p + 1)]);
        }
      }
      if (SPLIT && !RED) {
#pragma unroll
        for (int p = 0; p < ND / 2; ++p) {
          const double2 v2 = lds_f64x2(vd_a + doff + 16 * p);
          Dr[WBC_DX(2 * p)] = fma(nbD, v2.x, Dr[WBC_DX(2 * p)]);
          Dr[WBC_DX(2 * p + 1)] = fma(nbD, v2.y, Dr[WBC_DX(2 * p + 1)]);
        }
      }
      if (!is_eq) {                                                   // R column (signed): [d1 ; -sigma]
        const int slot_new = __shfl_sync(WBC_FULL_MASK, slot, iq);
        sts_f64_if(lane < iq, R_a + 8 * (lane * LD + slot_new), sgn * d_own);
        sts_f64_if(lane == iq, R_a + 8 * (iq * LD + slot_new), -sgn * sigma);
      }
      if (lane == iq) {
        rinv = -sgn * isig;                                           // 1 / R_iq,iq
        ws_c = ip;
        u = u_new;
      }
      const int st = is_eq ? 3 : (side > 0 ? 2 : 1);
      if (is_box) { if (lane == ip) bstat = st; }
      else if (lane == owner) cstat = st;
      iq++;
      if (is_eq) p_eq = iq;
      have = false;
    }
    __syncwarp();
  }
  };
  if (KEQ >= 2 && p_eq == KEQ) ineq_loop(QpIntC<(KEQ & ~1)>{});
  else ineq_loop(QpIntC<0>{});

  x_out = (fixed || sfix) ? lds_f64(lo_a + 8 * lane) : x;
  res.act_box = res.act_rows = 0ull;
  if (!S.skip_act) pack_active_sets(lane, NQ + NF, nC, bstat, cstat, res);   // (~130 instructions: only on request)
  return res;
}

template <int NV, bool SPLIT, bool SYNC = false, int NF = 0>
__device__ __forceinline__ QpResult warp_qp_solve_reg(const QpRegShared S, double (&h)[NV], const double hdiag,
                                                      const int nC, double g, const double lb_in, const double ub_in,
                                                      const int max_iter, double& x_out) {
  const double ta[36] = {};
  return warp_qp_solve_reg_impl<NV, SPLIT, SYNC, NF, false>(S, h, hdiag, nC, g, lb_in, ub_in, max_iter, x_out, ta, 0.0, 0.0);
}

// Fallback of the reduced front: H = A^T A + aj^2 I and g = -A^T b - aj b_joint assembled column by column from the
// task-row columns (compact, not fast: this path runs for states that fail the preconditions of wbc_qp_red.inc), then
// the general solver.
template <int NV, bool SPLIT, int NF>
__device__ __noinline__ QpResult warp_qp_solve_reg_cold(const QpRegShared S, const double* a_in, const double aj,
                                                        const double bj, const int nC, const double lb_in,
                                                        const double ub_in, const int max_iter, double* x_out) {
  constexpr int NQ = NV - NF;
  const int lane = threadIdx.x & 31;
  double a[36];
#pragma unroll
  for (int r = 0; r < 36; ++r) a[r] = a_in[r];
  double g0 = 0.0, g1 = 0.0;
#pragma unroll
  for (int r = 0; r < 18; ++r) {
    const double2 b2 = lds_f64x2(S.b + 16 * r);
    g0 = fma(-a[2 * r], b2.x, g0);
    g1 = fma(-a[2 * r + 1], b2.y, g1);
  }
  const double g = (g0 + g1) - aj * bj;
  double h[NV];
  double hdiag = 0.0;
  const uint32_t buf = S.col;                // 64 doubles, free until the factorisation
#pragma unroll
  for (int l = 0; l < NV; ++l) {
    if (lane == l) {
#pragma unroll
      for (int r = 0; r < 18; ++r) sts_f64x2(buf + 16 * r, a[2 * r], a[2 * r + 1]);
    }
    __syncwarp();
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int r = 0; r < 18; ++r) {
      const double2 c2 = lds_f64x2(buf + 16 * r);
      s0 = fma(a[2 * r], c2.x, s0);
      s1 = fma(a[2 * r + 1], c2.y, s1);
    }
    const double hl = (s0 + s1) + ((l == lane) ? aj * aj : 0.0);
    if (l == lane) hdiag = hl;
    h[l] = (lane < NQ) ? hl : 0.0;
    __syncwarp();
  }
  if (lane >= NV) hdiag = 0.0;
  double x;
  const QpResult res = warp_qp_solve_reg<NV, SPLIT, false, NF>(S, h, hdiag, nC, g, lb_in, ub_in, max_iter, x);
  *x_out = x;
  return res;
}
