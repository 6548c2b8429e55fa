// wbc_step.cuh -- the fused WBC tick for one robot state per warp:
//   FK + joint-Jacobian columns + frame placements            (updateState,        Robot_Wrapper4.py:387-428)
//   weighted task rows A and targets b                        (qpA / qpb,          :1271-1294, 474-490, 948-1157)
//   H = A^T A, g = -A^T b kept in shared memory               (QP.__init__,        QP_Wrapper.py:17-18)
//   velocity-damper box bounds                                (velDamperJointConstraints, :572-637)
//   constraint rows C, Clb, Cub                               (findConstraints,    :764-836)
//   dual active-set QP                                        (QP.solveQP,         QP_Wrapper.py:23-53)
//   integrate + base estimate                                 (jointVelocitiestoConfig :440, trunkWorldPos :1297)
// A never touches HBM; per state the kernel reads q, targets, task memory and references and
// writes qdot, status, iters (+ optional memory / q_next).
#pragma once
#include "wbc_qp_reg.cuh"

#define WBC_IO_TARGETS 0
#define WBC_IO_MEM 18
#define WBC_IO_REF 90
#define WBC_IO_TOTAL 114
#define WBC_HOT_FRAMES 6

// offsets inside mem / ref blocks
#define MEM_PREV_EE_POS 0
#define MEM_PREV_EE_ROT 15
#define MEM_PREV_TRUNK_REF 60
#define MEM_OLD_TRUNK_ROT 63
#define REF_DEF_EE_ORI 0
#define REF_DEF_TRUNK_ORI 15
#define REF_INIT_TRUNK_POS 18
#define REF_INIT_TRUNK_EUL 21



struct StepLayout {   // per-warp shared-memory layout in doubles (every offset even: 16-byte aligned)
  int hs, ast, col, omf, vec, io, total;
};

__host__ __device__ inline int wbc_ld(int nv) { return nv | 1; }
#define WBC_LDT 38    // doubles per column of the transposed task rows: 304 B = 19 x 16 B, conflict-free 128-bit accesses

__host__ __device__ inline StepLayout step_layout(int nv, int nC) {
  StepLayout L;
  const int ld = wbc_ld(nv);
  int r0 = nv * (nv + 2);                              // H rows, later the L columns / R factor of the QP ...
  const int fk = WBC_MAX_JOINTS * WBC_T_STRIDE;        // ... aliased by the oMi scratch (dead before H is written)
  if (r0 < fk) r0 = fk;
  r0 = (r0 + 1) & ~1;
  int r1 = nv * WBC_LDT;                               // transposed task rows AsT, later the constraint rows C
  if (r1 < nC * ld) r1 = nC * ld;
  r1 = (r1 + 1) & ~1;
  L.hs = 0;
  L.ast = r0;
  L.col = r0 + r1;
  L.omf = L.col + 64;
  L.vec = L.omf + WBC_HOT_FRAMES * WBC_T_STRIDE + 2;   // 6 * 13 + 2 = 80
  L.io = L.vec + 40 + 32 + 32 + 32 + 40;               // qs[40] vd[32] clb[32] cub[32] bs[40]
  L.total = L.io + WBC_IO_TOTAL + 2;
  L.total = (L.total + 1) & ~1;
  return L;
}

struct StepParams {
  const DevModel* model;
  WbcConfig cfg;
  WbcStepIO io;
  WbcAssembleOut dbg;
  long long N;
  int nC, m_rows, flags;
};

// Number of rows of C and A implied by the masks.
__host__ __device__ inline int cfg_nc(const WbcConfig& c) {
  int n = 0;
  if (c.constraint_mask & WBC_CON_COM) n += 2;
  if (c.constraint_mask & WBC_CON_TRUNK) n += 4;
  for (int i = 0; i < 5; ++i)
    if (c.constraint_mask & (WBC_CON_FR << i)) n += 3;
  return n + c.n_extra_rows;
}
__host__ __device__ inline int cfg_m(const WbcConfig& c, int nv) {
  int n = 0;
  for (int t = 0; t < 6; ++t)
    if (c.task_mask & (1 << t)) n += 6;
  if (c.task_mask & WBC_TASK_JOINT) n += nv;
  return n;
}

// pin.integrate for the free-flyer part (oracle/pin.py: integrate, exp6, exp3). One lane.
__device__ __forceinline__ void integrate_freeflyer(const double* q, const double* v, double* out) {
  const double w[3] = {v[3], v[4], v[5]};
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double t = sqrt(t2);
  double a_wxv, a_v, a_w, ct;
  if (t < 1.220703125e-4) {          // sqrt(sqrt(eps))
    a_wxv = 0.5 - t2 / 24.0;
    a_v = 1.0 - t2 / 6.0;
    a_w = 1.0 / 6.0 - t2 / 120.0;
    ct = 1.0 - t2 / 2.0;
  } else {
    double st;
    sincos(t, &st, &ct);
    a_wxv = (1.0 - ct) / t2;
    a_v = st / t;
    a_w = (1.0 - a_v) / t2;
  }
  double wxv[3];
  cross3(w, v, wxv);
  const double wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
  double pe[3], Re[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) pe[i] = a_v * v[i] + a_wxv * wxv[i] + a_w * wv * w[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Re[3 * i + j] = a_wxv * w[i] * w[j];
  Re[1] -= a_v * w[2]; Re[3] += a_v * w[2];
  Re[2] += a_v * w[1]; Re[6] -= a_v * w[1];
  Re[5] -= a_v * w[0]; Re[7] += a_v * w[0];
  Re[0] += ct; Re[4] += ct; Re[8] += ct;
  double R0[9], R1[9], p1[3];
  quat_to_matrix_eigen(q[3], q[4], q[5], q[6], R0);
  mat3_mul(R0, Re, R1);
  mat3_vec(R0, pe, p1);
  out[0] = q[0] + p1[0]; out[1] = q[1] + p1[1]; out[2] = q[2] + p1[2];
  double rq[4];
  matrix_to_quat_eigen(R1, rq);
  const double dot = rq[0] * q[3] + rq[1] * q[4] + rq[2] * q[5] + rq[3] * q[6];
  if (dot < 0) { rq[0] = -rq[0]; rq[1] = -rq[1]; rq[2] = -rq[2]; rq[3] = -rq[3]; }
  const double n2 = rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3];
  const double al = (3.0 - n2) / 2.0;
  out[3] = rq[0] * al; out[4] = rq[1] * al; out[5] = rq[2] * al; out[6] = rq[3] * al;
}

// calcTargetVelEE3 (Robot_Wrapper4.py:1052-1157): lane i in 0..4.  Writes b (6) and updates the staged memory.
__device__ __forceinline__ void ee_task_target(const WbcConfig& cfg, int i, const double* __restrict__ oMf,
                                               double* __restrict__ io, double inv_dt, double* b) {
  const double* target = io + WBC_IO_TARGETS + 3 * i;
  double* prev = io + WBC_IO_MEM + MEM_PREV_EE_POS + 3 * i;
  double* prevR = io + WBC_IO_MEM + MEM_PREV_EE_ROT + 9 * i;
  const double* fk = oMf + i * WBC_T_STRIDE + 9;
  double ref_vel[3], err[3], ge[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ref_vel[k] = (target[k] - prev[k]) * inv_dt;
    err[k] = (target[k] - fk[k]) * inv_dt;
  }
  mat3_vec(cfg.ee_gain_pos[i], err, ge);
  double qref[4], Rref[9];
  scipy_quat_from_euler_xyz(io + WBC_IO_REF + REF_DEF_EE_ORI + 3 * i, qref);
  scipy_matrix_from_quat(qref, Rref);
  double dR[9], sk[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) dR[k] = (Rref[k] - prevR[k]) * inv_dt;
  mat3_mul_bt(dR, Rref, sk);                                   // (dR/dt) * Rref^T   (:1125)
  const double wgt = cfg.cart_task_weight[i];
  b[0] = (ref_vel[0] + ge[0]) * wgt;
  b[1] = (ref_vel[1] + ge[1]) * wgt;
  b[2] = (ref_vel[2] + ge[2]) * wgt;
  b[3] = sk[7] * wgt;                                          // skew[2,1]
  b[4] = sk[2] * wgt;                                          // skew[0,2]
  b[5] = sk[3] * wgt;                                          // skew[1,0]
#pragma unroll
  for (int k = 0; k < 3; ++k) prev[k] = target[k];             // :1151
#pragma unroll
  for (int k = 0; k < 9; ++k) prevR[k] = Rref[k];              // :1152
}

// calcTargetVelTrunk2 (Robot_Wrapper4.py:948-1015): one lane.
__device__ __forceinline__ void trunk_task_target(const WbcConfig& cfg, const double* __restrict__ oMf,
                                                  const double* fkq, double* __restrict__ io, double inv_dt, double* b) {
  const double* target = io + WBC_IO_TARGETS + 15;
  double* prev = io + WBC_IO_MEM + MEM_PREV_TRUNK_REF;
  double* oldR = io + WBC_IO_MEM + MEM_OLD_TRUNK_ROT;
  const double* fk = oMf + WBC_FRAME_TRUNK * WBC_T_STRIDE + 9;
  double ref_vel[3], err[3], ge[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ref_vel[k] = (target[k] - prev[k]) * inv_dt;
    err[k] = (target[k] - fk[k]) * inv_dt;
  }
  mat3_vec(cfg.trunk_gain_pos, err, ge);
  double r[4], Rref[9];
  scipy_quat_from_euler_xyz(io + WBC_IO_REF + REF_DEF_TRUNK_ORI, r);
  scipy_matrix_from_quat(r, Rref);
  const double* f = fkq;
  double qe[3];
  qe[0] = (f[3] * r[0]) - (f[0] * r[3]) + (f[1] * r[2]) - (f[2] * r[1]);
  qe[1] = (f[3] * r[1]) - (f[1] * r[3]) - (f[0] * r[2]) + (f[2] * r[0]);
  qe[2] = (f[0] * r[1]) - (f[1] * r[0]);                       // :976 -- the w*z terms cancel exactly
  double dR[9], sk[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) dR[k] = (Rref[k] - oldR[k]) * inv_dt;
  mat3_mul(dR, Rref, sk);                                      // (dR/dt) * Rref, NOT transposed (:984)
  const double wgt = cfg.cart_task_weight[5];
  b[0] = (ref_vel[0] + ge[0]) * wgt;
  b[1] = (ref_vel[1] + ge[1]) * wgt;
  b[2] = (ref_vel[2] + ge[2]) * wgt;
  b[3] = (sk[7] + cfg.trunk_gain_ori[0] * qe[0]) * wgt;
  b[4] = (sk[2] + cfg.trunk_gain_ori[1] * qe[1]) * wgt;
  b[5] = (sk[3] + cfg.trunk_gain_ori[2] * qe[2]) * wgt;
#pragma unroll
  for (int k = 0; k < 3; ++k) prev[k] = target[k];             // :995
#pragma unroll
  for (int k = 0; k < 9; ++k) oldR[k] = Rref[k];               // :996
}

// velDamperJointConstraints (Robot_Wrapper4.py:572-637) for velocity index k.
__device__ __forceinline__ void damper_bounds(const DevModel* __restrict__ M, const WbcConfig& cfg,
                                              const double* __restrict__ qs, int k, double& lbv, double& ubv) {
  const int grip = cfg.gripper_joint_id;
  const int qidx = (k < 6) ? k : k + 1;                        // np.delete(., 6)
  double lo, up;
  if (qidx < 7) { lo = -5.0; up = 5.0; }
  else if (qidx >= grip - 2 + 7) { lo = 0.0; up = 0.0; }
  else { lo = M->lower[qidx]; up = M->upper[qidx]; }
  const double vel = (k < 7) ? 5.0 : M->velocity[k];           // vel_lim[i] = 5 for i < 7 (:593)
  const double c = (cfg.compat_flags & WBC_COMPAT_DAMPER_OFF_BY_ONE) ? qs[k] : qs[qidx];   // quirk D.2
  const double coef = cfg.damper_coef, qi = cfg.damper_qi, qsv = cfg.damper_qs;
  if (c <= lo + qi) {
    lbv = -coef * (c - lo - qsv) / (qi - qsv);
    if (lbv > vel) lbv = vel;
    if (lbv < -vel) lbv = -vel;
  } else {
    lbv = -vel;
  }
  if (c >= up - qi) {
    ubv = coef * (up - c - qsv) / (qi - qsv);
    if (ubv < -vel) ubv = -vel;
    if (ubv > vel) ubv = vel;
  } else {
    ubv = vel;
  }
  if (lbv > 0) lbv = lbv * -1;
  if (ubv < 0) ubv = ubv * -1;
  if (k >= grip - 2 + 6) { lbv = 0.0; ubv = 0.0; }
}

// One WBC tick for the state `sidx`, executed by one warp.  ws: this warp's shared workspace.
// The warps of a CTA are re-aligned with a block barrier at the phase boundaries (PHASE_SYNC): the tick is
// ~10^4 mostly straight-line instructions, far more than the instruction cache holds, so warps that drift apart
// each stream the whole code through the cache on their own (ncu: 50 % of stall samples were `no_instruction`);
// in step, one fetch feeds all warps.  `valid` == false: a padding warp that shadows the last state, no writes.
template <bool ON>
__device__ __forceinline__ void phase_sync() {
  if (ON) __syncthreads();
  else __syncwarp();
}

template <int NV, bool DEBUG_OUT, bool SPLIT>
__device__ __forceinline__ void warp_wbc_step(const StepParams& P, const DevModel* __restrict__ M,
                                              double* __restrict__ ws, const StepLayout L, long long sidx,
                                              const bool valid) {
#ifndef WBC_PHASE_SYNC
#define WBC_PHASE_SYNC 1
#endif
  constexpr bool PS = WBC_PHASE_SYNC && !DEBUG_OUT;
  constexpr int LD = NV | 1;
  const int lane = threadIdx.x & 31;
  const WbcConfig& cfg = P.cfg;
  const int nq = NV + 1;
  const double dt = P.io.dt;
  const double inv_dt = 1.0 / dt;       // the reference divides by dt; multiplying by 1/dt differs by <= 1 ulp
  double* Hs = ws + L.hs;           // [NV][LD] rows of H; the QP reuses it for its R factor
  double* AsT = ws + L.ast;         // [NV][WBC_LDT] column k of the Cartesian task rows (36 values)
  double* Cs = ws + L.ast;          // [nC][LD] constraint rows (after AsT is dead)
  double* oMi = ws + L.hs;
  double* oMf = ws + L.omf;
  double* qs = ws + L.vec;          // [40]
  double* vd = qs + 40;             // [32]
  double* clbs = vd + 32;
  double* cubs = clbs + 32;
  double* bs = cubs + 32;           // [40]
  double* io = ws + L.io;

  // ---------------------------------------------------------------- stage inputs (coalesced)
  {
    const double* qg = P.io.q + sidx * nq;
    for (int i = lane; i < nq; i += 32) qs[i] = qg[i];
    const double* tg = P.io.targets + sidx * WBC_TARGETS_STRIDE;
    if (lane < WBC_TARGETS_STRIDE) io[WBC_IO_TARGETS + lane] = tg[lane];
    const double* mg = P.io.mem_in + sidx * WBC_MEM_STRIDE;
    for (int i = lane; i < WBC_MEM_STRIDE; i += 32) io[WBC_IO_MEM + i] = mg[i];
    const double* rg = P.io.ref + sidx * WBC_REF_STRIDE;
    if (lane < WBC_REF_STRIDE) io[WBC_IO_REF + lane] = rg[lane];
  }
  phase_sync<PS>();

  // ---------------------------------------------------------------- kinematics
  warp_fk(M, qs, oMi, lane);
  double Sc[6];
  warp_jac_column(M, oMi, lane, Sc);
  if (lane < WBC_HOT_FRAMES) {       // hot frames: 5 EE + trunk
    const int par = M->frame_parent[lane];
    const double* Pm = oMi + par * WBC_T_STRIDE;
    double Rp[9], R[9], p[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rp[i] = Pm[i];
    mat3_mul(Rp, M->frR[lane], R);
    mat3_vec(Rp, M->frp[lane], p);
    double* out = oMf + lane * WBC_T_STRIDE;
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    out[9] = p[0] + Pm[9]; out[10] = p[1] + Pm[10]; out[11] = p[2] + Pm[11];
  }
  // centre of mass (only when the CoM constraint is on): lane j -> m_j * c_j in the world
  double com_w[3] = {0, 0, 0};
  double Jcom[2] = {0, 0};
  if (cfg.constraint_mask & WBC_CON_COM) {
    double mc[4] = {0, 0, 0, 0};
    if (lane >= 1 && lane < M->njoints) {
      const double* T = oMi + lane * WBC_T_STRIDE;
      double R[9], c[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = T[i];
      mat3_vec(R, M->com[lane], c);
      const double m = M->mass[lane];
      mc[0] = m; mc[1] = m * (c[0] + T[9]); mc[2] = m * (c[1] + T[10]); mc[3] = m * (c[2] + T[11]);
    }
    // subtree sums for my column: sum over joints j in sub_joints[lane]
    double sm = 0, s1 = 0, s2 = 0, s3 = 0;
    const uint32_t sub = (lane < NV) ? M->sub_joints[lane] : 0u;
    for (int j = 1; j < M->njoints; ++j) {
      const double m0 = __shfl_sync(WBC_FULL_MASK, mc[0], j), m1 = __shfl_sync(WBC_FULL_MASK, mc[1], j);
      const double m2 = __shfl_sync(WBC_FULL_MASK, mc[2], j), m3 = __shfl_sync(WBC_FULL_MASK, mc[3], j);
      if ((sub >> j) & 1u) { sm += m0; s1 += m1; s2 += m2; s3 += m3; }
    }
    const double Mt = M->total_mass;
    com_w[0] = warp_sum(mc[1]) / Mt; com_w[1] = warp_sum(mc[2]) / Mt; com_w[2] = warp_sum(mc[3]) / Mt;
    // Jcom column = (sm * lin - (sum m c) x ang) / M   (only x, y rows are used, :670)
    const double msc[3] = {s1, s2, s3};
    double cx[3];
    cross3(msc, Sc + 3, cx);
    Jcom[0] = (sm * Sc[0] - cx[0]) / Mt;
    Jcom[1] = (sm * Sc[1] - cx[1]) / Mt;
  }
  phase_sync<PS>();

  // ---------------------------------------------------------------- task rows for my column (registers)
  double a[36];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const bool on = (cfg.task_mask >> t) & 1;
    double Jc[6];
    if (on) {
      // EE tasks: LOCAL_WORLD_ALIGNED (:476-480); trunk task: WORLD (:488)
      frame_jac_column(Sc, M->frame_supp[t], lane, oMf + t * WBC_T_STRIDE,
                       t < 5 ? WBC_RF_LOCAL_WORLD_ALIGNED : WBC_RF_WORLD, Jc);
      const double w = cfg.cart_task_weight[t];
      const double* W = (t < 5) ? cfg.ee_weight[t] : cfg.trunk_weight;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        if (t < 5) {                                             // A = W (J w)          (:480-482)
#pragma unroll
          for (int c = 0; c < 6; ++c) s += W[6 * r + c] * (Jc[c] * w);
        } else {                                                 // A = (W J) w          (:488-490)
#pragma unroll
          for (int c = 0; c < 6; ++c) s += W[6 * r + c] * Jc[c];
          s *= w;
        }
        a[6 * t + r] = s;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) a[6 * t + r] = 0.0;
    }
  }

  // ---------------------------------------------------------------- targets b (lanes 0..5), constraint bounds
  double fkq[4] = {0, 0, 0, 1};
  if (lane == 5) scipy_quat_from_matrix(oMf + WBC_FRAME_TRUNK * WBC_T_STRIDE, fkq);
  if (lane < 6 && ((cfg.task_mask >> lane) & 1)) {
    double b6[6];
    if (lane < 5) ee_task_target(cfg, lane, oMf, io, inv_dt, b6);
    else trunk_task_target(cfg, oMf, fkq, io, inv_dt, b6);
#pragma unroll
    for (int r = 0; r < 6; ++r) bs[6 * lane + r] = b6[r];
  } else if (lane < 6) {
#pragma unroll
    for (int r = 0; r < 6; ++r) bs[6 * lane + r] = 0.0;
  }
  // row layout of C
  int row_com = -1, row_trunk = -1, row_ee[5], nrows = 0;
  if (cfg.constraint_mask & WBC_CON_COM) { row_com = nrows; nrows += 2; }
  if (cfg.constraint_mask & WBC_CON_TRUNK) { row_trunk = nrows; nrows += 4; }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    row_ee[i] = -1;
    if (cfg.constraint_mask & (WBC_CON_FR << i)) { row_ee[i] = nrows; nrows += 3; }
  }
  const int row_extra = nrows;
  const int nC = nrows + cfg.n_extra_rows;
  if (lane < 32) { clbs[lane] = 0.0; cubs[lane] = 0.0; }
  __syncwarp();
  if (row_trunk >= 0 && lane == 5) {                             // trunkConstraint (:707-754)
    double eul[3];
    scipy_euler_xyz_from_quat(fkq, eul);
    const double* ip = io + WBC_IO_REF + REF_INIT_TRUNK_POS;
    const double* ie = io + WBC_IO_REF + REF_INIT_TRUNK_EUL;
    const double cur[4] = {oMf[WBC_FRAME_TRUNK * WBC_T_STRIDE + 11], eul[0], eul[1], eul[2]};
    const double z_var = ip[2] * 0.25;
    const double var = 1.5 * 0.1;
    const double lo[4] = {ip[2] - z_var, ie[0] - var, ie[1] - var, ie[2] - var};
    const double up[4] = {ip[2] + z_var, ie[0] + var, ie[1] + var, ie[2] + var};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      clbs[row_trunk + r] = ((lo[r] - cur[r]) * inv_dt) * 0.5;
      cubs[row_trunk + r] = ((up[r] - cur[r]) * inv_dt) * 0.5;
    }
  }
  if (row_com >= 0 && lane == 0) {                               // CoMConstraint (:669-694)
    const double* FL = oMf + 1 * WBC_T_STRIDE + 9;
    const double* RR = oMf + 2 * WBC_T_STRIDE + 9;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      clbs[row_com + r] = ((RR[r] - com_w[r]) * inv_dt) * 0.8;
      cubs[row_com + r] = ((FL[r] - com_w[r]) * inv_dt) * 0.8;
    }
  }
  if (lane < cfg.n_extra_rows) {
    clbs[row_extra + lane] = cfg.extra_lo[lane];
    cubs[row_extra + lane] = cfg.extra_hi[lane];
  }

  // ---------------------------------------------------------------- box bounds, joint task
  double lbv = 0.0, ubv = 0.0;
  if (lane < NV) damper_bounds(M, cfg, qs, lane, lbv, ubv);
  const bool joint_on = (cfg.task_mask & WBC_TASK_JOINT) != 0;
  const double aj = joint_on ? (1.0 / NV) * cfg.joint_task_weight : 0.0;       // qpJointA (:1199-1206)
  double bj = 0.0;
  if (joint_on && cfg.joint_mode == WBC_JOINT_PREV && lane < NV)               // qpJointb "PREV" (:1216-1217)
    bj = ((1.0 / NV) * qs[(lane < 6) ? lane : lane + 1]) * cfg.joint_task_weight;

  // ---------------------------------------------------------------- H = A^T A, g = -A^T b
  // Column `lane` of the 36 Cartesian rows goes to shared memory transposed (AsT[lane][r]); then for every
  // (task t, supporting column l) pair lane i adds  sum_r A[6t+r][i] A[6t+r][l]  to H[i][l]: three 128-bit
  // broadcast loads and six FMAs.  Pairs outside the frame's support are structural zeros and are skipped.
  __syncwarp();                      // all reads of oMi (aliased by Hs) and bs writes are done
  if (lane < NV) {
    double2* dst = reinterpret_cast<double2*>(AsT + lane * WBC_LDT);
#pragma unroll
    for (int r = 0; r < 18; ++r) dst[r] = make_double2(a[2 * r], a[2 * r + 1]);
    double* Hrow = Hs + lane * LD;
#pragma unroll
    for (int l = 0; l < NV; ++l) Hrow[l] = (l == lane) ? aj * aj : 0.0;
  }
  __syncwarp();
  double gk = 0.0;
  if (lane < NV) {
#pragma unroll
    for (int r = 0; r < 36; ++r) gk -= a[r] * bs[r];
    gk -= aj * bj;
  }
  {
    double* Hrow = Hs + (lane < NV ? lane : 0) * LD;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      if (!((cfg.task_mask >> t) & 1)) continue;
      uint32_t mask = M->frame_supp[t];
      while (mask) {                                               // warp-uniform
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        const double2* c2 = reinterpret_cast<const double2*>(AsT + l * WBC_LDT + 6 * t);
        const double2 v0 = c2[0], v1 = c2[1], v2 = c2[2];
        double h0 = a[6 * t] * v0.x, h1 = a[6 * t + 1] * v0.y;
        h0 = fma(a[6 * t + 2], v1.x, h0); h1 = fma(a[6 * t + 3], v1.y, h1);
        h0 = fma(a[6 * t + 4], v2.x, h0); h1 = fma(a[6 * t + 5], v2.y, h1);
        if (lane < NV) Hrow[l] += h0 + h1;
      }
    }
  }
  __syncwarp();

  if (DEBUG_OUT) {
    const WbcAssembleOut& D = P.dbg;
    const int m = P.m_rows;
    if (lane < NV) {
      int row = 0;
      for (int t = 0; t < 6; ++t) {
        if (!((cfg.task_mask >> t) & 1)) continue;
        for (int r = 0; r < 6; ++r, ++row) {
          if (D.A) D.A[(sidx * m + row) * NV + lane] = AsT[lane * WBC_LDT + 6 * t + r];
          if (D.b && lane == 0) D.b[sidx * m + row] = bs[6 * t + r];
        }
      }
      if (joint_on) {
        for (int r = 0; r < NV; ++r) {
          if (D.A) D.A[(sidx * m + row + r) * NV + lane] = (r == lane) ? aj : 0.0;
        }
        if (D.b) D.b[sidx * m + row + lane] = bj;
      }
      if (D.lb) D.lb[sidx * NV + lane] = lbv;
      if (D.ub) D.ub[sidx * NV + lane] = ubv;
      if (D.g) D.g[sidx * NV + lane] = gk;
      if (D.H)
        for (int l = 0; l < NV; ++l) D.H[(sidx * NV + lane) * NV + l] = Hs[lane * LD + l];
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- constraint rows (AsT is dead now)
  if (lane < NV) {
    if (row_com >= 0) { Cs[(row_com + 0) * LD + lane] = Jcom[0]; Cs[(row_com + 1) * LD + lane] = Jcom[1]; }
    if (row_trunk >= 0) {                                        // LWA rows z, wx, wy, wz of the trunk frame (:709)
      double Jc[6];
      frame_jac_column(Sc, M->frame_supp[WBC_FRAME_TRUNK], lane, oMf + WBC_FRAME_TRUNK * WBC_T_STRIDE,
                       WBC_RF_LOCAL_WORLD_ALIGNED, Jc);
#pragma unroll
      for (int r = 0; r < 4; ++r) Cs[(row_trunk + r) * LD + lane] = Jc[2 + r];
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if (row_ee[i] >= 0) {                                      // WORLD linear rows (:758)
        const bool sup = (M->frame_supp[i] >> lane) & 1u;
#pragma unroll
        for (int r = 0; r < 3; ++r) Cs[(row_ee[i] + r) * LD + lane] = sup ? Sc[r] : 0.0;
      }
    }
    for (int e = 0; e < cfg.n_extra_rows; ++e) {                 // extension rows (not in the reference)
      const int f = cfg.extra_frame[e];
      double Jc[6];
      frame_jac_column(Sc, M->frame_supp[f], lane, oMf + f * WBC_T_STRIDE, cfg.extra_rf[e], Jc);
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) s += cfg.extra_coeff[e][c] * Jc[c];
      Cs[(row_extra + e) * LD + lane] = s;
    }
  }
  phase_sync<PS>();
  const double clb_r = (lane < nC) ? clbs[lane] : 0.0;
  const double cub_r = (lane < nC) ? cubs[lane] : 0.0;

  if (DEBUG_OUT) {
    const WbcAssembleOut& D = P.dbg;
    if (D.C && lane < NV)
      for (int r = 0; r < nC; ++r) D.C[(sidx * nC + r) * NV + lane] = Cs[r * LD + lane];
    if (D.Clb && lane < nC) D.Clb[sidx * nC + lane] = clb_r;
    if (D.Cub && lane < nC) D.Cub[sidx * nC + lane] = cub_r;
    if (P.io.mem_out)
      for (int i = lane; i < WBC_MEM_STRIDE; i += 32) P.io.mem_out[sidx * WBC_MEM_STRIDE + i] = io[WBC_IO_MEM + i];
    return;
  }

  // ---------------------------------------------------------------- QP
  double x;
  QpResult res;
  {
    double h[NV];
    const double* Hrow = Hs + (lane < NV ? lane : 0) * LD;
#pragma unroll
    for (int l = 0; l < NV; ++l) h[l] = (lane < NV) ? Hrow[l] : 0.0;
    const double hdiag = (lane < NV) ? Hrow[lane] : 0.0;
    __syncwarp();                    // Hs becomes the solver's R factor
    QpRegShared S;
    S.R = Hs; S.col = ws + L.col; S.vd = vd; S.C = Cs;
    res = warp_qp_solve_reg<NV, SPLIT>(S, h, hdiag, nC, gk, lbv, ubv, clb_r, cub_r, cfg.max_iter, x);
  }

  phase_sync<PS>();
  if (valid && lane < NV) P.io.qdot[sidx * NV + lane] = x;
  if (valid && lane == 0) {
    P.io.status[sidx] = res.status;
    P.io.iters[sidx] = res.iters;
    if (P.io.active_set) {
      P.io.active_set[2 * sidx] = res.act_box;
      P.io.active_set[2 * sidx + 1] = res.act_rows;
    }
  }
  if (valid && P.io.mem_out)
    for (int i = lane; i < WBC_MEM_STRIDE; i += 32) P.io.mem_out[sidx * WBC_MEM_STRIDE + i] = io[WBC_IO_MEM + i];

  // ---------------------------------------------------------------- integrate + base estimate
  if (P.io.q_next) {
    __syncwarp();
    const double v = x * dt;
    double vb[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) vb[i] = __shfl_sync(WBC_FULL_MASK, v, i);
    double* qn = ws + L.col;         // [<= 33] new configuration (the 64-double column buffer is free now)
    if (lane == 0) {
      double o7[7];
      integrate_freeflyer(qs, vb, o7);
#pragma unroll
      for (int i = 0; i < 7; ++i) qn[i] = o7[i];
    }
    if (lane >= 6 && lane < NV) {
      const int iq = M->col_q[lane];
      qn[iq] = qs[iq] + v;
    }
    __syncwarp();
    if (!(P.flags & WBC_STEP_FLAG_PLAIN_INTEGRATE)) {
      // updateState(joint_config, imu, running=True): q = [old xyz, imu quat, joints], FK, trunkWorldPos (:387-428)
      if (lane < 3) qn[lane] = qs[lane];
      if (P.io.imu_quat && lane < 4) qn[3 + lane] = P.io.imu_quat[sidx * 4 + lane];
      __syncwarp();
      warp_fk(M, qn, oMi, lane);
      double bp[3] = {0, 0, 0};
      if (lane < 4) {                // foot frame positions at the new configuration
        const int par = M->frame_parent[lane];
        const double* Pm = oMi + par * WBC_T_STRIDE;
        double Rp[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rp[i] = Pm[i];
        mat3_vec(Rp, M->frp[lane], bp);
        bp[0] += Pm[9]; bp[1] += Pm[10]; bp[2] += Pm[11];
      }
      // trunkWorldPos (:1297-1327): order of the sums follows the reference (FR + FL + RR + RL) / 4
      double BPA[3], WPA[3];
      const double* Tt = oMi + M->frame_parent[WBC_FRAME_TRUNK] * WBC_T_STRIDE;   // trunk frame has identity offset?
      double Rt[9], pt[3];
      {
        double Rp[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rp[i] = Tt[i];
        mat3_mul(Rp, M->frR[WBC_FRAME_TRUNK], Rt);
        mat3_vec(Rp, M->frp[WBC_FRAME_TRUNK], pt);
        pt[0] += Tt[9]; pt[1] += Tt[10]; pt[2] += Tt[11];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double d0 = __shfl_sync(WBC_FULL_MASK, bp[c], 0) - pt[c];
        const double d1 = __shfl_sync(WBC_FULL_MASK, bp[c], 1) - pt[c];
        const double d2 = __shfl_sync(WBC_FULL_MASK, bp[c], 2) - pt[c];
        const double d3 = __shfl_sync(WBC_FULL_MASK, bp[c], 3) - pt[c];
        BPA[c] = (d0 + d1 + d2 + d3) / 4;
        const double* tg = io + WBC_IO_TARGETS;
        WPA[c] = (tg[c] + tg[3 + c] + tg[6 + c] + tg[9 + c]) / 4;
      }
      double rb[3];
      mat3_vec(Rt, BPA, rb);
      if (lane < 3) qn[lane] = WPA[lane] - rb[lane];
      __syncwarp();
    }
    if (valid)
      for (int i = lane; i < nq; i += 32) P.io.q_next[sidx * nq + i] = qn[i];
  }
}
