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

// per-state input block in shared memory (double-buffered: the next state's block is fetched with cp.async
// while the QP of the current one runs)
#define WBC_IN_Q 0          // q[nq <= 33]
#define WBC_IN_TARGETS 34   // 18
#define WBC_IN_MEM 52       // 72
#define WBC_IN_REF 124      // 24
#define WBC_IN_IMU 28       // 4: the IMU quaternion fed back after the tick (runWBC's base_config) -- in the tail of the q
                            //    slot (the fused kernel is instantiated for nq <= 28 only), so that the block stays 148 wide:
                            //    the general instantiation's 12 warps per SM fill the 227 KB of shared memory to the byte
#define WBC_IN_MBAR 32      // 1: the mbarrier the bulk (TMA) copies of this buffer complete on -- also in the tail of the q slot
#define WBC_IN_TOTAL 148
#define WBC_HOT_FRAMES 6
#define WBC_STEP_FLAG_WEIGHTS_IDENTITY 0x10000   // internal (set by the host wrapper): every 6x6 task weight is I

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
  int hs, ast, col, omf, vec, in, total;
};

__host__ __device__ inline int wbc_ld(int nv) { return nv | 1; }

// (sized for WBC_MAX_NC rows whatever the configuration: the layout is a compile-time constant of the kernel, so every
//  shared-memory address of the tick is [per-warp base + immediate])
// red: the instantiation with the reduced QP front (three locked DoFs, <= 16 rows): no H rows and no transposed task
// rows; the first block holds the oMi scratch, then the front's scratch (576 doubles, wbc_qp_red.inc), then the R factor
// of the inequality block -- or, for a state on the fallback path, the L columns / R factor of the general solver
// ((nv - 3) rows of nv | 1 doubles + the 32-lane store overhang); the second block only the constraint rows.
__host__ __device__ constexpr StepLayout step_layout(int nv, int nC = WBC_MAX_NC, bool red = false) {
  StepLayout L{};
  const int ld = nv | 1;
  int r0 = red ? (nv - 3) * ld + 32 : nv * (nv + 2);   // H rows, later the L columns / R factor of the QP ...
  const int fk = WBC_MAX_JOINTS * WBC_T_STRIDE;        // ... aliased by the oMi scratch (dead before H is written)
  if (r0 < fk) r0 = fk;
  if (red && r0 < 576) r0 = 576;
  r0 = (r0 + 1) & ~1;
  int r1 = red ? 16 * ld + 2 : nv * WBC_LDT;           // transposed task rows AsT, later the constraint rows C
  if (!red && r1 < nC * ld + 2) r1 = nC * ld + 2;
  r1 = (r1 + 1) & ~1;
  L.hs = 0;
  L.ast = r0;
  L.col = r0 + r1;
  L.omf = L.col + 64;
  L.vec = L.omf + WBC_HOT_FRAMES * WBC_T_STRIDE + 2;   // 6 * 13 + 2 = 80
  L.in = L.vec + 32 + 32 + 32 + 64;                    // vd[32] clb[32] cub[32] bs[64] (b, then the QP's |d|^2)
  L.total = L.in + 2 * WBC_IN_TOTAL;
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
  int row_com, row_trunk, row_ee[5], row_extra;   // first row of each constraint block in C (-1: off), set_rows()
  // reduced (null-space) QP front (wbc_qp_red.inc): usable for this model + configuration, first C row of the foot
  // owning limb columns [6 + 3 j, 9 + 3 j) (one byte per j), bit mask of the foot rows; set_reduced()
  int red_ok;
  unsigned red_rows, red_feet_mask, red_blk, red_other;
  int grid_cap;                                   // > 0: at most this many CTAs (wbc_step_host runs two slices side by side)
  // optional FP32 I/O (wbc_step_host with WBC_HOST_F32): which arrays hold float32 elements instead of float64
  int f32_in;                                     // WBC_F32_Q | _TARGETS | _MEM | _REF | _IMU
  int f32_out;                                    // WBC_F32_QDOT | _JOINTS
  // targets / task memory / references (/ IMU quaternion) are staged by bulk asynchronous copies (cp.async.bulk, the
  // TMA engine: one instruction per array and state, completion on an mbarrier) instead of one 8-byte cp.async per lane
  // and 32 doubles: float64 arrays whose base pointers are 16-byte aligned (their rows are 144 / 576 / 192 / 32 B).  Rows
  // of q (8 nq = 216 B for A1+WX200) are only 8-byte aligned and stay on cp.async.
  int bulk_in;                                    // WBC_BULK_TARGETS | _MEM | _REF | _IMU (0: everything on cp.async)
  // closed-loop horizon in ONE launch (the MULTI instantiation, wbc_rollout): K ticks; io.targets / io.imu_quat are then
  // trajectories [K, N, 18] / [K, N, 4], q and the task memory are advanced in place (io.q_next == io.q, io.mem_out == io.mem_in)
  int K;
  // zero-copy host call: the first warp of each barrier group writes the status / iteration count of the group's (consecutive)
  // states in one coalesced store behind the post-QP barrier, instead of one 4-byte store per warp -- over PCIe every store
  // is a transaction of its own, and 110 M four-byte writes per second cost the closed-loop host tick 9 %
  int group_report;
};
#define WBC_GROUP_REPORT_BYTES 512   // CTA-shared: [2 parities][32 warps] status + the same for iters (reduced-front instantiations)
#define WBC_BULK_TARGETS 1
#define WBC_BULK_MEM 2
#define WBC_BULK_REF 4
#define WBC_BULK_IMU 8
#define WBC_F32_Q 1
#define WBC_F32_TARGETS 2
#define WBC_F32_MEM 4
#define WBC_F32_REF 8
#define WBC_F32_IMU 16
#define WBC_F32_DELTA 32       // float32 targets / IMU quaternion are increments over the previous targets / the resident quaternion
#define WBC_F32_QDOT 1
#define WBC_F32_JOINTS 2

// Row layout of C implied by the constraint mask (findConstraints order, Robot_Wrapper4.py:764-836): computed once on
// the host so that the kernel reads the offsets straight from the parameter bank.
inline void set_rows(StepParams* P) {
  int nrows = 0;
  P->row_com = P->row_trunk = -1;
  if (P->cfg.constraint_mask & WBC_CON_COM) { P->row_com = nrows; nrows += 2; }
  if (P->cfg.constraint_mask & WBC_CON_TRUNK) { P->row_trunk = nrows; nrows += 4; }
  for (int i = 0; i < 5; ++i) {
    P->row_ee[i] = -1;
    if (P->cfg.constraint_mask & (WBC_CON_FR << i)) { P->row_ee[i] = nrows; nrows += 3; }
  }
  P->row_extra = nrows;
}

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

// 8-byte asynchronous global -> shared copy (LDGSTS); rows of q are only 8-byte aligned (nq is odd)
__device__ __forceinline__ void cp_async8(uint32_t dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- bulk asynchronous copies (TMA engine, non-tensor form) + mbarrier ------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mb, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mb, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` of transactions on `mb`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mb) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mb)
               : "memory");
}
// all lanes poll; a copy that never completes (a bug, not a data-dependent event) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(mb), "r"(parity)
                 : "memory");
    if (!done && ++spins > (1 << 22)) __trap();
  } while (!done);
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

// one input array of the state: n elements into the slot at `slot_a` (n doubles wide).  float64: element i lands in its
// place; float32 (optional FP32 I/O): the n floats land in the upper half of the slot, widen_slot() spreads them out.
__device__ __forceinline__ void prefetch_slot(uint32_t slot_a, const double* src, long long s, int n, bool f32, int lane) {
  if (!f32) {
    const double* g = src + s * n;
    for (int i = lane; i < n; i += 32) cp_async8(slot_a + 8 * i, g + i);
  } else {
    const float* g = reinterpret_cast<const float*>(src) + s * n;
    for (int i = lane; i < n; i += 32) cp_async4(slot_a + 4 * n + 4 * i, g + i);
  }
}

// fetch the input block of state s: q, targets, task memory, references (+ the IMU quaternion): 1128 (+ 32) B for
// nq = 27, coalesced
// (sv: row of the state in io.targets / io.imu_quat -- s itself, or k N + s for tick k of a multi-tick launch)
template <int NV>
__device__ __forceinline__ void prefetch_inputs(const StepParams& P, long long s, long long sv, uint32_t in_a, int lane) {
  constexpr int nq = NV + 1;
  const WbcStepIO& io = P.io;
  if (P.bulk_in) {                               // (uniform) q by cp.async, the selected 16-byte aligned rows by the TMA engine
    const int bm = P.bulk_in;
    const double* qg = io.q + s * nq;
    if (lane < nq) cp_async8(in_a + 8 * (WBC_IN_Q + lane), qg + lane);
    const double* tg = io.targets + sv * WBC_TARGETS_STRIDE;
    const double* mg = io.mem_in + s * WBC_MEM_STRIDE;
    const double* rg = io.ref + s * WBC_REF_STRIDE;
    const bool imu = io.imu_quat != nullptr;
    if (lane == 0) {
      const uint32_t mb = in_a + 8 * WBC_IN_MBAR;
      mbar_expect_tx(mb, ((bm & WBC_BULK_TARGETS) ? 8u * WBC_TARGETS_STRIDE : 0u) + ((bm & WBC_BULK_MEM) ? 8u * WBC_MEM_STRIDE : 0u) +
                             ((bm & WBC_BULK_REF) ? 8u * WBC_REF_STRIDE : 0u) + ((imu && (bm & WBC_BULK_IMU)) ? 32u : 0u));
      if (bm & WBC_BULK_TARGETS) bulk_g2s(in_a + 8 * WBC_IN_TARGETS, tg, 8 * WBC_TARGETS_STRIDE, mb);
      if (bm & WBC_BULK_MEM) bulk_g2s(in_a + 8 * WBC_IN_MEM, mg, 8 * WBC_MEM_STRIDE, mb);
      if (bm & WBC_BULK_REF) bulk_g2s(in_a + 8 * WBC_IN_REF, rg, 8 * WBC_REF_STRIDE, mb);
      if (imu && (bm & WBC_BULK_IMU)) bulk_g2s(in_a + 8 * WBC_IN_IMU, io.imu_quat + sv * 4, 32, mb);
    }
    if (!(bm & WBC_BULK_TARGETS) && lane < WBC_TARGETS_STRIDE) cp_async8(in_a + 8 * (WBC_IN_TARGETS + lane), tg + lane);
    if (!(bm & WBC_BULK_MEM)) {
      cp_async8(in_a + 8 * (WBC_IN_MEM + lane), mg + lane);
      cp_async8(in_a + 8 * (WBC_IN_MEM + 32 + lane), mg + 32 + lane);
      if (lane < WBC_MEM_STRIDE - 64) cp_async8(in_a + 8 * (WBC_IN_MEM + 64 + lane), mg + 64 + lane);
    }
    if (!(bm & WBC_BULK_REF) && lane < WBC_REF_STRIDE) cp_async8(in_a + 8 * (WBC_IN_REF + lane), rg + lane);
    if (imu && !(bm & WBC_BULK_IMU) && lane < 4) cp_async8(in_a + 8 * (WBC_IN_IMU + lane), io.imu_quat + sv * 4 + lane);
  } else if (P.f32_in == 0) {                    // (uniform) the float64 layout: straight-line, one instruction per 32 doubles
    const double* qg = io.q + s * nq;
    if (lane < nq) cp_async8(in_a + 8 * (WBC_IN_Q + lane), qg + lane);
    if (nq > 32 && lane == 0) cp_async8(in_a + 8 * (WBC_IN_Q + 32), qg + 32);
    const double* tg = io.targets + sv * WBC_TARGETS_STRIDE;
    if (lane < WBC_TARGETS_STRIDE) cp_async8(in_a + 8 * (WBC_IN_TARGETS + lane), tg + lane);
    const double* mg = io.mem_in + s * WBC_MEM_STRIDE;
    cp_async8(in_a + 8 * (WBC_IN_MEM + lane), mg + lane);
    cp_async8(in_a + 8 * (WBC_IN_MEM + 32 + lane), mg + 32 + lane);
    if (lane < WBC_MEM_STRIDE - 64) cp_async8(in_a + 8 * (WBC_IN_MEM + 64 + lane), mg + 64 + lane);
    const double* rg = io.ref + s * WBC_REF_STRIDE;
    if (lane < WBC_REF_STRIDE) cp_async8(in_a + 8 * (WBC_IN_REF + lane), rg + lane);
    if (io.imu_quat && lane < 4) cp_async8(in_a + 8 * (WBC_IN_IMU + lane), io.imu_quat + sv * 4 + lane);
  } else {
    prefetch_slot(in_a + 8 * WBC_IN_Q, io.q, s, nq, P.f32_in & WBC_F32_Q, lane);
    prefetch_slot(in_a + 8 * WBC_IN_TARGETS, io.targets, sv, WBC_TARGETS_STRIDE, P.f32_in & WBC_F32_TARGETS, lane);
    prefetch_slot(in_a + 8 * WBC_IN_MEM, io.mem_in, s, WBC_MEM_STRIDE, P.f32_in & WBC_F32_MEM, lane);
    prefetch_slot(in_a + 8 * WBC_IN_REF, io.ref, s, WBC_REF_STRIDE, P.f32_in & WBC_F32_REF, lane);
    if (io.imu_quat) prefetch_slot(in_a + 8 * WBC_IN_IMU, io.imu_quat, sv, 4, P.f32_in & WBC_F32_IMU, lane);
  }
  cp_async_commit();
}

// FP32 I/O: the float32 elements that prefetch_slot() parked in the upper halves of their slots become float64 in
// place (all reads of the warp precede all writes: the two ranges overlap)
template <int NV>
__device__ __forceinline__ void widen_inputs(const StepParams& P, uint32_t in_a, int lane) {
  constexpr int nq = NV + 1;
  const int m = P.f32_in;
  float vq = 0.f, vq2 = 0.f, vt = 0.f, vm0 = 0.f, vm1 = 0.f, vm2 = 0.f, vr = 0.f, vi = 0.f;
  if ((m & WBC_F32_Q) && lane < nq) vq = lds_f32(in_a + 8 * WBC_IN_Q + 4 * nq + 4 * lane);
  if ((m & WBC_F32_Q) && nq > 32 && lane == 0) vq2 = lds_f32(in_a + 8 * WBC_IN_Q + 4 * nq + 4 * 32);
  if ((m & WBC_F32_TARGETS) && lane < WBC_TARGETS_STRIDE) vt = lds_f32(in_a + 8 * WBC_IN_TARGETS + 4 * WBC_TARGETS_STRIDE + 4 * lane);
  if (m & WBC_F32_MEM) {
    const uint32_t a = in_a + 8 * WBC_IN_MEM + 4 * WBC_MEM_STRIDE + 4 * lane;
    vm0 = lds_f32(a); vm1 = lds_f32(a + 128);
    if (lane < WBC_MEM_STRIDE - 64) vm2 = lds_f32(a + 256);
  }
  if ((m & WBC_F32_REF) && lane < WBC_REF_STRIDE) vr = lds_f32(in_a + 8 * WBC_IN_REF + 4 * WBC_REF_STRIDE + 4 * lane);
  if ((m & WBC_F32_IMU) && P.io.imu_quat && lane < 4) vi = lds_f32(in_a + 8 * WBC_IN_IMU + 16 + 4 * lane);
  __syncwarp();
  if ((m & WBC_F32_Q) && lane < nq) sts_f64(in_a + 8 * (WBC_IN_Q + lane), (double)vq);
  if ((m & WBC_F32_Q) && nq > 32 && lane == 0) sts_f64(in_a + 8 * (WBC_IN_Q + 32), (double)vq2);
  if ((m & WBC_F32_TARGETS) && lane < WBC_TARGETS_STRIDE) {
    double t = (double)vt;
    if (m & WBC_F32_DELTA)       // increment over the previous tick's target: prev_EE_pos[5][3] | prev_trunk_ref[3] of the task memory
      t += lds_f64(in_a + 8 * (WBC_IN_MEM + (lane < 15 ? MEM_PREV_EE_POS + lane : MEM_PREV_TRUNK_REF + lane - 15)));
    sts_f64(in_a + 8 * (WBC_IN_TARGETS + lane), t);
  }
  if (m & WBC_F32_MEM) {
    sts_f64(in_a + 8 * (WBC_IN_MEM + lane), (double)vm0);
    sts_f64(in_a + 8 * (WBC_IN_MEM + 32 + lane), (double)vm1);
    if (lane < WBC_MEM_STRIDE - 64) sts_f64(in_a + 8 * (WBC_IN_MEM + 64 + lane), (double)vm2);
  }
  if ((m & WBC_F32_REF) && lane < WBC_REF_STRIDE) sts_f64(in_a + 8 * (WBC_IN_REF + lane), (double)vr);
  if ((m & WBC_F32_IMU) && P.io.imu_quat && lane < 4) {
    double t = (double)vi;
    if (m & WBC_F32_DELTA) t += lds_f64(in_a + 8 * (WBC_IN_Q + 3 + lane));     // increment over the resident base quaternion
    sts_f64(in_a + 8 * (WBC_IN_IMU + lane), t);
  }
  __syncwarp();
}

#define WBC_MOFF(f) ((uint32_t)offsetof(DevModel, f))

#ifndef WBC_FK_SCAN
#define WBC_FK_SCAN 1
#endif
// forwardKinematics into shared memory (stride WBC_T_STRIDE); 32-bit shared addressing.
__device__ __forceinline__ void warp_fk_a(uint32_t M_a, uint32_t q_a, uint32_t oMi_a, int lane) {
  double Rl[9], pl[3];
  const int nj = lds_s32(M_a + WBC_MOFF(njoints));
  const bool active = lane >= 1 && lane < nj;
  if (active) {
    const int jt = lds_s32(M_a + WBC_MOFF(jtype) + 4 * lane);
    const int iq = lds_s32(M_a + WBC_MOFF(idx_q) + 4 * lane);
    const bool ident = lds_s32(M_a + WBC_MOFF(pl_ident) + 4 * lane) != 0;
    const uint32_t qa = q_a + 8 * iq;
    double Rj[9], pj[3] = {0.0, 0.0, 0.0}, plp[3];
    lds_vec3(M_a + WBC_MOFF(plp) + 24 * lane, plp);
    if (jt == WBC_JT_FREEFLYER) {
      quat_to_matrix_eigen(lds_f64(qa + 24), lds_f64(qa + 32), lds_f64(qa + 40), lds_f64(qa + 48), Rj);
      lds_vec3(qa, pj);
    } else {
      double ax[3];
      lds_vec3(M_a + WBC_MOFF(axis) + 24 * lane, ax);
      const double qv = lds_f64(qa);
      if (jt == WBC_JT_REVOLUTE) {
        double sn, cs;
        sincos(qv, &sn, &cs);
        axis_angle_matrix(ax, sn, cs, Rj);
      } else {  // prismatic
        Rj[0] = 1; Rj[1] = 0; Rj[2] = 0; Rj[3] = 0; Rj[4] = 1; Rj[5] = 0; Rj[6] = 0; Rj[7] = 0; Rj[8] = 1;
        pj[0] = ax[0] * qv; pj[1] = ax[1] * qv; pj[2] = ax[2] * qv;
      }
    }
    if (ident) {                       // pure-translation placement: I * Rj and I * pj are exact, skip them
#pragma unroll
      for (int i = 0; i < 9; ++i) Rl[i] = Rj[i];
      pl[0] = pj[0]; pl[1] = pj[1]; pl[2] = pj[2];
    } else {
      double PR[9];
      lds_mat3(M_a + WBC_MOFF(plR) + 72 * lane, PR);
      mat3_mul(PR, Rj, Rl);
      mat3_vec(PR, pj, pl);
    }
    if (jt == WBC_JT_REVOLUTE) { pl[0] = plp[0]; pl[1] = plp[1]; pl[2] = plp[2]; }   // (placement * 0 + p, exactly p)
    else { pl[0] += plp[0]; pl[1] += plp[1]; pl[2] += plp[2]; }
  }
#if !WBC_FK_SCAN
  {
    const int par = active ? lds_s32(M_a + WBC_MOFF(parent) + 4 * lane) : 0;
    const int myDepth = active ? lds_s32(M_a + WBC_MOFF(depth) + 4 * lane) : 0;
    const int maxdepth = lds_s32(M_a + WBC_MOFF(maxdepth));
    const uint32_t out_a = oMi_a + 8 * WBC_T_STRIDE * lane;
    for (int d = 1; d <= maxdepth; ++d) {
      if (active && myDepth == d) {
        if (par == 0) {
#pragma unroll
          for (int i = 0; i < 9; ++i) sts_f64(out_a + 8 * i, Rl[i]);
          sts_f64(out_a + 72, pl[0]); sts_f64(out_a + 80, pl[1]); sts_f64(out_a + 88, pl[2]);
        } else {
          const uint32_t pa = oMi_a + 8 * WBC_T_STRIDE * par;
          double Rp[9], R[9], p[3], pp[3];
          lds_mat3(pa, Rp);
          lds_vec3(pa + 72, pp);
          mat3_mul(Rp, Rl, R);
          mat3_vec(Rp, pl, p);
#pragma unroll
          for (int i = 0; i < 9; ++i) sts_f64(out_a + 8 * i, R[i]);
          sts_f64(out_a + 72, p[0] + pp[0]); sts_f64(out_a + 80, p[1] + pp[1]); sts_f64(out_a + 88, p[2] + pp[2]);
        }
      }
      __syncwarp();
    }
  }
}
#else
  // oMi[j] = liMi[root] ... liMi[parent(j)] liMi[j] by pointer jumping: after round r every joint holds the product
  // over its 2^(r+1) nearest ancestors (itself included), so ceil(log2(depth)) rounds replace `depth` serial levels and
  // every lane works in every round.  The own transform stays in registers; only the ancestor's comes from shared
  // memory.  (The association order of the products differs from Pinocchio's root-to-leaf sweep: rounding-level, 1e-16.)
  // Divergence-free: slot 0 (the universe) and the slots of the idle lanes hold the identity, and a joint without a
  // 2^r-th ancestor composes with slot 0 (1 * x + 0 * y + 0 * z is exact), so all 32 lanes run the same code.
  // Slot index = joint index = lane, so the ancestor's transform comes by warp shuffle straight out of its lane's registers:
  // no stores between the rounds, no barrier pairs, half the shared-memory pipe time (the pipe is the busiest unit of the
  // tick, ncu); only the final transforms go to shared memory for the Jacobian columns and the frames.
  const uint32_t out_a = oMi_a + 8 * WBC_T_STRIDE * lane;
  if (!active) {
#pragma unroll
    for (int i = 0; i < 9; ++i) Rl[i] = (i % 4 == 0) ? 1.0 : 0.0;
    pl[0] = pl[1] = pl[2] = 0.0;
  }
  const int nrounds = lds_s32(M_a + WBC_MOFF(nrounds));
#pragma unroll 1
  for (int r = 0; r < nrounds; ++r) {
    const int anc = lds_s32(M_a + WBC_MOFF(anc) + 4 * (WBC_MAX_JOINTS * r + lane));
    double Rp[9], pp[3], R[9], p[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rp[i] = __shfl_sync(WBC_FULL_MASK, Rl[i], anc);
#pragma unroll
    for (int i = 0; i < 3; ++i) pp[i] = __shfl_sync(WBC_FULL_MASK, pl[i], anc);
    mat3_mul(Rp, Rl, R);
    mat3_vec(Rp, pl, p);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rl[i] = R[i];
    pl[0] = p[0] + pp[0]; pl[1] = p[1] + pp[1]; pl[2] = p[2] + pp[2];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) sts_f64(out_a + 8 * i, Rl[i]);
  sts_f64(out_a + 72, pl[0]); sts_f64(out_a + 80, pl[1]); sts_f64(out_a + 88, pl[2]);
  __syncwarp();
}
#endif

// velDamperJointConstraints (Robot_Wrapper4.py:572-637) for velocity index k.
__device__ __forceinline__ void damper_bounds_a(uint32_t M_a, const WbcConfig& cfg, uint32_t q_a, int k, double& lbv,
                                                double& ubv) {
  const int grip = cfg.gripper_joint_id;
  const int qidx = (k < 6) ? k : k + 1;                        // np.delete(., 6)
  double lo, up;
  if (qidx < 7) { lo = -5.0; up = 5.0; }
  else if (qidx >= grip - 2 + 7) { lo = 0.0; up = 0.0; }
  else { lo = lds_f64(M_a + WBC_MOFF(lower) + 8 * qidx); up = lds_f64(M_a + WBC_MOFF(upper) + 8 * qidx); }
  const double vel = (k < 7) ? 5.0 : lds_f64(M_a + WBC_MOFF(velocity) + 8 * k);   // vel_lim[i] = 5 for i < 7 (:593)
  const double c = lds_f64(q_a + 8 * ((cfg.compat_flags & WBC_COMPAT_DAMPER_OFF_BY_ONE) ? k : qidx));   // quirk D.2
  const double coef = cfg.damper_coef, qi = cfg.damper_qi, qsv = cfg.damper_qs;
  const double inv_zone = 1.0 / (qi - qsv);      // uniform over the launch (hoisted); x * (1 / d) is within 1 ulp of x / d
  if (c <= lo + qi) {
    lbv = (-coef * (c - lo - qsv)) * inv_zone;
    if (lbv > vel) lbv = vel;
    if (lbv < -vel) lbv = -vel;
  } else {
    lbv = -vel;
  }
  if (c >= up - qi) {
    ubv = (coef * (up - c - qsv)) * inv_zone;
    if (ubv < -vel) ubv = -vel;
    if (ubv > vel) ubv = vel;
  } else {
    ubv = vel;
  }
  if (lbv > 0) lbv = lbv * -1;
  if (ubv < 0) ubv = ubv * -1;
  if (k >= grip - 2 + 6) { lbv = 0.0; ubv = 0.0; }
}

// getFrameJacobian column from the WORLD joint-Jacobian column S for a frame at position p (LOCAL_WORLD_ALIGNED)
// or as is (WORLD); zero outside the support.
__device__ __forceinline__ void frame_jac_lwa(const double* S, bool sup, const double* p, bool world, double* Jc) {
  if (!sup) {
#pragma unroll
    for (int i = 0; i < 6; ++i) Jc[i] = 0.0;
    return;
  }
  if (world) {
#pragma unroll
    for (int i = 0; i < 6; ++i) Jc[i] = S[i];
    return;
  }
  double pxw[3];
  cross3(p, S + 3, pxw);
  Jc[0] = S[0] - pxw[0]; Jc[1] = S[1] - pxw[1]; Jc[2] = S[2] - pxw[2];
  Jc[3] = S[3]; Jc[4] = S[4]; Jc[5] = S[5];
}

// column `lane` of data.J (computeJointJacobians, WORLD frame): [lin; ang]
template <int NV>
__device__ __forceinline__ void warp_jac_col_a(uint32_t M_a, uint32_t omi_a, int lane, double* Sc) {
#pragma unroll
  for (int i = 0; i < 6; ++i) Sc[i] = 0.0;
  if (lane < NV) {
    const uint32_t Ta = omi_a + 8 * WBC_T_STRIDE * lds_s32(M_a + WBC_MOFF(col_joint) + 4 * lane);
    double R[9], axl[3], aw[3];
    lds_mat3(Ta, R);
    lds_vec3(M_a + WBC_MOFF(col_axis) + 24 * lane, axl);
    mat3_vec(R, axl, aw);
    if (lds_s32(M_a + WBC_MOFF(col_ang) + 4 * lane)) {
      double p[3];
      lds_vec3(Ta + 72, p);
      cross3(p, aw, Sc);
      Sc[3] = aw[0]; Sc[4] = aw[1]; Sc[5] = aw[2];
    } else {
      Sc[0] = aw[0]; Sc[1] = aw[1]; Sc[2] = aw[2];
    }
  }
}

// sqrt(det(J J^T)) of getJointJacobian(joint, LOCAL_WORLD_ALIGNED): the manipulability measure of qpJointb's "MANI" /
// "HYBRID" modes (Robot_Wrapper4.py:1233-1236).  G_a: [32][6] scratch, Mm_a: [36] scratch.  Uniform result.
template <int NV>
__device__ __forceinline__ double warp_manip_a(uint32_t M_a, uint32_t omi_a, uint32_t G_a, uint32_t Mm_a, int jid, int lane) {
  double Sc[6], Jc[6], pj[3];
  warp_jac_col_a<NV>(M_a, omi_a, lane, Sc);
  const uint32_t supp = (uint32_t)lds_s32(M_a + WBC_MOFF(joint_supp) + 4 * jid);
  lds_vec3(omi_a + 8 * (WBC_T_STRIDE * jid + 9), pj);
  frame_jac_lwa(Sc, lane < NV && ((supp >> lane) & 1u), pj, false, Jc);
#pragma unroll
  for (int r = 0; r < 6; ++r) sts_f64(G_a + 8 * (6 * lane + r), Jc[r]);
  __syncwarp();
  if (lane < 21) {                       // upper triangle of the 6 x 6 Gram matrix, one entry per lane
    int r = 0, rem = lane;
    while (rem >= 6 - r) { rem -= 6 - r; ++r; }
    const int c = r + rem;
    double acc = 0.0;
#pragma unroll 2
    for (int k = 0; k < NV; ++k) acc = fma(lds_f64(G_a + 8 * (6 * k + r)), lds_f64(G_a + 8 * (6 * k + c)), acc);
    sts_f64(Mm_a + 8 * (6 * r + c), acc);
    sts_f64(Mm_a + 8 * (6 * c + r), acc);
  }
  __syncwarp();
  double f = 0.0;
  if (lane == 0) {                       // determinant by LU with partial pivoting (np.linalg.det)
    double det = 1.0;
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
      int piv = k;
      double best = fabs(lds_f64(Mm_a + 8 * (6 * k + k)));
#pragma unroll 1
      for (int i = k + 1; i < 6; ++i) {
        const double v = fabs(lds_f64(Mm_a + 8 * (6 * i + k)));
        if (v > best) { best = v; piv = i; }
      }
      if (piv != k) {
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
          const double t0 = lds_f64(Mm_a + 8 * (6 * k + c)), t1 = lds_f64(Mm_a + 8 * (6 * piv + c));
          sts_f64(Mm_a + 8 * (6 * k + c), t1);
          sts_f64(Mm_a + 8 * (6 * piv + c), t0);
        }
        det = -det;
      }
      const double d = lds_f64(Mm_a + 8 * (6 * k + k));
      det *= d;
      if (d == 0.0) break;
#pragma unroll 1
      for (int i = k + 1; i < 6; ++i) {
        const double l = lds_f64(Mm_a + 8 * (6 * i + k)) / d;
#pragma unroll 1
        for (int c = k + 1; c < 6; ++c)
          sts_f64(Mm_a + 8 * (6 * i + c), lds_f64(Mm_a + 8 * (6 * i + c)) - l * lds_f64(Mm_a + 8 * (6 * k + c)));
      }
    }
    f = sqrt(det);
  }
  __syncwarp();
  return __shfl_sync(WBC_FULL_MASK, f, 0);
}

// updateState (Robot_Wrapper4.py:387-428) for one state: forwardKinematics into oMi, column `lane` of the WORLD joint
// Jacobian (computeJointJacobians), the six hot frame placements (updateFramePlacements) and, when the CoM constraint
// is on, the centre of mass and column `lane` of jacobianCenterOfMass (rows x, y).
template <int NV>
__device__ __forceinline__ void warp_kin_a(uint32_t M_a, uint32_t q_a, uint32_t omi_a, uint32_t omf_a, int lane, bool com_on,
                                           double* Sc, double* com_w, double* Jcom) {
  warp_fk_a(M_a, q_a, omi_a, lane);
  warp_jac_col_a<NV>(M_a, omi_a, lane, Sc);
  if (lane < WBC_HOT_FRAMES) {         // hot frames: 5 EE + trunk
    const uint32_t Pa = omi_a + 8 * WBC_T_STRIDE * lds_s32(M_a + WBC_MOFF(frame_parent) + 4 * lane);
    double Rp[9], R[9], p[3], fp[3], pp[3];
    lds_mat3(Pa, Rp);
    lds_vec3(Pa + 72, pp);
    lds_vec3(M_a + WBC_MOFF(frp) + 24 * lane, fp);
    if (lds_s32(M_a + WBC_MOFF(fr_ident) + 4 * lane)) {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = Rp[i];
    } else {
      double FR[9];
      lds_mat3(M_a + WBC_MOFF(frR) + 72 * lane, FR);
      mat3_mul(Rp, FR, R);
    }
    mat3_vec(Rp, fp, p);
    const uint32_t oa = omf_a + 8 * WBC_T_STRIDE * lane;
#pragma unroll
    for (int i = 0; i < 9; ++i) sts_f64(oa + 8 * i, R[i]);
    sts_f64(oa + 72, p[0] + pp[0]); sts_f64(oa + 80, p[1] + pp[1]); sts_f64(oa + 88, p[2] + pp[2]);
  }
  // centre of mass (only when the CoM constraint is on): lane j -> m_j * c_j in the world
  com_w[0] = com_w[1] = com_w[2] = 0.0;
  Jcom[0] = Jcom[1] = 0.0;
  if (com_on) {
    double mc[4] = {0, 0, 0, 0};
    const int nj = lds_s32(M_a + WBC_MOFF(njoints));
    if (lane >= 1 && lane < nj) {
      const uint32_t Ta = omi_a + 8 * WBC_T_STRIDE * lane;
      double R[9], c[3], cl[3], pj[3];
      lds_mat3(Ta, R);
      lds_vec3(Ta + 72, pj);
      lds_vec3(M_a + WBC_MOFF(com) + 24 * lane, cl);
      mat3_vec(R, cl, c);
      const double m = lds_f64(M_a + WBC_MOFF(mass) + 8 * lane);
      mc[0] = m; mc[1] = m * (c[0] + pj[0]); mc[2] = m * (c[1] + pj[1]); mc[3] = m * (c[2] + pj[2]);
    }
    // subtree sums for my column: sum over joints j in sub_joints[lane]
    double sm = 0, s1 = 0, s2 = 0, s3 = 0;
    const uint32_t sub = (lane < NV) ? (uint32_t)lds_s32(M_a + WBC_MOFF(sub_joints) + 4 * lane) : 0u;
    for (int j = 1; j < nj; ++j) {
      const double m0 = __shfl_sync(WBC_FULL_MASK, mc[0], j), m1 = __shfl_sync(WBC_FULL_MASK, mc[1], j);
      const double m2 = __shfl_sync(WBC_FULL_MASK, mc[2], j), m3 = __shfl_sync(WBC_FULL_MASK, mc[3], j);
      if ((sub >> j) & 1u) { sm += m0; s1 += m1; s2 += m2; s3 += m3; }
    }
    const double Mt = lds_f64(M_a + WBC_MOFF(total_mass));
    com_w[0] = warp_sum(mc[1]) / Mt; com_w[1] = warp_sum(mc[2]) / Mt; com_w[2] = warp_sum(mc[3]) / Mt;
    // Jcom column = (sm * lin - (sum m c) x ang) / M   (only x, y rows are used, :670)
    const double msc[3] = {s1, s2, s3};
    double cx[3];
    cross3(msc, Sc + 3, cx);
    Jcom[0] = (sm * Sc[0] - cx[0]) / Mt;
    Jcom[1] = (sm * Sc[1] - cx[1]) / Mt;
  }
}

#ifndef WBC_SYNC_TOP
#define WBC_SYNC_TOP 1         // phase barrier at the top of a tick (after the input prefetch has landed)
#endif
#ifndef WBC_SYNC_KIN
#define WBC_SYNC_KIN 1         // phase barrier after the kinematics
#endif
#ifndef WBC_SYNC_PREQP
#define WBC_SYNC_PREQP 1       // phase barrier before the QP
#endif
// Reduced-front instantiations (round 2): with their hot code down from 25 k to 21 k instructions the three barriers in front
// of the QP cost more than the re-alignment they buy -- the warps of a group leave the barrier after the QP together and stay
// close enough until the next one: device-resident +2.5 %, closed-loop host tick +0.6 %, config 5 +2.3 %.  The barrier
// after the QP stays (without it the closed-loop tick loses 9 %), and the general instantiations keep all four (bootstrap P1
// loses 2.6 % without the front three).
#ifndef WBC_SYNC_FRONT_RED
#define WBC_SYNC_FRONT_RED 0
#endif
#ifndef WBC_SYNC_POSTQP
#define WBC_SYNC_POSTQP 1      // phase barrier after the QP
#endif
#ifndef WBC_QP_MID_SYNC
#define WBC_QP_MID_SYNC 0
#endif
#ifndef WBC_PHASE_SYNC
#define WBC_PHASE_SYNC 1
#endif

// All WBC ticks of one warp: states first, first + stride, ...  `ws`: this warp's shared workspace.
// The warps of a CTA are re-aligned with a block barrier at the phase boundaries (PHASE_SYNC): the tick is
// ~10^4 mostly straight-line instructions, far more than the 32 KB instruction cache holds, so warps that drift
// apart each stream the whole code from L2 on their own (measured: 2.4x slower without the barriers); in step,
// one fetch feeds all warps.  Padding warps shadow the last state (no writes) so the barriers stay uniform.
// Shared memory is addressed through 32-bit shared-window addresses (wbc_device.cuh: smem_addr, lds_*, sts_*).
// NF: the last NF velocity DoFs are locked by the configuration (gripper + fingers, lb = ub = 0): the QP runs on NV - NF variables
// MULTI: P.K consecutive closed-loop ticks in this one launch.  A robot belongs to one warp for the whole horizon (its q and
// task memory are advanced in place by that warp alone), so no grid-wide synchronisation separates the ticks: a CTA walks
// (tick 0: its rounds), (tick 1: its rounds), ... and only the last tick of the horizon has a drain tail.
template <int NV, bool DEBUG_OUT, bool SPLIT, bool FD, int NF = 0, bool RED = false, bool MULTI = false>
__device__ __forceinline__ void warp_wbc_states(const StepParams& P, const DevModel* Ms, double* ws) {
  constexpr StepLayout L = step_layout(NV, WBC_MAX_NC, RED);
  static_assert(NV + 1 <= WBC_IN_IMU, "the IMU quaternion sits behind q inside the q slot");
  constexpr bool PS = WBC_PHASE_SYNC && !DEBUG_OUT;
  constexpr bool PSF = PS && (!RED || WBC_SYNC_FRONT_RED);     // the three barriers in front of the QP
  constexpr int LD = NV | 1;
  constexpr int nq = NV + 1;
  const int lane = threadIdx.x & 31;
  // (through a warp reduction: its result is warp-uniform for ptxas, so everything derived from it -- the per-warp shared base,
  //  the state index, the input buffer -- can live in uniform registers instead of the 128 per-thread ones)
  //  (measured: 4 096-state, bootstrap, HYBRID and closed-loop launches +4-5 %, headline +0.7 %; the full-width layout -6 %: plain there)
  const int warp = SPLIT ? (int)__reduce_max_sync(WBC_FULL_MASK, threadIdx.x >> 5) : (int)(threadIdx.x >> 5), wpc = blockDim.x >> 5;
  const WbcConfig& cfg = P.cfg;
  const double dt = P.io.dt;
  const double inv_dt = 1.0 / dt;       // the reference divides by dt; multiplying by 1/dt differs by <= 1 ulp
  const uint32_t M_a = smem_addr(Ms);
  const uint32_t ws_a = smem_addr(ws);
  const uint32_t hs_a = ws_a + 8 * L.hs;      // [NV][LD] rows of H; then the QP's L columns / R factor; oMi scratch before
  const uint32_t ast_a = ws_a + 8 * L.ast;    // [NV][WBC_LDT] transposed task rows; then [nC][LD] constraint rows
  const uint32_t omi_a = hs_a;
  const uint32_t omf_a = ws_a + 8 * L.omf;
  const uint32_t vd_a = ws_a + 8 * L.vec;
  const uint32_t clb_a = vd_a + 8 * 32, cub_a = clb_a + 8 * 32, bs_a = cub_a + 8 * 32;
  const uint32_t in0_a = ws_a + 8 * L.in;

  // row layout of C (uniform over the launch; parameter bank)
#define row_com P.row_com
#define row_trunk P.row_trunk
#define row_ee P.row_ee
#define row_extra P.row_extra
  const int nC = P.nC;
  const bool joint_on = (cfg.task_mask & WBC_TASK_JOINT) != 0;
  const double aj = joint_on ? (1.0 / NV) * cfg.joint_task_weight : 0.0;       // qpJointA (:1199-1206)
  const bool w_ident = (P.flags & WBC_STEP_FLAG_WEIGHTS_IDENTITY) != 0;        // every 6x6 task weight is the identity

  // (32-bit state indices: the host refuses N >= 2^31 - 2^20; 64-bit arithmetic only where an address is formed)
  const int NS = (int)P.N;
  const int stride = (int)gridDim.x * wpc;
  const int first = (int)blockIdx.x * wpc;
  // MULTI: the next tick of a robot reads what this tick's tail wrote.  With two or more rounds per tick the prefetch at the
  // top of an iteration fetches a state whose previous tick ended at least one iteration ago; with a single round it is the
  // state in flight, so the prefetch moves behind the tail (`late`)
  const bool late = MULTI && first + stride >= NS;
  int buf = 0;
  uint32_t mb_phase = 0;                        // bit b: parity the next wait on buffer b's mbarrier expects
  if (P.bulk_in) {
    if (lane == 0) {
      mbar_init(in0_a + 8 * WBC_IN_MBAR, 1);
      mbar_init(in0_a + 8 * (WBC_IN_TOTAL + WBC_IN_MBAR), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();
  }
  {
    int s0 = first + warp;
    if (s0 >= NS) s0 = NS - 1;
    prefetch_inputs<NV>(P, s0, s0, in0_a, lane);
  }
  // (toff = k N: row offset of tick k in the target / IMU trajectories; 0 and dead code unless MULTI)
  long long toff = 0;
  int tick = 0;
  for (int base = first; MULTI ? (tick < P.K && first < NS) : (base < NS);
       base += stride, (MULTI && base >= NS) ? (base = first, toff += NS, ++tick) : 0) {
    int sidx32 = base + warp;
    const bool valid = sidx32 < NS;
    if (!valid) sidx32 = NS - 1;
    const long long sidx = sidx32;
    const uint32_t in_a = in0_a + 8 * WBC_IN_TOTAL * buf;
    const uint32_t q_a = in_a + 8 * WBC_IN_Q, tg_a = in_a + 8 * WBC_IN_TARGETS;
    const uint32_t mem_a = in_a + 8 * WBC_IN_MEM, ref_a = in_a + 8 * WBC_IN_REF;
    cp_async_wait_all();
    if (P.bulk_in) {                                     // (uniform) the bulk copies of this buffer have landed
      mbar_wait(in_a + 8 * WBC_IN_MBAR, (mb_phase >> buf) & 1u);
      mb_phase ^= 1u << buf;
    }
    if (DEBUG_OUT && !valid) break;                      // (after the waits: no copy may be in flight when the CTA exits)
    if (P.f32_in) widen_inputs<NV>(P, in_a, lane);       // (uniform) optional FP32 I/O
    phase_sync<PSF && (WBC_SYNC_TOP != 0)>();
    // the next state's inputs travel while this tick computes: the other buffer is free (its last reader was the tail
    // of the previous tick), and a whole tick -- ~20 us per warp -- hides the latency even of a PCIe read (zero-copy
    // host buffers).  Issued here, not in front of the QP, so that its address arithmetic does not sit in the kernel's
    // region of highest register pressure.
    auto prefetch_next = [&]() {
      bool more = base + stride < NS;                    // another state in this tick ...
      int ns = base + stride + warp;
      long long nv_off = MULTI ? toff : 0;
      if (MULTI && !more && tick + 1 < P.K) {            // ... or the first one of the next tick
        more = true;
        ns = first + warp;
        nv_off = toff + NS;
      }
      if (ns >= NS) ns = NS - 1;
      if (more) {
        if (P.bulk_in) {         // the previous tick's generic-proxy writes into that buffer (prev targets / rotations) are
          fence_proxy_async();   // ordered before the TMA engine's writes
          __syncwarp();
        }
        prefetch_inputs<NV>(P, ns, nv_off + ns, in0_a + 8 * WBC_IN_TOTAL * (buf ^ 1), lane);
      }
    };
    if (!late) prefetch_next();

    // ---------------------------------------------------------------- kinematics
    double Sc[6];                        // column `lane` of data.J (WORLD): [lin; ang]
    double com_w[3], Jcom[2];
    warp_kin_a<NV>(M_a, q_a, omi_a, omf_a, lane, (cfg.constraint_mask & WBC_CON_COM) != 0, Sc, com_w, Jcom);
    phase_sync<PSF && (WBC_SYNC_KIN != 0)>();

    // ---------------------------------------------------------------- task rows for my column (registers)
    // EE tasks: A = W (J_LWA w) (:476-482); trunk task: A = (W J_WORLD) w (:488-490)
    uint32_t supp[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) supp[t] = (uint32_t)lds_s32(M_a + WBC_MOFF(frame_supp) + 4 * t);
    double a[36];
    // (computed here only in the finite-difference modes, whose perturbed state must not reach A; otherwise after the
    //  targets and bounds, right before its consumers: 72 registers less across the rotation chains)
    auto task_rows = [&]() {
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      const bool on = (cfg.task_mask >> t) & 1;
      if (on) {
        double Jc[6], fp[3];
        lds_vec3(omf_a + 8 * (WBC_T_STRIDE * t + 9), fp);
        // (outside the frame's support the column is zero: folded into the scalar weight, one select instead of six)
        frame_jac_lwa(Sc, true, fp, t == 5, Jc);
        const double w = ((supp[t] >> lane) & 1u) ? cfg.cart_task_weight[t] : 0.0;
        if (w_ident) {
#pragma unroll
          for (int r = 0; r < 6; ++r) a[6 * t + r] = Jc[r] * w;
        } else {
          const double* W = (t < 5) ? cfg.ee_weight[t] : cfg.trunk_weight;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            double sacc = 0.0;
            if (t < 5) {
#pragma unroll
              for (int c = 0; c < 6; ++c) sacc += W[6 * r + c] * (Jc[c] * w);
            } else {
#pragma unroll
              for (int c = 0; c < 6; ++c) sacc += W[6 * r + c] * Jc[c];
              sacc *= w;
            }
            a[6 * t + r] = sacc;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) a[6 * t + r] = 0.0;
      }
    }
    };
    if (FD) task_rows();

    // ---------------------------------------------------------------- targets b (lanes 0..5), constraint bounds
    // calcTargetVelEE3 (:1052-1157) on lanes 0..4 and calcTargetVelTrunk2 (:948-1015) on lane 5 share the
    // position law and the Euler -> quaternion -> matrix chain; they differ in the skew product and the trunk's
    // quaternion-error term.
    sts_f64(clb_a + 8 * lane, 0.0);
    sts_f64(cub_a + 8 * lane, 0.0);
    // Rotation.from_euler('xyz', default_ori) of the six targets needs 18 half-angle sincos: one per lane (lane 3 t + k:
    // target t, angle k), gathered by lane t (the default Euler triples are contiguous in `ref`)
    double hs0, hc0, hs1, hc1, hs2, hc2;
    {
      double hs, hc;
      sincos(lds_f64(ref_a + 8 * (REF_DEF_EE_ORI + (lane < 18 ? lane : 0))) / 2.0, &hs, &hc);
      const int t3 = 3 * (lane < 6 ? lane : 0);
      hs0 = __shfl_sync(WBC_FULL_MASK, hs, t3);     hc0 = __shfl_sync(WBC_FULL_MASK, hc, t3);
      hs1 = __shfl_sync(WBC_FULL_MASK, hs, t3 + 1); hc1 = __shfl_sync(WBC_FULL_MASK, hc, t3 + 1);
      hs2 = __shfl_sync(WBC_FULL_MASK, hs, t3 + 2); hc2 = __shfl_sync(WBC_FULL_MASK, hc, t3 + 2);
    }
    double fkq[4] = {0, 0, 0, 1};
    __syncwarp();
    if (lane < 6) {
      const bool on = (cfg.task_mask >> lane) & 1;
      const bool is_trunk = lane == 5;
      const uint32_t T_a = omf_a + 8 * WBC_T_STRIDE * lane;
      double b6[6] = {0, 0, 0, 0, 0, 0};
      if (is_trunk) {
        double Rf[9];
        lds_mat3(T_a, Rf);
        scipy_quat_from_matrix(Rf, fkq);
      }
      if (on) {
        const uint32_t prev_a = mem_a + 8 * (is_trunk ? MEM_PREV_TRUNK_REF : MEM_PREV_EE_POS + 3 * lane);
        const uint32_t prevR_a = mem_a + 8 * (is_trunk ? MEM_OLD_TRUNK_ROT : MEM_PREV_EE_ROT + 9 * lane);
        double target[3], prev[3], fk[3], ref_vel[3], err[3], ge[3];
        lds_vec3(tg_a + 24 * lane, target);
        lds_vec3(prev_a, prev);
        lds_vec3(T_a + 72, fk);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          ref_vel[k] = (target[k] - prev[k]) * inv_dt;
          err[k] = (target[k] - fk[k]) * inv_dt;
        }
        const double* G = is_trunk ? cfg.trunk_gain_pos : cfg.ee_gain_pos[lane < 5 ? lane : 0];
        mat3_vec(G, err, ge);
        double qref[4], Rref[9], prevR[9], dR[9], X[9], sk[9];
        scipy_quat_from_half_sincos(hs0, hc0, hs1, hc1, hs2, hc2, qref);
        scipy_matrix_from_quat(qref, Rref);
        lds_mat3(prevR_a, prevR);
#pragma unroll
        for (int k = 0; k < 9; ++k) dR[k] = (Rref[k] - prevR[k]) * inv_dt;
        // EE: (dR/dt) Rref^T (:1125); trunk: (dR/dt) Rref, NOT transposed (:984)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) X[3 * i + j] = is_trunk ? Rref[3 * i + j] : Rref[3 * j + i];
        mat3_mul(dR, X, sk);
        const double wgt = cfg.cart_task_weight[lane];
        double o3 = sk[7], o4 = sk[2], o5 = sk[3];                      // skew[2,1], skew[0,2], skew[1,0]
        if (is_trunk) {                                                 // quaternion error (:976), z term: the w*z products cancel
          const double* f = fkq; const double* r = qref;
          o3 += cfg.trunk_gain_ori[0] * ((f[3] * r[0]) - (f[0] * r[3]) + (f[1] * r[2]) - (f[2] * r[1]));
          o4 += cfg.trunk_gain_ori[1] * ((f[3] * r[1]) - (f[1] * r[3]) - (f[0] * r[2]) + (f[2] * r[0]));
          o5 += cfg.trunk_gain_ori[2] * ((f[0] * r[1]) - (f[1] * r[0]));
        }
        b6[0] = (ref_vel[0] + ge[0]) * wgt;
        b6[1] = (ref_vel[1] + ge[1]) * wgt;
        b6[2] = (ref_vel[2] + ge[2]) * wgt;
        b6[3] = o3 * wgt; b6[4] = o4 * wgt; b6[5] = o5 * wgt;
#pragma unroll
        for (int k = 0; k < 3; ++k) sts_f64(prev_a + 8 * k, target[k]);   // :995 / :1151
#pragma unroll
        for (int k = 0; k < 9; ++k) sts_f64(prevR_a + 8 * k, Rref[k]);    // :996 / :1152
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) sts_f64(bs_a + 8 * (6 * lane + r), b6[r]);
    }
    if (row_trunk >= 0) {                                              // trunkConstraint (:707-754), warp level
      double fq[4], eul[3], ip[3], ie[3];
#pragma unroll
      for (int k = 0; k < 4; ++k) fq[k] = __shfl_sync(WBC_FULL_MASK, fkq[k], WBC_FRAME_TRUNK);
      warp_scipy_euler_xyz_from_quat(fq, lane, eul);
      lds_vec3(ref_a + 8 * REF_INIT_TRUNK_POS, ip);
      lds_vec3(ref_a + 8 * REF_INIT_TRUNK_EUL, ie);
      const double cur[4] = {lds_f64(omf_a + 8 * (WBC_T_STRIDE * WBC_FRAME_TRUNK) + 88), eul[0], eul[1], eul[2]};
      const double z_var = ip[2] * 0.25;
      const double var = 1.5 * 0.1;
      const double lo[4] = {ip[2] - z_var, ie[0] - var, ie[1] - var, ie[2] - var};
      const double up[4] = {ip[2] + z_var, ie[0] + var, ie[1] + var, ie[2] + var};
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          sts_f64(clb_a + 8 * (row_trunk + r), ((lo[r] - cur[r]) * inv_dt) * 0.5);
          sts_f64(cub_a + 8 * (row_trunk + r), ((up[r] - cur[r]) * inv_dt) * 0.5);
        }
      }
    }
    if (row_com >= 0 && lane == 0) {                               // CoMConstraint (:669-694)
      double FL[3], RR[3];
      lds_vec3(omf_a + 8 * (1 * WBC_T_STRIDE + 9), FL);
      lds_vec3(omf_a + 8 * (2 * WBC_T_STRIDE + 9), RR);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        sts_f64(clb_a + 8 * (row_com + r), ((RR[r] - com_w[r]) * inv_dt) * 0.8);
        sts_f64(cub_a + 8 * (row_com + r), ((FL[r] - com_w[r]) * inv_dt) * 0.8);
      }
    }
    if (lane < cfg.n_extra_rows) {
      sts_f64(clb_a + 8 * (row_extra + lane), cfg.extra_lo[lane]);
      sts_f64(cub_a + 8 * (row_extra + lane), cfg.extra_hi[lane]);
    }

    // ---------------------------------------------------------------- qpJointb "MANI" / "HYBRID" (:1219-1260)
    // Central finite differences of the manipulability sqrt(det(J J^T)) of one joint Jacobian per probed DoF: two
    // extra FK passes each.  Reference quirks kept: the configuration is indexed with the velocity
    // index, the perturbations accumulate (each probed entry ends at q - dq), and the perturbed state is what the
    // constraint rows, the bounds and the integration that follow see (A and the Cartesian targets were taken before).
    double u_fd = 0.0;
    const bool fd_on = FD && joint_on && (cfg.joint_mode == WBC_JOINT_MANI || cfg.joint_mode == WBC_JOINT_HYBRID);
    if (FD && fd_on) {                                               // FD: separate kernel instantiation, so the
                                                                     // common modes do not carry its register pressure
      const bool mani = cfg.joint_mode == WBC_JOINT_MANI;
      const uint32_t G_a = ast_a, Mm_a = ast_a + 8 * 192, qp_a = ast_a + 8 * 232;   // AsT is not written yet
      const double deltaq = 0.0002;
      for (int i = lane; i < nq; i += 32) sts_f64(qp_a + 8 * i, lds_f64(q_a + 8 * i));
      if (!mani && lane < NV) u_fd = lds_f64(q_a + 8 * ((lane < 6) ? lane : lane + 1));   // u = np.delete(q, 6)
      __syncwarp();
      bool any = false;
#pragma unroll 1
      for (int i = 0; i < NV; ++i) {
        const int jid = mani ? (i < 6 ? 1 : i + 1 - 5) : (i - 6);
        if (!mani && jid < cfg.arm_base_id) continue;
        any = true;
        double f1 = 0.0, f2 = 0.0;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
          if (lane == 0) {
            const double qi = lds_f64(qp_a + 8 * i);
            sts_f64(qp_a + 8 * i, side == 0 ? qi + deltaq : qi - (deltaq * 2));
          }
          __syncwarp();
          warp_fk_a(M_a, qp_a, omi_a, lane);
          const double f = warp_manip_a<NV>(M_a, omi_a, G_a, Mm_a, jid, lane);
          if (side == 0) f1 = f; else f2 = f;
        }
        if (lane == i) u_fd = 0.5 * (f1 - f2) / deltaq;
      }
      if (any) {          // current_joint_config is the perturbed array from here on: refresh what depends on it
        for (int i = lane; i < nq; i += 32) sts_f64(q_a + 8 * i, lds_f64(qp_a + 8 * i));
        __syncwarp();
        warp_kin_a<NV>(M_a, q_a, omi_a, omf_a, lane, (cfg.constraint_mask & WBC_CON_COM) != 0, Sc, com_w, Jcom);
        if (row_trunk >= 0 && lane == WBC_FRAME_TRUNK) {             // trunkConstraint (:707-754) at the perturbed state
          const uint32_t T_a = omf_a + 8 * WBC_T_STRIDE * WBC_FRAME_TRUNK;
          double Rf[9], fq[4], eul[3], ip[3], ie[3];
          lds_mat3(T_a, Rf);
          scipy_quat_from_matrix(Rf, fq);
          scipy_euler_xyz_from_quat(fq, eul);
          lds_vec3(ref_a + 8 * REF_INIT_TRUNK_POS, ip);
          lds_vec3(ref_a + 8 * REF_INIT_TRUNK_EUL, ie);
          const double cur[4] = {lds_f64(T_a + 88), eul[0], eul[1], eul[2]};
          const double z_var = ip[2] * 0.25;
          const double var = 1.5 * 0.1;
          const double lo[4] = {ip[2] - z_var, ie[0] - var, ie[1] - var, ie[2] - var};
          const double up[4] = {ip[2] + z_var, ie[0] + var, ie[1] + var, ie[2] + var};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            sts_f64(clb_a + 8 * (row_trunk + r), ((lo[r] - cur[r]) * inv_dt) * 0.5);
            sts_f64(cub_a + 8 * (row_trunk + r), ((up[r] - cur[r]) * inv_dt) * 0.5);
          }
        }
        if (row_com >= 0 && lane == 0) {
          double FL[3], RR[3];
          lds_vec3(omf_a + 8 * (1 * WBC_T_STRIDE + 9), FL);
          lds_vec3(omf_a + 8 * (2 * WBC_T_STRIDE + 9), RR);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            sts_f64(clb_a + 8 * (row_com + r), ((RR[r] - com_w[r]) * inv_dt) * 0.8);
            sts_f64(cub_a + 8 * (row_com + r), ((FL[r] - com_w[r]) * inv_dt) * 0.8);
          }
        }
        __syncwarp();
      }
    }

    // ---------------------------------------------------------------- box bounds, joint task
    double lbv = 0.0, ubv = 0.0;
    if (lane < NV) damper_bounds_a(M_a, cfg, q_a, lane, lbv, ubv);
    double bj = 0.0;
    if (joint_on && cfg.joint_mode == WBC_JOINT_PREV && lane < NV)               // qpJointb "PREV" (:1216-1217)
      bj = ((1.0 / NV) * lds_f64(q_a + 8 * ((lane < 6) ? lane : lane + 1))) * cfg.joint_task_weight;
    if (fd_on && lane < NV) bj = ((1.0 / NV) * u_fd) * cfg.joint_task_weight;

    // ---------------------------------------------------------------- constraint rows
    // (general instantiation: into the block of the transposed task rows once those are dead; reduced instantiation: that
    //  block only ever holds C, so the rows are written before the task rows are computed -- shorter live range of `a`)
    auto constraint_rows = [&]() {
    if (lane < NV) {
      const uint32_t c_a = ast_a + 8 * lane;
      if (row_com >= 0) { sts_f64(c_a + 8 * LD * row_com, Jcom[0]); sts_f64(c_a + 8 * LD * (row_com + 1), Jcom[1]); }
      if (row_trunk >= 0) {                                        // LWA rows z, wx, wy, wz of the trunk frame (:709)
        double Jc[6], fp[3];
        lds_vec3(omf_a + 8 * (WBC_T_STRIDE * WBC_FRAME_TRUNK + 9), fp);
        frame_jac_lwa(Sc, (supp[WBC_FRAME_TRUNK] >> lane) & 1u, fp, false, Jc);
#pragma unroll
        for (int r = 0; r < 4; ++r) sts_f64(c_a + 8 * LD * (row_trunk + r), Jc[2 + r]);
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (row_ee[i] >= 0) {                                      // WORLD linear rows (:758)
          const bool sup = (supp[i] >> lane) & 1u;
#pragma unroll
          for (int r = 0; r < 3; ++r) sts_f64(c_a + 8 * LD * (row_ee[i] + r), sup ? Sc[r] : 0.0);
        }
      }
      for (int e = 0; e < cfg.n_extra_rows; ++e) {                 // extension rows (not in the reference)
        const int f = cfg.extra_frame[e];
        double T[12], Jc[6];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = lds_f64(omf_a + 8 * (WBC_T_STRIDE * f + i));
        frame_jac_column(Sc, (uint32_t)lds_s32(M_a + WBC_MOFF(frame_supp) + 4 * f), lane, T, cfg.extra_rf[e], Jc);
        double sacc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) sacc += cfg.extra_coeff[e][c] * Jc[c];
        sts_f64(c_a + 8 * LD * (row_extra + e), sacc);
      }
    }
    };
    if (RED) constraint_rows();
    if (!FD) task_rows();

    // ---------------------------------------------------------------- H = A^T A, g = -A^T b
    // Column `lane` of the 36 Cartesian rows goes to shared memory transposed (AsT[lane][r]); then for every
    // (task t, supporting column l) pair lane i adds  sum_r A[6t+r][i] A[6t+r][l]  to H[i][l]: three 128-bit
    // broadcast loads and six FMAs.  Pairs outside the frame's support are structural zeros and are skipped.
    __syncwarp();                      // all reads of oMi (aliased by Hs) and bs writes are done
    const uint32_t hrow_a = hs_a + 8 * LD * (lane < NV ? lane : 0);
    double gk = 0.0;
    if (!RED) {                        // (RED: the solver's front assembles the reduced Hessian from `a` itself)
    if (lane < NV) {
      const uint32_t da = ast_a + 8 * WBC_LDT * lane;
#pragma unroll
      for (int r = 0; r < 18; ++r) sts_f64x2(da + 16 * r, a[2 * r], a[2 * r + 1]);
#pragma unroll
      for (int l = 0; l < NV; ++l) sts_f64(hrow_a + 8 * l, 0.0);
    }
    __syncwarp();
    {
      double g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int r = 0; r < 18; ++r) {
        const double2 b2 = lds_f64x2(bs_a + 16 * r);
        g0 = fma(-a[2 * r], b2.x, g0);
        g1 = fma(-a[2 * r + 1], b2.y, g1);
      }
      gk = (g0 + g1) - aj * bj;
    }
    // The six free-flyer columns support every frame: 36 of the 53 (task, column) pairs.  They accumulate in
    // registers as independent chains; only the limb columns go through the shared-memory read-modify-write.
    double hb[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      if (!((cfg.task_mask >> t) & 1)) continue;
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        if ((supp[t] >> l) & 1u) {                                 // warp-uniform
          const uint32_t ca = ast_a + 8 * (WBC_LDT * l + 6 * t);
          const double2 v0 = lds_f64x2(ca), v1 = lds_f64x2(ca + 16), v2 = lds_f64x2(ca + 32);
          double h0 = fma(a[6 * t], v0.x, hb[l]), h1 = a[6 * t + 1] * v0.y;
          h0 = fma(a[6 * t + 2], v1.x, h0); h1 = fma(a[6 * t + 3], v1.y, h1);
          h0 = fma(a[6 * t + 4], v2.x, h0); h1 = fma(a[6 * t + 5], v2.y, h1);
          hb[l] = h0 + h1;
        }
      }
      uint32_t mask = supp[t] & ~0x3Fu;
      while (mask) {                                               // warp-uniform
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t ca = ast_a + 8 * (WBC_LDT * l + 6 * t);
        const double2 v0 = lds_f64x2(ca), v1 = lds_f64x2(ca + 16), v2 = lds_f64x2(ca + 32);
        const uint32_t ha = hrow_a + 8 * l;
        const double hprev = lds_f64(ha);
        double h0 = fma(a[6 * t], v0.x, hprev), h1 = a[6 * t + 1] * v0.y;
        h0 = fma(a[6 * t + 2], v1.x, h0); h1 = fma(a[6 * t + 3], v1.y, h1);
        h0 = fma(a[6 * t + 4], v2.x, h0); h1 = fma(a[6 * t + 5], v2.y, h1);
        sts_f64_if(lane < NV, ha, h0 + h1);
      }
    }
    if (lane < NV) {
#pragma unroll
      for (int l = 0; l < 6; ++l) sts_f64(hrow_a + 8 * l, hb[l]);
    }
    __syncwarp();
    }   // !RED

    if (DEBUG_OUT) {
      const WbcAssembleOut& D = P.dbg;
      const int m = P.m_rows;
      if (lane < NV) {
        int row = 0;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          if (!((cfg.task_mask >> t) & 1)) continue;
#pragma unroll
          for (int r = 0; r < 6; ++r, ++row) {
            if (D.A) D.A[(sidx * m + row) * NV + lane] = a[6 * t + r];
            if (D.b && lane == 0) D.b[sidx * m + row] = lds_f64(bs_a + 8 * (6 * t + r));
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
          for (int l = 0; l < NV; ++l)
            D.H[(sidx * NV + lane) * NV + l] = lds_f64(hrow_a + 8 * l) + ((l == lane) ? aj * aj : 0.0);
      }
      __syncwarp();
    }

    if (!RED) constraint_rows();
    phase_sync<PSF && (WBC_SYNC_PREQP != 0)>();
    const double clb_r = (DEBUG_OUT && lane < nC) ? lds_f64(clb_a + 8 * lane) : 0.0;
    const double cub_r = (DEBUG_OUT && lane < nC) ? lds_f64(cub_a + 8 * lane) : 0.0;

    if (DEBUG_OUT) {
      const WbcAssembleOut& D = P.dbg;
      if (D.C && lane < NV)
        for (int r = 0; r < nC; ++r) D.C[(sidx * nC + r) * NV + lane] = lds_f64(ast_a + 8 * (LD * r + lane));
      if (D.Clb && lane < nC) D.Clb[sidx * nC + lane] = clb_r;
      if (D.Cub && lane < nC) D.Cub[sidx * nC + lane] = cub_r;
      if (P.io.mem_out)
        for (int i = lane; i < WBC_MEM_STRIDE; i += 32) P.io.mem_out[sidx * WBC_MEM_STRIDE + i] = lds_f64(mem_a + 8 * i);
      __syncwarp();
      buf ^= 1;
      continue;
    }

    // ---------------------------------------------------------------- the QP
    // The loop state (state index, input buffer, mbarrier phases, tick) is parked in two spare shared-memory words across
    // the QP and re-derived behind it: carried in registers it is spilled to local memory there anyway (the QP is the
    // kernel's region of highest register pressure), and local loads miss the 28 KB of L1 half of the time.
    const uint32_t stash_a = in0_a + 8 * 33, stash_b = in0_a + 8 * (WBC_IN_TOTAL + 33);
    if (lane == 0) {
      sts_s32(stash_a, base);
      sts_s32(stash_a + 4, buf | (int)(mb_phase << 1));
      if (MULTI) sts_s32(stash_b, tick);
    }
    double x;
    QpResult res;
    {
      double h[NV];
#pragma unroll
      for (int l = 0; l < NV; ++l)
        h[l] = (!RED && lane < NV - NF) ? lds_f64(hrow_a + 8 * l) + ((l == lane) ? aj * aj : 0.0) : 0.0;
      const double hdiag = (!RED && lane < NV) ? lds_f64(hrow_a + 8 * lane) + aj * aj : 0.0;
      __syncwarp();                    // Hs becomes the solver's R factor
      QpRegShared S;
      S.R = hs_a; S.col = ws_a + 8 * L.col; S.vd = vd_a; S.C = ast_a;
      S.clb = clb_a; S.cub = cub_a; S.dd = bs_a;
      S.red_rows = P.red_rows; S.feet_mask = P.red_feet_mask; S.red_blk = P.red_blk; S.red_other = P.red_other;
      S.b = bs_a; S.skip_act = P.io.active_set == nullptr;
      res = warp_qp_solve_reg_impl<NV, SPLIT, PS && (WBC_QP_MID_SYNC != 0), NF, RED>(S, h, hdiag, nC, gk, lbv, ubv, cfg.max_iter,
                                                                                    x, a, aj, bj);
    }

    {
      const int bm = lds_s32(stash_a + 4);
      base = lds_s32(stash_a);
      buf = bm & 1;
      mb_phase = (uint32_t)bm >> 1;
      if (MULTI) { tick = lds_s32(stash_b); toff = (long long)tick * NS; }
    }
    constexpr bool GRP = RED && PS && (WBC_SYNC_POSTQP != 0) && (WBC_SYNC_GROUP < 0);
    const uint32_t grp_a = ws_a + 8 * L.total * (wpc - warp) + 128 * buf;      // behind the last warp's workspace; slots alternate
    if (GRP && P.group_report && lane == 0) {
      sts_s32(grp_a + 4 * warp, res.status);
      sts_s32(grp_a + 256 + 4 * warp, res.iters);
    }
    phase_sync<PS && (WBC_SYNC_POSTQP != 0)>();
    {   // (scope of the re-derived loop state: it shadows what the first half of the tick used)
    int sidx_p = base + warp;
    const bool valid = sidx_p < NS;
    if (!valid) sidx_p = NS - 1;
    const long long sidx = sidx_p;
    const uint32_t in_a = in0_a + 8 * WBC_IN_TOTAL * buf;
    const uint32_t q_a = in_a + 8 * WBC_IN_Q, tg_a = in_a + 8 * WBC_IN_TARGETS, mem_a = in_a + 8 * WBC_IN_MEM;
    if (valid && lane < NV) {
      if (P.f32_out & WBC_F32_QDOT) reinterpret_cast<float*>(P.io.qdot)[sidx * NV + lane] = (float)x;
      else P.io.qdot[sidx * NV + lane] = x;
    }
    if (GRP && P.group_report) {
      const int half = (wpc + 1) >> 1;
      const int g0 = warp >= half ? half : 0, gn = warp >= half ? wpc - half : half;
      const int sg = base + g0 + lane;
      if (warp == g0 && lane < gn && sg < NS) {
        P.io.status[sg] = lds_s32(grp_a + 4 * (g0 + lane));
        P.io.iters[sg] = lds_s32(grp_a + 256 + 4 * (g0 + lane));
      }
    }
    if (valid && lane == 0) {
      if (!(GRP && P.group_report)) {
        P.io.status[sidx] = res.status;
        P.io.iters[sidx] = res.iters;
      }
      if (P.io.active_set) {
        P.io.active_set[2 * sidx] = res.act_box;
        P.io.active_set[2 * sidx + 1] = res.act_rows;
      }
    }
    if (valid && P.io.mem_out)
      for (int i = lane; i < WBC_MEM_STRIDE; i += 32) P.io.mem_out[sidx * WBC_MEM_STRIDE + i] = lds_f64(mem_a + 8 * i);

    // ---------------------------------------------------------------- integrate + base estimate
    if (P.io.q_next) {
      __syncwarp();
      const double v = x * dt;
      const uint32_t qn_a = ws_a + 8 * L.col;          // [<= 33] new configuration (the 64-double column buffer is free now)
      // (runWBC with IMU feedback keeps nothing of the integrated free-flyer pose: updateState(running=True) puts the old xyz
      //  and the IMU quaternion in its place and re-estimates xyz (:387-428) -- ~250 serial instructions of lane 0 skipped)
      const bool need_ff = (P.flags & WBC_STEP_FLAG_PLAIN_INTEGRATE) || !P.io.imu_quat;
      if (need_ff) {                                   // (uniform)
        double vb[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) vb[i] = __shfl_sync(WBC_FULL_MASK, v, i);
        if (lane == 0) {
          double q7[7], o7[7];
#pragma unroll
          for (int i = 0; i < 7; ++i) q7[i] = lds_f64(q_a + 8 * i);
          integrate_freeflyer(q7, vb, o7);
#pragma unroll
          for (int i = 0; i < 7; ++i) sts_f64(qn_a + 8 * i, o7[i]);
        }
      }
      if (lane >= 6 && lane < NV) {
        const int iq = lds_s32(M_a + WBC_MOFF(col_q) + 4 * lane);
        sts_f64(qn_a + 8 * iq, lds_f64(q_a + 8 * iq) + v);
      }
      __syncwarp();
      if (!(P.flags & WBC_STEP_FLAG_PLAIN_INTEGRATE)) {
        // updateState(joint_config, imu, running=True): q = [old xyz, imu quat, joints], FK, trunkWorldPos (:387-428)
        if (lane < 3) sts_f64(qn_a + 8 * lane, lds_f64(q_a + 8 * lane));
        if (P.io.imu_quat && lane < 4) sts_f64(qn_a + 8 * (3 + lane), lds_f64(in_a + 8 * (WBC_IN_IMU + lane)));
        __syncwarp();
        warp_fk_a(M_a, qn_a, omi_a, lane);
        double bp[3] = {0, 0, 0};
        if (lane < 4) {                // foot frame positions at the new configuration
          const uint32_t Pa = omi_a + 8 * WBC_T_STRIDE * lds_s32(M_a + WBC_MOFF(frame_parent) + 4 * lane);
          double Rp[9], fp[3];
          lds_mat3(Pa, Rp);
          lds_vec3(M_a + WBC_MOFF(frp) + 24 * lane, fp);
          mat3_vec(Rp, fp, bp);
          bp[0] += lds_f64(Pa + 72); bp[1] += lds_f64(Pa + 80); bp[2] += lds_f64(Pa + 88);
        }
        // trunkWorldPos (:1297-1327): order of the sums follows the reference (FR + FL + RR + RL) / 4
        double BPA[3], WPA[3], Rt[9], pt[3];
        {
          const uint32_t Ta = omi_a + 8 * WBC_T_STRIDE * lds_s32(M_a + WBC_MOFF(frame_parent) + 4 * WBC_FRAME_TRUNK);
          double Rp[9], FR[9], fp[3];
          lds_mat3(Ta, Rp);
          lds_mat3(M_a + WBC_MOFF(frR) + 72 * WBC_FRAME_TRUNK, FR);
          lds_vec3(M_a + WBC_MOFF(frp) + 24 * WBC_FRAME_TRUNK, fp);
          mat3_mul(Rp, FR, Rt);
          mat3_vec(Rp, fp, pt);
          pt[0] += lds_f64(Ta + 72); pt[1] += lds_f64(Ta + 80); pt[2] += lds_f64(Ta + 88);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double d0 = __shfl_sync(WBC_FULL_MASK, bp[c], 0) - pt[c];
          const double d1 = __shfl_sync(WBC_FULL_MASK, bp[c], 1) - pt[c];
          const double d2 = __shfl_sync(WBC_FULL_MASK, bp[c], 2) - pt[c];
          const double d3 = __shfl_sync(WBC_FULL_MASK, bp[c], 3) - pt[c];
          BPA[c] = (d0 + d1 + d2 + d3) / 4;
          WPA[c] = (lds_f64(tg_a + 8 * c) + lds_f64(tg_a + 8 * (3 + c)) + lds_f64(tg_a + 8 * (6 + c)) + lds_f64(tg_a + 8 * (9 + c))) / 4;
        }
        double rb[3];
        mat3_vec(Rt, BPA, rb);
        __syncwarp();
        if (lane < 3) sts_f64(qn_a + 8 * lane, WPA[lane] - rb[lane]);
        __syncwarp();
      }
      if (valid)
        for (int i = lane; i < nq; i += 32) P.io.q_next[sidx * nq + i] = lds_f64(qn_a + 8 * i);
      if (valid && P.io.joint_targets && lane < nq - 7) {      // what runWBC returns: q_next[7:] (:1405-1412)
        const double jt = lds_f64(qn_a + 8 * (7 + lane));
        if (P.f32_out & WBC_F32_JOINTS) reinterpret_cast<float*>(P.io.joint_targets)[sidx * (nq - 7) + lane] = (float)jt;
        else P.io.joint_targets[sidx * (nq - 7) + lane] = jt;
      }
      __syncwarp();
    }
    }   // re-derived loop state
    if (MULTI && late) {               // single round per tick: the state just written is the next one to read
      __syncwarp();
      prefetch_next();
    }
    buf ^= 1;
  }
  cp_async_wait_all();
#undef row_com
#undef row_trunk
#undef row_ee
#undef row_extra
}
