// wbc_kernels.cu -- kernels + C ABI of libwbc_b200.so (see include/wbc_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>
#include "wbc_step.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
#define CUDA_TRY(expr)                                                          \
  do {                                                                          \
    cudaError_t e__ = (expr);                                                   \
    if (e__ != cudaSuccess) return fail(WBC_ERR_CUDA, #expr ": %s", cudaGetErrorString(e__)); \
  } while (0)

#define WBC_PIPE_STREAMS 3
struct WbcModel {
  DevModel host;
  DevModel* dev;
  int device;
  int sm_count;
  // wbc_step_host: copy / compute pipeline (created on first use)
  bool pipe_ready;
  cudaStream_t pipe[WBC_PIPE_STREAMS];
  cudaEvent_t pipe_start, pipe_done[WBC_PIPE_STREAMS];
  // wbc_step_host, chunks == 0: which host path is faster HERE is measured, not assumed (one GPU alone: zero-copy; eight GPUs
  // pulling on one host NUMA node: staged slices).  Calls 0-1 of a new problem shape run zero-copy, calls 2-3 staged; calls 1
  // and 3 are timed with events on the caller's stream; from then on the faster one runs.
  struct HostTune {
    int64_t N; int key;          // problem shape the measurement belongs to (states; dtype / form / flags)
    int calls;                   // calls seen with this shape
    int choice;                  // -1 undecided, 0 zero-copy, 8 staged slices
    bool ev_ready;
    cudaEvent_t ev[4];           // [zero-copy begin, end, staged begin, end]
  } tune;
};
#define WBC_HOST_AUTO_STAGED 8

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_model(const DevModel* __restrict__ g, DevModel* s) {
  const int nwords = sizeof(DevModel) / 4;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
  uint32_t* dst = reinterpret_cast<uint32_t*>(s);
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// fused tick / assembly accessor: persistent CTAs, one state per warp.
// SPLIT (nC <= 16): the QP keeps each row's d vector in two lanes, fits 168 registers -> 12 warps per SM;
// otherwise the full-width layout needs ~210 registers -> 8 warps per SM.
#ifndef WBC_STEP_WARPS
#define WBC_STEP_WARPS 12
#endif
#ifndef WBC_STEP_CTAS
#define WBC_STEP_CTAS 1       // CTAs per SM of the SPLIT kernel (WBC_STEP_WARPS warps each)
#endif
#ifndef WBC_STEP_WARPS_RED
#define WBC_STEP_WARPS_RED 16 // the reduced-front instantiation: 128 registers, 13.5 KB of shared memory per warp
#endif
#ifndef WBC_STEP_WARPS_FULL
#define WBC_STEP_WARPS_FULL 12 // the full-width solver layout (more than 16 rows of C): 168 registers with ~0.8 KB of spills, yet 14 % faster than 8 warps at 255 (config 3)
#endif
template <bool SPLIT, bool RED = false> struct StepWarps {
  static constexpr int value = RED ? WBC_STEP_WARPS_RED : (SPLIT ? WBC_STEP_WARPS : WBC_STEP_WARPS_FULL);
  static constexpr int ctas = SPLIT ? WBC_STEP_CTAS : 1;
};

#ifdef WBC_STEP_MAXNREG       // A/B builds: explicit register cap instead of the launch-bounds heuristic
#define WBC_STEP_BOUNDS(SPLIT, RED) __maxnreg__(WBC_STEP_MAXNREG)
#else
#define WBC_STEP_BOUNDS(SPLIT, RED) __launch_bounds__(32 * StepWarps<SPLIT, RED>::value, StepWarps<SPLIT, RED>::ctas)
#endif
template <int NV, bool DEBUG_OUT, bool SPLIT, bool FD, int NF, bool RED, bool MULTI>
__global__ void WBC_STEP_BOUNDS(SPLIT, RED) wbc_step_kernel(const __grid_constant__ StepParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevModel* Ms = reinterpret_cast<DevModel*>(smem_raw);
  stage_model(P.model, Ms);
  constexpr StepLayout L = step_layout(NV, WBC_MAX_NC, RED);
  const int warp = SPLIT ? (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5) : (int)(threadIdx.x >> 5);   // warp-uniform for ptxas (see warp_wbc_states)
  double* ws = reinterpret_cast<double*>(smem_raw + ((sizeof(DevModel) + 15) & ~size_t(15))) + (size_t)warp * L.total;
  warp_wbc_states<NV, DEBUG_OUT, SPLIT, FD, NF, RED, MULTI>(P, Ms, ws);
}

// FK + frame Jacobians accessor (HBM-write bound): one state per warp
struct FkJacParams {
  const DevModel* model;
  const double* q;
  long long N;
  int nsel, rf;
  int slots[WBC_MAX_FRAMES];
  double* out_oMf;
  double* out_J;
  double* out_oMi;   // joint_jacobians variant
  double* out_Jw;
};

#ifndef WBC_FKJ_CTAS
#define WBC_FKJ_CTAS 2      // 128 registers: 16 warps per SM (3 CTAs spill, measured slower)
#endif
__global__ void __launch_bounds__(256, WBC_FKJ_CTAS) wbc_fk_jac_kernel(const __grid_constant__ FkJacParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevModel* Ms = reinterpret_cast<DevModel*>(smem_raw);
  stage_model(P.model, Ms);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int per_warp = WBC_MAX_JOINTS * WBC_T_STRIDE + WBC_MAX_FRAMES * WBC_T_STRIDE + 40;
  double* ws = reinterpret_cast<double*>(smem_raw + ((sizeof(DevModel) + 15) & ~size_t(15))) + (size_t)warp * per_warp;
  const uint32_t M_a = smem_addr(Ms);
  const uint32_t omi_a = smem_addr(ws);
  const uint32_t omf_a = omi_a + 8 * WBC_MAX_JOINTS * WBC_T_STRIDE;      // placement of selected frame f at slot f
  const uint32_t q_a = omf_a + 8 * WBC_MAX_FRAMES * WBC_T_STRIDE;
  const int nq = Ms->nq, nv = Ms->nv, nj = Ms->njoints;
  // per selected frame (lane f < nsel): parent joint, support mask, offset -- loop invariant
  int fpar = 0;
  uint32_t fsupp_mine = 0;
  double frR[9], frp[3];
  bool fident = true;
  if (lane < P.nsel) {
    const int slot = P.slots[lane];
    fpar = Ms->frame_parent[slot];
    fsupp_mine = Ms->frame_supp[slot];
    fident = Ms->fr_ident[slot] != 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) frR[i] = Ms->frR[slot][i];
    frp[0] = Ms->frp[slot][0]; frp[1] = Ms->frp[slot][1]; frp[2] = Ms->frp[slot][2];
  }
  for (long long s = (long long)blockIdx.x * wpc + warp; s < P.N; s += (long long)gridDim.x * wpc) {
    for (int i = lane; i < nq; i += 32) sts_f64(q_a + 8 * i, P.q[s * nq + i]);
    __syncwarp();
    warp_fk_a(M_a, q_a, omi_a, lane);
    double Sc[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) Sc[i] = 0.0;
    if (lane < nv) {                                  // column `lane` of data.J (run-time nv: same code as warp_jac_col_a)
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
    if (lane < P.nsel) {                              // updateFramePlacements for the selected frames
      double Rp[9], R[9], p[3], pp[3] = {0.0, 0.0, 0.0};
      if (fpar == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) Rp[i] = (i % 4 == 0) ? 1.0 : 0.0;
      } else {
        lds_mat3(omi_a + 8 * WBC_T_STRIDE * fpar, Rp);
        lds_vec3(omi_a + 8 * (WBC_T_STRIDE * fpar + 9), pp);
      }
      if (fident) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = Rp[i];
      } else {
        mat3_mul(Rp, frR, R);
      }
      mat3_vec(Rp, frp, p);
      const uint32_t oa = omf_a + 8 * WBC_T_STRIDE * lane;
#pragma unroll
      for (int i = 0; i < 9; ++i) sts_f64(oa + 8 * i, R[i]);
      sts_f64(oa + 72, p[0] + pp[0]); sts_f64(oa + 80, p[1] + pp[1]); sts_f64(oa + 88, p[2] + pp[2]);
    }
    __syncwarp();
    if (P.out_oMi) {
      for (int i = lane; i < nj * 12; i += 32) {
        const int j = i / 12, e = i % 12;
        double v = lds_f64(omi_a + 8 * (j * WBC_T_STRIDE + e));
        if (j == 0) v = (e == 0 || e == 4 || e == 8) ? 1.0 : 0.0;
        P.out_oMi[s * nj * 12 + i] = v;
      }
    }
    if (P.out_Jw && lane < nv)
      for (int r = 0; r < 6; ++r) P.out_Jw[(s * 6 + r) * nv + lane] = Sc[r];
    if (P.out_oMf)                                    // [N, nsel, 12]: contiguous per state
      for (int i = lane; i < P.nsel * 12; i += 32)
        P.out_oMf[s * P.nsel * 12 + i] = lds_f64(omf_a + 8 * ((i / 12) * WBC_T_STRIDE + (i % 12)));
    if (P.out_J) {
      double* Jout = P.out_J + s * (long long)P.nsel * 6 * nv + lane;
      for (int f = 0; f < P.nsel; ++f) {
        const uint32_t supp = __shfl_sync(WBC_FULL_MASK, fsupp_mine, f);
        double T[12], Jc[6];
        const uint32_t ta = omf_a + 8 * WBC_T_STRIDE * f;
        if (P.rf == WBC_RF_LOCAL) {
#pragma unroll
          for (int i = 0; i < 9; ++i) T[i] = lds_f64(ta + 8 * i);
        }
        T[9] = lds_f64(ta + 72); T[10] = lds_f64(ta + 80); T[11] = lds_f64(ta + 88);
        frame_jac_column(Sc, supp, lane, T, P.rf, Jc);
        if (lane < nv) {
#pragma unroll
          for (int r = 0; r < 6; ++r) Jout[(f * 6 + r) * nv] = Jc[r];
        }
      }
    }
    __syncwarp();
  }
}

// initialiseWBC snapshot (Robot_Wrapper4.py:354-383)
__global__ void __launch_bounds__(256) wbc_init_memory_kernel(const DevModel* __restrict__ model, const double* __restrict__ q,
                                                              long long N, double* __restrict__ mem, double* __restrict__ ref) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevModel* Ms = reinterpret_cast<DevModel*>(smem_raw);
  stage_model(model, Ms);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int per_warp = WBC_MAX_JOINTS * WBC_T_STRIDE + WBC_MAX_FRAMES * WBC_T_STRIDE + 40;
  double* ws = reinterpret_cast<double*>(smem_raw + ((sizeof(DevModel) + 15) & ~size_t(15))) + (size_t)warp * per_warp;
  double* oMi = ws;
  double* oMf = ws + WBC_MAX_JOINTS * WBC_T_STRIDE;
  double* qs = oMf + WBC_MAX_FRAMES * WBC_T_STRIDE;
  const int nq = Ms->nq;
  for (long long s = (long long)blockIdx.x * wpc + warp; s < N; s += (long long)gridDim.x * wpc) {
    for (int i = lane; i < nq; i += 32) qs[i] = q[s * nq + i];
    __syncwarp();
    warp_fk(Ms, qs, oMi, lane);
    warp_frames(Ms, oMi, oMf, lane);
    double* m = mem + s * WBC_MEM_STRIDE;
    double* r = ref + s * WBC_REF_STRIDE;
    const double* Tt = oMf + WBC_FRAME_TRUNK * WBC_T_STRIDE;
    if (lane < 6) {
      const double* T = oMf + lane * WBC_T_STRIDE;
      double Rf[9], qf[4], eul[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rf[i] = T[i];
      scipy_quat_from_matrix(Rf, qf);
      scipy_euler_xyz_from_quat(qf, eul);
      if (lane < 5) {
        for (int i = 0; i < 3; ++i) {
          m[MEM_PREV_EE_POS + 3 * lane + i] = T[9 + i];              // prev_EE_pos (:373)
          r[REF_DEF_EE_ORI + 3 * lane + i] = eul[i];                 // default_EE_ori_list (:365-367)
        }
        double Rt[9], Rrel[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rt[i] = Tt[i];
        // prev_EE_CoM_rot = trunk_R^T * EE_R (:374-376)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) Rrel[3 * i + j] = Rt[i] * Rf[j] + Rt[3 + i] * Rf[3 + j] + Rt[6 + i] * Rf[6 + j];
        for (int i = 0; i < 9; ++i) m[MEM_PREV_EE_ROT + 9 * lane + i] = Rrel[i];
      } else {
        for (int i = 0; i < 3; ++i) {
          m[MEM_PREV_TRUNK_REF + i] = T[9 + i];                      // prev_trunk_ref (:370)
          r[REF_DEF_TRUNK_ORI + i] = eul[i];                         // default_trunk_ori (:363-364)
          r[REF_INIT_TRUNK_POS + i] = T[9 + i];                      // initial_trunk_pos (:379)
          r[REF_INIT_TRUNK_EUL + i] = eul[i];                        // initial_trunk_ori_euler (:381-383)
        }
        for (int i = 0; i < 9; ++i) m[MEM_OLD_TRUNK_ROT + i] = Rf[i];  // old_ref_trunk_rot_matrix (:371)
      }
    }
    __syncwarp();
  }
}

// updateState(..., running=True) minus the accessor refresh (Robot_Wrapper4.py:387-428): IMU quaternion into q, FK,
// trunkWorldPos (:1297-1327), q_out = [estimated base xyz, quaternion, joints].  One state per warp.
__global__ void __launch_bounds__(256) wbc_base_estimate_kernel(const DevModel* __restrict__ model, const double* q,
                                                                const double* __restrict__ imu, const double* __restrict__ targets,
                                                                long long N, double* q_out, double* __restrict__ base_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevModel* Ms = reinterpret_cast<DevModel*>(smem_raw);
  stage_model(model, Ms);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int per_warp = WBC_MAX_JOINTS * WBC_T_STRIDE + WBC_MAX_FRAMES * WBC_T_STRIDE + 40;
  double* ws = reinterpret_cast<double*>(smem_raw + ((sizeof(DevModel) + 15) & ~size_t(15))) + (size_t)warp * per_warp;
  double* oMi = ws;
  double* oMf = ws + WBC_MAX_JOINTS * WBC_T_STRIDE;
  double* qs = oMf + WBC_MAX_FRAMES * WBC_T_STRIDE;
  const int nq = Ms->nq;
  for (long long s = (long long)blockIdx.x * wpc + warp; s < N; s += (long long)gridDim.x * wpc) {
    for (int i = lane; i < nq; i += 32) qs[i] = (imu && i >= 3 && i < 7) ? imu[s * 4 + (i - 3)] : q[s * nq + i];
    __syncwarp();
    warp_fk(Ms, qs, oMi, lane);
    warp_frames(Ms, oMi, oMf, lane);
    if (lane == 0) {
      const double* Tt = oMf + WBC_FRAME_TRUNK * WBC_T_STRIDE;
      const double* tg = targets + s * WBC_TARGETS_STRIDE;
      double BPA[3], WPA[3], rb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {            // sums in the reference's order: FR + FL + RR + RL
        const double d0 = oMf[0 * WBC_T_STRIDE + 9 + c] - Tt[9 + c], d1 = oMf[1 * WBC_T_STRIDE + 9 + c] - Tt[9 + c];
        const double d2 = oMf[2 * WBC_T_STRIDE + 9 + c] - Tt[9 + c], d3 = oMf[3 * WBC_T_STRIDE + 9 + c] - Tt[9 + c];
        BPA[c] = (d0 + d1 + d2 + d3) / 4;
        WPA[c] = (tg[c] + tg[3 + c] + tg[6 + c] + tg[9 + c]) / 4;
      }
      mat3_vec(Tt, BPA, rb);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        qs[c] = WPA[c] - rb[c];
        if (base_out) base_out[s * 3 + c] = qs[c];
      }
    }
    __syncwarp();
    if (q_out)
      for (int i = lane; i < nq; i += 32) q_out[s * nq + i] = qs[i];
    __syncwarp();
  }
}

// pin.integrate(model, q, v)  (jointVelocitiestoConfig, Robot_Wrapper4.py:440-441): one state per warp
// (q and out may alias -- wbc_b200.h allows q_out == q -- so neither is __restrict__)
__global__ void __launch_bounds__(256) wbc_integrate_kernel(const DevModel* __restrict__ model, const double* q,
                                                            const double* __restrict__ v, long long N, double scale,
                                                            double* out) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nq = model->nq, nv = model->nv;
  for (long long s = wid; s < N; s += nw) {
    const double* qs = q + s * nq;
    const double* vs = v + s * nv;
    double* o = out + s * nq;
    if (lane == 0) {
      double q7[7], v6[6], o7[7];
      for (int i = 0; i < 7; ++i) q7[i] = qs[i];
      for (int i = 0; i < 6; ++i) v6[i] = vs[i] * scale;
      integrate_freeflyer(q7, v6, o7);
      for (int i = 0; i < 7; ++i) o[i] = o7[i];
    } else if (lane >= 6 && lane < nv) {
      const int iq = model->col_q[lane];
      o[iq] = qs[iq] + vs[lane] * scale;
    }
  }
}

// standalone batched QP: QP(A, b, lb, ub, C, Clb, Cub).solveQP()  (QP_Wrapper.py:10-53)
struct QpParams {
  long long N;
  int nv, m, nC, max_iter;
  const double *A, *b, *H, *g, *lb, *ub, *C, *Clb, *Cub;
  double* x;
  int* status;
  int* iters;
  unsigned long long* active_set;
};

__global__ void __launch_bounds__(256) wbc_qp_kernel(const __grid_constant__ QpParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int n = P.nv, LD = n | 1, nC = P.nC;
  const int per_warp = 2 * n * LD + nC * LD + 3 * 32 + ((2 * n * LD + nC * LD) & 1);
  double* ws = reinterpret_cast<double*>(smem_raw) + (size_t)warp * per_warp;
  QpShared S;
  S.M0 = ws;
  S.J = ws + n * LD;
  S.C = S.J + n * LD;
  S.vx = S.C + nC * LD;
  S.vd = S.vx + 32;
  S.vg = S.vd + 32;
  for (long long s = (long long)blockIdx.x * wpc + warp; s < P.N; s += (long long)gridDim.x * wpc) {
    double gk = 0.0;
    if (P.H) {
      for (int i = lane; i < n * n; i += 32) S.M0[(i / n) * LD + (i % n)] = P.H[s * n * n + i];
      if (lane < n) gk = P.g[s * n + lane];
    } else {
      // H = A^T A, g = -A^T b  (QP_Wrapper.py:17-18): row r of A staged in vx, lane k accumulates row k of H
      for (int i = lane; i < n * LD; i += 32) S.M0[i] = 0.0;
      __syncwarp();
      const double* Ag = P.A + s * (long long)P.m * n;
      const double* bg = P.b + s * (long long)P.m;
      for (int r = 0; r < P.m; ++r) {
        const double ark = (lane < n) ? Ag[r * n + lane] : 0.0;
        S.vx[lane] = ark;
        __syncwarp();
        if (lane < n) {
          double* Hrow = S.M0 + lane * LD;
          for (int l = 0; l < n; ++l) Hrow[l] += ark * S.vx[l];
          gk -= ark * bg[r];
        }
        __syncwarp();
      }
    }
    if (nC > 0)
      for (int i = lane; i < nC * n; i += 32) S.C[(i / n) * LD + (i % n)] = P.C[s * (long long)nC * n + i];
    const double lbv = (lane < n) ? P.lb[s * n + lane] : 0.0, ubv = (lane < n) ? P.ub[s * n + lane] : 0.0;
    const double clb = (lane < nC) ? P.Clb[s * nC + lane] : 0.0, cub = (lane < nC) ? P.Cub[s * nC + lane] : 0.0;
    __syncwarp();
    double x;
    const QpResult res = warp_qp_solve_rt(S, n, LD, nC, gk, lbv, ubv, clb, cub, P.max_iter, x);
    if (lane < n) P.x[s * n + lane] = x;
    if (lane == 0) {
      P.status[s] = res.status;
      P.iters[s] = res.iters;
      if (P.active_set) { P.active_set[2 * s] = res.act_box; P.active_set[2 * s + 1] = res.act_rows; }
    }
    __syncwarp();
  }
}

// the same drop-in with the register-resident solver (compile-time nv; wbc_qp_reg.cuh)
template <int NV, bool SPLIT>
__global__ void __launch_bounds__(256) wbc_qp_reg_kernel(const __grid_constant__ QpParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  constexpr int n = NV, LD = NV | 1;
  const int nC = P.nC;
  const int hs_sz = n * (n + 2) + (n & 1), cs_sz = (nC * LD + 1) & ~1;
  double* ws = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (hs_sz + cs_sz + 224);
  double* Hs = ws;
  double* Cs = ws + hs_sz;
  double* col = Cs + cs_sz;
  double* vd = col + 64;
  double* bnd = vd + 32;         // clb[32] cub[32] dd[64]
  for (long long s = (long long)blockIdx.x * wpc + warp; s < P.N; s += (long long)gridDim.x * wpc) {
    double gk = 0.0;
    if (P.H) {
      for (int i = lane; i < n * n; i += 32) Hs[(i / n) * LD + (i % n)] = P.H[s * n * n + i];
      if (lane < n) gk = P.g[s * n + lane];
    } else {
      for (int i = lane; i < n * LD; i += 32) Hs[i] = 0.0;
      __syncwarp();
      const double* Ag = P.A + s * (long long)P.m * n;
      const double* bg = P.b + s * (long long)P.m;
      for (int r = 0; r < P.m; ++r) {
        const double ark = (lane < n) ? Ag[r * n + lane] : 0.0;
        vd[lane] = ark;
        __syncwarp();
        if (lane < n) {
          double* Hrow = Hs + lane * LD;
          for (int l = 0; l < n; ++l) Hrow[l] += ark * vd[l];
          gk -= ark * bg[r];
        }
        __syncwarp();
      }
    }
    if (nC > 0)
      for (int i = lane; i < nC * n; i += 32) Cs[(i / n) * LD + (i % n)] = P.C[s * (long long)nC * n + i];
    const double lbv = (lane < n) ? P.lb[s * n + lane] : 0.0, ubv = (lane < n) ? P.ub[s * n + lane] : 0.0;
    bnd[lane] = (lane < nC) ? P.Clb[s * nC + lane] : 0.0;
    bnd[32 + lane] = (lane < nC) ? P.Cub[s * nC + lane] : 0.0;
    __syncwarp();
    double h[NV];
    const double* Hrow = Hs + (lane < n ? lane : 0) * LD;
#pragma unroll
    for (int l = 0; l < NV; ++l) h[l] = (lane < n) ? Hrow[l] : 0.0;
    const double hdiag = (lane < n) ? Hrow[lane] : 0.0;
    __syncwarp();
    QpRegShared S;
    S.R = smem_addr(Hs); S.col = smem_addr(col); S.vd = smem_addr(vd); S.C = smem_addr(Cs);
    S.clb = smem_addr(bnd); S.cub = S.clb + 8 * 32; S.dd = S.clb + 8 * 64;
    S.red_rows = S.feet_mask = S.red_blk = S.red_other = S.b = 0; S.skip_act = P.active_set == nullptr;
    double x;
    const QpResult res = warp_qp_solve_reg<NV, SPLIT>(S, h, hdiag, nC, gk, lbv, ubv, P.max_iter, x);
    if (lane < n) P.x[s * n + lane] = x;
    if (lane == 0) {
      P.status[s] = res.status;
      P.iters[s] = res.iters;
      if (P.active_set) { P.active_set[2 * s] = res.act_box; P.active_set[2 * s + 1] = res.act_rows; }
    }
    __syncwarp();
  }
}

template <int NV, bool SPLIT>
static int launch_qp_reg_k(const QpParams& P, int sms, cudaStream_t st) {
  const int LD = NV | 1, wpc = 8;
  const int per_warp = NV * (NV + 2) + (NV & 1) + ((P.nC * LD + 1) & ~1) + 224;
  const size_t smem = (size_t)wpc * per_warp * sizeof(double);
  auto kern = wbc_qp_reg_kernel<NV, SPLIT>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long need = (P.N + wpc - 1) / wpc;
  const long long cap = (long long)sms * 2;
  kern<<<(int)(need < cap ? need : cap), wpc * 32, smem, st>>>(P);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
}
template <int NV>
static int launch_qp_reg(const QpParams& P, int sms, cudaStream_t st) {
  return P.nC <= 16 ? launch_qp_reg_k<NV, true>(P, sms, st) : launch_qp_reg_k<NV, false>(P, sms, st);
}

// DFMA-saturating microkernel: 8 independent FMA chains per thread
__global__ void __launch_bounds__(256) wbc_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int build_dev_model(const WbcTreeTable* t, DevModel* m) {
  memset(m, 0, sizeof(*m));
  if (t->njoints < 2 || t->njoints > WBC_MAX_JOINTS) return fail(WBC_ERR_INVALID_ARG, "njoints out of range%s");
  if (t->nv < 6 || t->nv > WBC_MAX_NV || t->nq != t->nv + 1) return fail(WBC_ERR_INVALID_ARG, "need 6 <= nv <= 32 and nq == nv + 1%s");
  if (t->nframes < WBC_HOT_FRAMES || t->nframes > WBC_MAX_FRAMES) return fail(WBC_ERR_INVALID_ARG, "need 6..16 frame slots (5 EE + trunk first)%s");
  if (t->jtype[1] != WBC_JT_FREEFLYER || t->parent[1] != 0) return fail(WBC_ERR_INVALID_ARG, "joint 1 must be the free-flyer root%s");
  m->njoints = t->njoints; m->nq = t->nq; m->nv = t->nv; m->nframes = t->nframes;
  int maxd = 0;
  for (int j = 0; j < t->njoints; ++j) {
    m->parent[j] = t->parent[j];
    m->jtype[j] = t->jtype[j];
    m->idx_q[j] = t->idx_q[j];
    if (j > 0) {
      if (t->parent[j] < 0 || t->parent[j] >= j) return fail(WBC_ERR_INVALID_ARG, "parents must precede children%s");
      if (j > 1 && t->jtype[j] != WBC_JT_REVOLUTE && t->jtype[j] != WBC_JT_PRISMATIC)
        return fail(WBC_ERR_UNSUPPORTED, "only revolute / prismatic joints below the free-flyer root%s");
      m->depth[j] = m->depth[t->parent[j]] + 1;
      if (m->depth[j] > maxd) maxd = m->depth[j];
    }
    memcpy(m->plR[j], t->placement_R[j], sizeof(double) * 9);
    {
      static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      m->pl_ident[j] = memcmp(t->placement_R[j], I3, sizeof(I3)) == 0;
    }
    memcpy(m->plp[j], t->placement_p[j], sizeof(double) * 3);
    memcpy(m->axis[j], t->axis[j], sizeof(double) * 3);
    m->mass[j] = t->mass[j];
    memcpy(m->com[j], t->com[j], sizeof(double) * 3);
    m->total_mass += t->mass[j];
  }
  m->maxdepth = maxd;
  m->nrounds = 0;
  while ((1 << m->nrounds) < maxd) ++m->nrounds;       // 2^nrounds >= depth of the deepest joint
  for (int j = 0; j < WBC_MAX_JOINTS; ++j) m->anc[0][j] = (j > 0 && j < t->njoints) ? t->parent[j] : 0;
  for (int r = 1; r < WBC_FK_ROUNDS; ++r)
    for (int j = 0; j < WBC_MAX_JOINTS; ++j) m->anc[r][j] = m->anc[r - 1][m->anc[r - 1][j]];
  for (int k = 0; k < WBC_MAX_NV; ++k) m->col_q[k] = -1;
  for (int j = 1; j < t->njoints; ++j) {
    const int iv = t->idx_v[j];
    if (t->jtype[j] == WBC_JT_FREEFLYER) {
      for (int e = 0; e < 3; ++e) {
        m->col_joint[iv + e] = j; m->col_ang[iv + e] = 0; m->col_axis[iv + e][e] = 1.0;
        m->col_joint[iv + 3 + e] = j; m->col_ang[iv + 3 + e] = 1; m->col_axis[iv + 3 + e][e] = 1.0;
      }
    } else {
      m->col_joint[iv] = j;
      m->col_ang[iv] = (t->jtype[j] == WBC_JT_REVOLUTE);
      memcpy(m->col_axis[iv], t->axis[j], sizeof(double) * 3);
      m->col_q[iv] = t->idx_q[j];
    }
  }
  for (int j = 1; j < t->njoints; ++j) {
    uint32_t mask = 0;
    for (int a = j; a > 0; a = t->parent[a]) {
      const int nvj = (t->jtype[a] == WBC_JT_FREEFLYER) ? 6 : 1;
      for (int k = t->idx_v[a]; k < t->idx_v[a] + nvj; ++k) {
        mask |= 1u << k;
        m->sub_joints[k] |= 1u << j;           // column k moves joint j
      }
    }
    m->joint_supp[j] = mask;
  }
  for (int f = 0; f < t->nframes; ++f) {
    if (t->frame_parent[f] < 0 || t->frame_parent[f] >= t->njoints) return fail(WBC_ERR_INVALID_ARG, "frame parent out of range%s");
    m->frame_parent[f] = t->frame_parent[f];
    m->frame_supp[f] = m->joint_supp[t->frame_parent[f]];
    memcpy(m->frR[f], t->frame_R[f], sizeof(double) * 9);
    {
      static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      m->fr_ident[f] = memcmp(t->frame_R[f], I3, sizeof(I3)) == 0;
    }
    memcpy(m->frp[f], t->frame_p[f], sizeof(double) * 3);
  }
  memcpy(m->lower, t->lower, sizeof(double) * WBC_MAX_NQ);
  memcpy(m->upper, t->upper, sizeof(double) * WBC_MAX_NQ);
  memcpy(m->velocity, t->velocity, sizeof(double) * WBC_MAX_NV);
  return WBC_OK;
}

// Every entry point launches with the model's device pointers: the caller's current device must be the one the model
// was created on (wbc_b200.h).
static int check_device(const WbcModel* model) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != model->device)
    return fail(WBC_ERR_INVALID_ARG, "the model was created on another CUDA device than the current one%s");
  return WBC_OK;
}

static size_t model_smem_bytes() { return (sizeof(DevModel) + 15) & ~size_t(15); }

#ifndef WBC_ONLY_HOT
#define WBC_ONLY_HOT 0         // 1: A/B builds -- only the bench instantiation (nv = 26, reduced front) is compiled (30 s instead of 3 min)
#endif
template <int NV, bool DBG, bool SPLIT, bool FD, int NF = 0, bool RED = false, bool MULTI = false>
static int launch_step_k(const WbcModel* model, const StepParams& P, cudaStream_t st, int* info) {
  if constexpr (WBC_ONLY_HOT && !(NV == 26 && !DBG && SPLIT && !FD && NF == 3 && RED)) {
    return fail(WBC_ERR_UNSUPPORTED, "this is a WBC_ONLY_HOT build: only the nv = 26 reduced-front instantiation exists%s");
  } else {
  constexpr StepLayout L = step_layout(NV, WBC_MAX_NC, RED);
  const size_t per_warp = (size_t)L.total * sizeof(double);
  int max_optin = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, model->device));
  constexpr int ctas = StepWarps<SPLIT, RED>::ctas;
  const size_t extra = RED ? WBC_GROUP_REPORT_BYTES : 0;      // CTA-shared group-report slots (StepParams::group_report)
  int warps = (int)(((size_t)(max_optin + 1024) / ctas - 1024 - model_smem_bytes() - extra) / per_warp);
  if (warps > StepWarps<SPLIT, RED>::value) warps = StepWarps<SPLIT, RED>::value;
  if (warps < 1) return fail(WBC_ERR_UNSUPPORTED, "shared memory too small for one state%s");
  const size_t smem = model_smem_bytes() + warps * per_warp + extra;
  auto kern = wbc_step_kernel<NV, DBG, SPLIT, FD, NF, RED, MULTI>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long need = (P.N + warps - 1) / warps;
  long long cap = (long long)model->sm_count * ctas;
  if (P.grid_cap > 0 && P.grid_cap < cap) cap = P.grid_cap;
  int grid = (int)(need < cap ? need : cap);
  if (grid < 1) grid = 1;
  if (info) {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
    info[0] = (int)cap; info[1] = warps * 32; info[2] = (int)smem; info[3] = fa.numRegs;
    return WBC_OK;
  }
  if (P.N == 0) return WBC_OK;
  kern<<<grid, warps * 32, smem, st>>>(P);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
  }
}

template <int NV, bool DBG>
static int launch_step_t(const WbcModel* model, const StepParams& P, cudaStream_t st, int* info) {
  // FD: the finite-difference joint-task modes ("MANI" / "HYBRID") are their own instantiation
  const bool fd = (P.cfg.task_mask & WBC_TASK_JOINT) && (P.cfg.joint_mode == WBC_JOINT_MANI || P.cfg.joint_mode == WBC_JOINT_HYBRID);
  if (DBG)                                                                     // the accessor never reaches the solver
    return fd ? launch_step_k<NV, DBG, true, true>(model, P, st, info) : launch_step_k<NV, DBG, true, false>(model, P, st, info);
  if (P.nC <= 16) {
    // velDamperJointConstraints locks v >= gripper_joint_id - 2 + 6 (lb = ub = 0, :627-631): when these are exactly the
    // last three DoFs (both arms) the solver eliminates them at compile time
    if (NV - (P.cfg.gripper_joint_id - 2 + 6) == 3) {
      // the twelve foot equality rows eliminated up front (wbc_qp_red.inc) when model and configuration allow it
      // (also with the finite-difference joint task, "MANI" / "HYBRID": what sim3.py:145-148 runs)
      if constexpr (!DBG) {
        // the whole closed-loop horizon in one launch (wbc_rollout)
        if (P.red_ok && P.K > 1) return fd ? launch_step_k<NV, DBG, true, true, 3, true, true>(model, P, st, info)
                                           : launch_step_k<NV, DBG, true, false, 3, true, true>(model, P, st, info);
        if (P.red_ok) return fd ? launch_step_k<NV, DBG, true, true, 3, true>(model, P, st, info)
                                : launch_step_k<NV, DBG, true, false, 3, true>(model, P, st, info);
      }
      return fd ? launch_step_k<NV, DBG, true, true, 3>(model, P, st, info) : launch_step_k<NV, DBG, true, false, 3>(model, P, st, info);
    }
    return fd ? launch_step_k<NV, DBG, true, true>(model, P, st, info) : launch_step_k<NV, DBG, true, false>(model, P, st, info);
  }
  return fd ? launch_step_k<NV, DBG, false, true>(model, P, st, info) : launch_step_k<NV, DBG, false, false>(model, P, st, info);
}

// TMA staging of the input block: float64 arrays with 16-byte aligned bases.  WBC_B200_BULK=<mask> overrides the policy
// (bit 0 targets, 1 task memory, 2 references, 3 IMU quaternion; 0 = everything on cp.async): A/B runs.
static void set_bulk(StepParams* P) {
  static const int forced = [] { const char* e = getenv("WBC_B200_BULK"); return e && e[0] ? atoi(e) : -1; }();
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  int want = forced >= 0 ? forced : P->bulk_in;
  if (P->f32_in) want = 0;
  if (!al(P->io.targets)) want &= ~WBC_BULK_TARGETS;
  if (!al(P->io.mem_in)) want &= ~WBC_BULK_MEM;
  if (!al(P->io.ref)) want &= ~WBC_BULK_REF;
  if (!P->io.imu_quat || !al(P->io.imu_quat)) want &= ~WBC_BULK_IMU;
  P->bulk_in = want;
}

template <bool DBG>
static int launch_step(const WbcModel* model, const StepParams& P_in, cudaStream_t st, int* info) {
  StepParams P = P_in;
  if (P.N >= (1LL << 31) - (1LL << 20)) return fail(WBC_ERR_UNSUPPORTED, "more than 2^31 - 2^20 states per call (the kernel indexes states with 32 bits)%s");
  set_bulk(&P);
  switch (model->host.nv) {
    case 25: return launch_step_t<25, DBG>(model, P, st, info);
    case 26: return launch_step_t<26, DBG>(model, P, st, info);
    default: return fail(WBC_ERR_UNSUPPORTED, "the fused step kernel is instantiated for nv = 25 (A1 + PX100) and nv = 26 (A1 + WX200)%s");
  }
}

// Can the QP start from the reduced (null-space) front of wbc_qp_red.inc?  Needs: the four foot constraints on and the
// gripper's off (its rows would be three more equalities), at most 16 rows, every foot frame supported by the six
// free-flyer columns plus three consecutive limb columns [6 + 3 j, 9 + 3 j) (each j once), and no other row of C
// touching a leg column (so that C Z is a column selection).  WBC_B200_NO_REDUCED=1 switches it off (A/B runs).
static void set_reduced(const DevModel& M, StepParams* P) {
  P->red_ok = 0;
  P->red_rows = 0;
  P->red_feet_mask = 0;
  P->red_blk = 0;
  P->red_other = 0;
  const char* off = getenv("WBC_B200_NO_REDUCED");
  if (off && off[0] == '1') return;
  if (P->nC > 16 || P->row_com >= 0 || P->row_ee[4] >= 0) return;
  unsigned seen = 0;
  for (int t = 0; t < 4; ++t) {
    if (P->row_ee[t] < 0) return;
    const unsigned supp = (unsigned)M.frame_supp[t];
    if ((supp & 0x3Fu) != 0x3Fu) return;
    const unsigned limb = supp >> 6;
    int j = -1;
    for (int b = 0; b < 4; ++b)
      if (limb == (7u << (3 * b))) j = b;
    if (j < 0 || ((seen >> j) & 1u)) return;
    seen |= 1u << j;
    P->red_rows |= (unsigned)P->row_ee[t] << (8 * j);
    P->red_blk |= (unsigned)j << (2 * t);
    P->red_feet_mask |= 7u << P->row_ee[t];
  }
  const unsigned legs = 0xFFFu << 6;
  if (P->row_trunk >= 0 && ((unsigned)M.frame_supp[WBC_FRAME_TRUNK] & legs)) return;
  for (int e = 0; e < P->cfg.n_extra_rows; ++e)
    if ((unsigned)M.frame_supp[P->cfg.extra_frame[e]] & legs) return;
  {
    int m = 0;                               // the rows of C that are not foot rows, in order: one per spare lane NV + m
    for (int r = 0; r < P->nC; ++r)
      if (!((P->red_feet_mask >> r) & 1u)) {
        if (m >= 4) return;                  // (nC <= 16 with 12 foot rows: cannot happen)
        P->red_other |= (unsigned)r << (8 * m++);
      }
  }
  P->red_ok = 1;
}

static int check_cfg(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, StepParams* P) {
  if (!model || !cfg || !io) return fail(WBC_ERR_INVALID_ARG, "null model / config / io%s");
  if (cfg->joint_mode < WBC_JOINT_ZERO || cfg->joint_mode > WBC_JOINT_HYBRID) return fail(WBC_ERR_INVALID_ARG, "unknown joint_mode%s");
  if (cfg->n_extra_rows < 0 || cfg->n_extra_rows > WBC_MAX_EXTRA_ROWS) return fail(WBC_ERR_INVALID_ARG, "n_extra_rows out of range%s");
  for (int e = 0; e < cfg->n_extra_rows; ++e)
    if (cfg->extra_frame[e] < 0 || cfg->extra_frame[e] >= WBC_HOT_FRAMES) return fail(WBC_ERR_INVALID_ARG, "extra row frame must be a hot frame slot 0..5%s");
  for (int e = 0; e < cfg->n_extra_rows; ++e)
    if (cfg->extra_rf[e] < WBC_RF_WORLD || cfg->extra_rf[e] > WBC_RF_LOCAL_WORLD_ALIGNED)
      return fail(WBC_ERR_INVALID_ARG, "extra row reference frame must be WORLD, LOCAL or LOCAL_WORLD_ALIGNED%s");
  if (!(cfg->task_mask & 0x7f)) return fail(WBC_ERR_INVALID_ARG, "no task selected%s");
  // end_effector_index_list_joint[4]: a joint id, or njoints when the model has no such joint (pin.getJointId's answer
  // for an unknown name: nothing gets locked, laikago_vx300)
  if (cfg->gripper_joint_id < 2 || cfg->gripper_joint_id > model->host.njoints)
    return fail(WBC_ERR_INVALID_ARG, "gripper_joint_id must be in [2, njoints]%s");
  if ((cfg->task_mask & WBC_TASK_JOINT) && cfg->joint_mode == WBC_JOINT_HYBRID &&
      (cfg->arm_base_id < 1 || cfg->arm_base_id >= model->host.njoints))
    return fail(WBC_ERR_INVALID_ARG, "arm_base_id must be a joint id in [1, njoints) for the HYBRID joint task%s");
  if (!io->q || !io->targets || !io->mem_in || !io->ref) return fail(WBC_ERR_INVALID_ARG, "q / targets / mem_in / ref are required%s");
  if (!(io->dt > 0.0)) return fail(WBC_ERR_INVALID_ARG, "dt must be positive%s");
  if (int rc = check_device(model)) return rc;
  P->model = model->dev;
  P->cfg = *cfg;
  if (P->cfg.max_iter <= 0) P->cfg.max_iter = 200;           // same default as wbc_qp_solve
  P->io = *io;
  memset(&P->dbg, 0, sizeof(P->dbg));
  P->nC = cfg_nc(*cfg);
  set_rows(P);
  P->m_rows = cfg_m(*cfg, model->host.nv);
  P->flags = (int)io->flags & ~WBC_STEP_FLAG_WEIGHTS_IDENTITY;
  {
    bool ident = true;
    for (int t = 0; t < 6 && ident; ++t) {
      const double* W = (t < 5) ? cfg->ee_weight[t] : cfg->trunk_weight;
      for (int k = 0; k < 36; ++k)
        if (W[k] != ((k % 7 == 0) ? 1.0 : 0.0)) { ident = false; break; }
    }
    if (ident) P->flags |= WBC_STEP_FLAG_WEIGHTS_IDENTITY;   // W = I: W (J w) == J w exactly, skip the 6x6 products
  }
  if (P->nC > WBC_MAX_NC) return fail(WBC_ERR_UNSUPPORTED, "more than 32 constraint rows%s");
  P->grid_cap = 0;
  P->f32_in = P->f32_out = 0;
  // TMA staging policy, by measurement (profiles/r2_ab_experiments.txt): device-resident inputs of a device-resident tick
  // are 1.8 % faster on plain cp.async, so it is off here; wbc_step_host switches it on for the closed-loop zero-copy tick,
  // where the per-tick inputs come out of pinned host memory (+4.7 % end to end)
  P->bulk_in = 0;
  P->K = 1;
  P->group_report = 0;
  if (io->joint_targets && !io->q_next) return fail(WBC_ERR_INVALID_ARG, "joint_targets needs q_next%s");
  set_reduced(model->host, P);
  return WBC_OK;
}

extern "C" {

int wbc_abi_version(void) { return WBC_ABI_VERSION; }
const char* wbc_last_error(void) { return g_err; }

int wbc_model_create(const WbcTreeTable* table, WbcModel** out_model) {
  if (!table || !out_model) return fail(WBC_ERR_INVALID_ARG, "null argument%s");
  WbcModel* m = new (std::nothrow) WbcModel;
  if (!m) return fail(WBC_ERR_INVALID_ARG, "out of host memory%s");
  m->pipe_ready = false;
  m->tune.N = -1; m->tune.key = 0; m->tune.calls = 0; m->tune.choice = -1; m->tune.ev_ready = false;
  int rc = build_dev_model(table, &m->host);
  if (rc != WBC_OK) { delete m; return rc; }
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device);
  if (e == cudaSuccess) e = cudaMalloc(&m->dev, sizeof(DevModel));
  if (e == cudaSuccess) e = cudaMemcpy(m->dev, &m->host, sizeof(DevModel), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete m;
    return fail(WBC_ERR_CUDA, "wbc_model_create: %s (no CUDA device? there is no CPU fallback)", cudaGetErrorString(e));
  }
  *out_model = m;
  return WBC_OK;
}

void wbc_model_destroy(WbcModel* model) {
  if (!model) return;
  if (model->tune.ev_ready)
    for (int k = 0; k < 4; ++k) cudaEventDestroy(model->tune.ev[k]);
  if (model->pipe_ready) {
    for (int s = 0; s < WBC_PIPE_STREAMS; ++s) {
      cudaStreamSynchronize(model->pipe[s]);
      cudaStreamDestroy(model->pipe[s]);
      cudaEventDestroy(model->pipe_done[s]);
    }
    cudaEventDestroy(model->pipe_start);
  }
  cudaFree(model->dev);
  delete model;
}

int wbc_config_rows(const WbcConfig* cfg, int32_t nv, int32_t* m_rows, int32_t* nc_rows) {
  if (!cfg) return fail(WBC_ERR_INVALID_ARG, "null config%s");
  if (m_rows) *m_rows = cfg_m(*cfg, nv);
  if (nc_rows) *nc_rows = cfg_nc(*cfg);
  return WBC_OK;
}

static int fk_common(const WbcModel* model, FkJacParams& P, void* stream) {
  if (int rc = check_device(model)) return rc;
  const int wpc = 8;
  const size_t per_warp = (WBC_MAX_JOINTS * WBC_T_STRIDE + WBC_MAX_FRAMES * WBC_T_STRIDE + 40) * sizeof(double);
  const size_t smem = model_smem_bytes() + wpc * per_warp;
  CUDA_TRY(cudaFuncSetAttribute(wbc_fk_jac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (P.N == 0) return WBC_OK;
  long long need = (P.N + wpc - 1) / wpc;
  const long long cap = (long long)model->sm_count * WBC_FKJ_CTAS;
  const int grid = (int)(need < cap ? need : cap);
  wbc_fk_jac_kernel<<<grid, wpc * 32, smem, (cudaStream_t)stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
}

int wbc_fk_jac(const WbcModel* model, const double* q, int64_t N, const int32_t* frame_slots, int32_t nsel,
               int32_t ref_frame, double* out_oMf, double* out_J, void* stream) {
  if (!model || !q || N < 0) return fail(WBC_ERR_INVALID_ARG, "null model / q or negative N%s");
  if (nsel < 0 || nsel > WBC_MAX_FRAMES || (nsel > 0 && !frame_slots)) return fail(WBC_ERR_INVALID_ARG, "bad frame selection%s");
  if (ref_frame < 0 || ref_frame > 2) return fail(WBC_ERR_INVALID_ARG, "ref_frame must be WORLD, LOCAL or LOCAL_WORLD_ALIGNED%s");
  FkJacParams P;
  memset(&P, 0, sizeof(P));
  P.model = model->dev; P.q = q; P.N = N; P.nsel = nsel; P.rf = ref_frame;
  for (int i = 0; i < nsel; ++i) {
    if (frame_slots[i] < 0 || frame_slots[i] >= model->host.nframes) return fail(WBC_ERR_INVALID_ARG, "frame slot out of range%s");
    P.slots[i] = frame_slots[i];
  }
  P.out_oMf = out_oMf; P.out_J = out_J;
  return fk_common(model, P, stream);
}

int wbc_joint_jacobians(const WbcModel* model, const double* q, int64_t N, double* out_oMi, double* out_J, void* stream) {
  if (!model || !q || N < 0) return fail(WBC_ERR_INVALID_ARG, "null model / q or negative N%s");
  FkJacParams P;
  memset(&P, 0, sizeof(P));
  P.model = model->dev; P.q = q; P.N = N;
  P.out_oMi = out_oMi; P.out_Jw = out_J;
  return fk_common(model, P, stream);
}

int wbc_init_memory(const WbcModel* model, const double* q, int64_t N, double* mem_out, double* ref_out, void* stream) {
  if (!model || !q || !mem_out || !ref_out || N < 0) return fail(WBC_ERR_INVALID_ARG, "null argument or negative N%s");
  if (int rc = check_device(model)) return rc;
  const int wpc = 8;
  const size_t per_warp = (WBC_MAX_JOINTS * WBC_T_STRIDE + WBC_MAX_FRAMES * WBC_T_STRIDE + 40) * sizeof(double);
  const size_t smem = model_smem_bytes() + wpc * per_warp;
  CUDA_TRY(cudaFuncSetAttribute(wbc_init_memory_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (N == 0) return WBC_OK;
  long long need = (N + wpc - 1) / wpc;
  const long long cap = (long long)model->sm_count * 4;
  wbc_init_memory_kernel<<<(int)(need < cap ? need : cap), wpc * 32, smem, (cudaStream_t)stream>>>(model->dev, q, N, mem_out, ref_out);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
}

int wbc_integrate(const WbcModel* model, const double* q, const double* v, int64_t N, double scale, double* q_out,
                  void* stream) {
  if (!model || !q || !v || !q_out || N < 0) return fail(WBC_ERR_INVALID_ARG, "null argument or negative N%s");
  if (int rc = check_device(model)) return rc;
  if (N == 0) return WBC_OK;
  long long need = (N + 7) / 8;
  const long long cap = (long long)model->sm_count * 8;
  wbc_integrate_kernel<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(model->dev, q, v, N, scale, q_out);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
}

int wbc_base_estimate(const WbcModel* model, const double* q, const double* imu_quat, const double* targets, int64_t N,
                      double* q_out, double* base_out, void* stream) {
  if (!model || !q || !targets || N < 0) return fail(WBC_ERR_INVALID_ARG, "null model / q / targets or negative N%s");
  if (!q_out && !base_out) return fail(WBC_ERR_INVALID_ARG, "no output requested%s");
  if (int rc = check_device(model)) return rc;
  const int wpc = 8;
  const size_t per_warp = (WBC_MAX_JOINTS * WBC_T_STRIDE + WBC_MAX_FRAMES * WBC_T_STRIDE + 40) * sizeof(double);
  const size_t smem = model_smem_bytes() + wpc * per_warp;
  CUDA_TRY(cudaFuncSetAttribute(wbc_base_estimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (N == 0) return WBC_OK;
  long long need = (N + wpc - 1) / wpc;
  const long long cap = (long long)model->sm_count * 4;
  wbc_base_estimate_kernel<<<(int)(need < cap ? need : cap), wpc * 32, smem, (cudaStream_t)stream>>>(model->dev, q, imu_quat, targets, N,
                                                                                                q_out, base_out);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
}

int wbc_assemble(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, int64_t N,
                 const WbcAssembleOut* out, void* stream) {
  StepParams P;
  int rc = check_cfg(model, cfg, io, &P);
  if (rc != WBC_OK) return rc;
  if (!out || N < 0) return fail(WBC_ERR_INVALID_ARG, "null outputs or negative N%s");
  P.dbg = *out;
  P.N = N;
  return launch_step<true>(model, P, (cudaStream_t)stream, nullptr);
}

int wbc_step(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, int64_t N, void* stream) {
  StepParams P;
  int rc = check_cfg(model, cfg, io, &P);
  if (rc != WBC_OK) return rc;
  if (N < 0 || !io->qdot || !io->status || !io->iters) return fail(WBC_ERR_INVALID_ARG, "qdot / status / iters are required%s");
  P.N = N;
  return launch_step<false>(model, P, (cudaStream_t)stream, nullptr);
}

// The fused tick for a caller that holds host arrays (the reference's calling convention: NumPy in, NumPy out,
// Robot_Wrapper4.py:1330-1412).  Zero-copy (page-locked arrays: the kernel reads / writes them over PCIe itself) or
// staged (the batch cut into `chunks` slices whose host -> device copies, kernel and device -> host copies overlap on
// three internal streams).  `stream` is ordered before the first access and after the last one.
int wbc_step_host(WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, const WbcHostIO* host, int64_t N,
                  int32_t chunks, void* stream) {
  StepParams P;
  int rc = check_cfg(model, cfg, io, &P);
  if (rc != WBC_OK) return rc;
  if (!host) return fail(WBC_ERR_INVALID_ARG, "null host io%s");
  if (N < 0 || !io->qdot || !io->status || !io->iters) return fail(WBC_ERR_INVALID_ARG, "qdot / status / iters are required%s");
  if (io->active_set) return fail(WBC_ERR_UNSUPPORTED, "wbc_step_host does not report active sets (io->active_set must be NULL)%s");
  if (host->dtype != WBC_HOST_F64 && host->dtype != WBC_HOST_F32) return fail(WBC_ERR_INVALID_ARG, "unknown host dtype%s");
  if (!host->targets) return fail(WBC_ERR_INVALID_ARG, "host targets are required (they are runWBC's arguments)%s");
  // closed loop: the configuration and the task memory are advanced in place on the device, as runWBC mutates its object
  if (io->q_next && (io->q_next != io->q || host->q))
    return fail(WBC_ERR_INVALID_ARG, "closed loop: q_next must alias the resident q (host q NULL)%s");
  if (io->mem_out && (io->mem_out != io->mem_in || host->mem_in))
    return fail(WBC_ERR_INVALID_ARG, "closed loop: mem_out must alias the resident mem_in (host mem_in NULL)%s");
  if ((host->imu_quat || host->joint_targets) && !io->q_next)
    return fail(WBC_ERR_INVALID_ARG, "imu_quat / joint_targets belong to the closed-loop tick (io->q_next)%s");
  if (N == 0) return WBC_OK;
  const bool f32 = host->dtype == WBC_HOST_F32;
  const int nq = model->host.nq, nv = model->host.nv;
  // the seven travelling arrays: host pointer, device twin, elements per state, FP32 flag bit, direction
  struct Arr { const void* h; const void** d; int n; int bit; bool out; bool flt; };
  Arr arr[9] = {
      {host->q, (const void**)&P.io.q, nq, WBC_F32_Q, false, true},
      {host->targets, (const void**)&P.io.targets, WBC_TARGETS_STRIDE, WBC_F32_TARGETS, false, true},
      {host->mem_in, (const void**)&P.io.mem_in, WBC_MEM_STRIDE, WBC_F32_MEM, false, true},
      {host->ref, (const void**)&P.io.ref, WBC_REF_STRIDE, WBC_F32_REF, false, true},
      {host->imu_quat, (const void**)&P.io.imu_quat, 4, WBC_F32_IMU, false, true},
      {host->qdot, (const void**)&P.io.qdot, nv, WBC_F32_QDOT, true, true},
      {host->joint_targets, (const void**)&P.io.joint_targets, nq - 7, WBC_F32_JOINTS, true, true},
      {host->status, (const void**)&P.io.status, 1, 0, true, false},
      {host->iters, (const void**)&P.io.iters, 1, 0, true, false}};
  P.f32_in = P.f32_out = 0;
  if (f32)
    for (int k = 0; k < 7; ++k)
      if (arr[k].h) (arr[k].out ? P.f32_out : P.f32_in) |= arr[k].bit;
  if (host->flags & WBC_HOST_FLAG_DELTA_INPUTS) {
    if (!f32 || !io->q_next || !io->mem_out || host->mem_in || host->q)
      return fail(WBC_ERR_INVALID_ARG, "increment inputs need float32 host arrays and the closed-loop tick (resident q / task memory)%s");
    P.f32_in |= WBC_F32_DELTA;
  }
  cudaEvent_t tune_end = nullptr;          // set: record it on the caller's stream when this call has been queued
  if (chunks <= 0) {
    // Zero-copy: when every host array is page-locked (and therefore mapped into the device's address space under
    // unified addressing) the kernel reads the inputs straight from host memory -- its cp.async prefetch runs a whole tick
    // ahead, which hides the PCIe latency -- and writes the outputs straight into host memory: one launch, no staging
    // copies, no per-slice ramp-up / ramp-down.
    bool mapped = true;
    const void* dptr[9];
    for (int k = 0; k < 9 && mapped; ++k) {
      dptr[k] = nullptr;
      if (!arr[k].h) continue;
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, arr[k].h) != cudaSuccess) { cudaGetLastError(); mapped = false; break; }
      if (a.type != cudaMemoryTypeHost || !a.devicePointer) mapped = false;
      else dptr[k] = a.devicePointer;
    }
    bool zero_copy = mapped;
    // chunks == 0: self-tuning between the zero-copy launch and WBC_HOST_AUTO_STAGED staged slices (chunks < 0: zero-copy
    // whenever the arrays are page-locked).  The staged candidate needs its staging twins.
    const bool can_stage = !(host->imu_quat && !io->imu_quat) && !(host->joint_targets && !io->joint_targets);
    if (mapped && chunks == 0 && can_stage && N >= 4096) {
      WbcModel::HostTune& T = model->tune;
      const int key = (f32 ? 1 : 0) | (io->q_next ? 2 : 0) | (host->flags << 2) | (host->q ? 64 : 0) | (host->mem_in ? 128 : 0);
      if (T.N != N || T.key != key) { T.N = N; T.key = key; T.calls = 0; T.choice = -1; }
      if (!T.ev_ready) {
        for (int k = 0; k < 4; ++k) CUDA_TRY(cudaEventCreate(&T.ev[k]));
        T.ev_ready = true;
      }
      if (T.choice < 0 && T.calls >= 4 && cudaEventQuery(T.ev[1]) == cudaSuccess && cudaEventQuery(T.ev[3]) == cudaSuccess) {
        float tz = 0.f, ts = 0.f;
        if (cudaEventElapsedTime(&tz, T.ev[0], T.ev[1]) == cudaSuccess && cudaEventElapsedTime(&ts, T.ev[2], T.ev[3]) == cudaSuccess)
          T.choice = (ts < 0.97f * tz) ? WBC_HOST_AUTO_STAGED : 0;     // one launch wins a tie
        else cudaGetLastError();
      }
      const int call = T.calls++;
      if (T.choice >= 0) zero_copy = T.choice == 0;
      else if (call == 2 || call == 3) zero_copy = false;               // the staged candidate's turn
      if (T.choice < 0 && (call == 1 || call == 3)) {                   // the second call of each candidate is the timed one
        CUDA_TRY(cudaEventRecord(T.ev[call == 1 ? 0 : 2], (cudaStream_t)stream));
        tune_end = T.ev[call == 1 ? 1 : 3];
      }
    }
    if (zero_copy) {
      for (int k = 0; k < 9; ++k)
        if (arr[k].h) *arr[k].d = dptr[k];
      // closed-loop tick from host buffers: targets / IMU quaternion (host) and task memory / references (device) are staged
      // by the TMA engine (cp.async.bulk + mbarrier); the open-loop call, whose q rows travel by cp.async next to them,
      // measured 9 % slower with it and keeps cp.async (set_bulk() drops what is not float64 / 16-byte aligned)
      if (io->q_next) P.bulk_in = WBC_BULK_TARGETS | WBC_BULK_MEM | WBC_BULK_REF | WBC_BULK_IMU;
      P.group_report = (host->status || host->iters) ? 1 : 0;   // 4-byte stores over PCIe: one per barrier group, not one per warp
      P.N = N;
      rc = launch_step<false>(model, P, (cudaStream_t)stream, nullptr);
      if (rc == WBC_OK && tune_end) CUDA_TRY(cudaEventRecord(tune_end, (cudaStream_t)stream));
      return rc;
    }
    chunks = WBC_HOST_AUTO_STAGED;         // pageable host memory, or the staged candidate of the self-tuning mode
  }
  // staged: the device twins of the travelling arrays are the staging space
  if (host->imu_quat && !io->imu_quat) return fail(WBC_ERR_INVALID_ARG, "staged copies: io->imu_quat is needed as staging space%s");
  if (host->joint_targets && !io->joint_targets) return fail(WBC_ERR_INVALID_ARG, "staged copies: io->joint_targets is needed as staging space%s");
  if (!model->pipe_ready) {
    for (int s = 0; s < WBC_PIPE_STREAMS; ++s) {
      CUDA_TRY(cudaStreamCreateWithFlags(&model->pipe[s], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&model->pipe_done[s], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&model->pipe_start, cudaEventDisableTiming));
    model->pipe_ready = true;
  }
  if (chunks < 1) chunks = 1;
  if (chunks > N) chunks = (int32_t)N;
  if (chunks > 64) chunks = 64;
  // Slice boundaries: whole waves of the persistent kernel (one state per resident warp), so that no slice ends on a
  // partly filled wave; the first and the last slice are a single wave -- they are the exposed parts of the pipeline
  // (nothing overlaps the first copy in and the last copy out).
  // (measured and dropped: capping each slice's kernel at half of the SMs so that two slices run side by side --
  //  64 M instead of 82 M steps/s end to end; P.grid_cap stays available for such experiments)
  int64_t bnd[65];
  {
    int info[4] = {0, 0, 0, 0};
    P.N = N;
    rc = launch_step<false>(model, P, nullptr, info);
    if (rc != WBC_OK) return rc;
    const int64_t wave = (int64_t)info[0] * (info[1] / 32);
    const int64_t W = (N + wave - 1) / wave;
    if (chunks >= 3 && W >= chunks) {
      bnd[0] = 0;
      bnd[1] = wave;
      for (int c = 2; c < chunks; ++c) bnd[c] = wave * (1 + (W - 2) * (c - 1) / (chunks - 2));
      bnd[chunks] = N;
    } else {
      for (int c = 0; c <= chunks; ++c) bnd[c] = N * c / chunks;
    }
  }
  const cudaStream_t user = (cudaStream_t)stream;
  const size_t fsz = f32 ? sizeof(float) : sizeof(double);
  const void* dev0[9];                       // device twins at state 0
  for (int k = 0; k < 9; ++k) dev0[k] = *arr[k].d;
  // any failure inside the slice loop still joins the internal streams back into the caller's stream before returning:
  // copies / kernels already queued must not outlive the call unobserved (the caller may free or reuse its buffers)
  cudaError_t ce = cudaEventRecord(model->pipe_start, user);
  int used = 0;
  for (int c = 0; c < chunks && ce == cudaSuccess && rc == WBC_OK; ++c) {
    const int64_t lo = bnd[c], n = bnd[c + 1] - lo;
    if (n <= 0) continue;
    const cudaStream_t st = model->pipe[c % WBC_PIPE_STREAMS];
    if (c < WBC_PIPE_STREAMS) { ce = cudaStreamWaitEvent(st, model->pipe_start, 0); used = c + 1; }
    for (int k = 0; k < 9 && ce == cudaSuccess; ++k) {
      const size_t esz = arr[k].flt ? fsz : sizeof(int32_t);
      // the kernel's view of this slice: a travelling array is indexed with the host element size, a resident one is float64
      const size_t step = (size_t)arr[k].n * (arr[k].h ? esz : (arr[k].flt ? sizeof(double) : sizeof(int32_t)));
      *arr[k].d = dev0[k] ? (const char*)dev0[k] + lo * step : nullptr;
      if (arr[k].h && !arr[k].out)           // NULL: resident on the device already
        ce = cudaMemcpyAsync(const_cast<void*>(*arr[k].d), (const char*)arr[k].h + lo * step, n * step, cudaMemcpyHostToDevice, st);
    }
    if (ce != cudaSuccess) break;
    if (io->q_next) P.io.q_next = const_cast<double*>(P.io.q);           // closed loop: in place, slice by slice
    if (io->mem_out) P.io.mem_out = const_cast<double*>(P.io.mem_in);
    P.N = n;
    rc = launch_step<false>(model, P, st, nullptr);
    for (int k = 0; k < 9 && ce == cudaSuccess && rc == WBC_OK; ++k) {
      if (!arr[k].h || !arr[k].out) continue;
      const size_t step = (size_t)arr[k].n * (arr[k].flt ? fsz : sizeof(int32_t));
      ce = cudaMemcpyAsync((char*)const_cast<void*>(arr[k].h) + lo * step, *arr[k].d, n * step, cudaMemcpyDeviceToHost, st);
    }
  }
  for (int s = 0; s < used; ++s) {
    cudaError_t e2 = cudaEventRecord(model->pipe_done[s], model->pipe[s]);
    if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(user, model->pipe_done[s], 0);
    if (e2 != cudaSuccess) { cudaStreamSynchronize(model->pipe[s]); if (ce == cudaSuccess) ce = e2; }
  }
  if (rc != WBC_OK) return rc;
  if (ce != cudaSuccess) return fail(WBC_ERR_CUDA, "wbc_step_host: %s", cudaGetErrorString(ce));
  if (tune_end) CUDA_TRY(cudaEventRecord(tune_end, user));
  return WBC_OK;
}

int wbc_step_host_path(const WbcModel* model) {
  if (!model) return -2;
  return model->tune.choice;
}

int wbc_rollout(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, const double* targets_traj,
                const double* imu_traj, int32_t K, int64_t N, void* stream) {
  if (!io || !targets_traj || K < 0) return fail(WBC_ERR_INVALID_ARG, "null io / targets_traj or negative K%s");
  if ((io->q_next && io->q_next != io->q) || (io->mem_out && io->mem_out != io->mem_in))
    return fail(WBC_ERR_INVALID_ARG, "wbc_rollout advances q and mem in place%s");
  WbcStepIO tick = *io;
  tick.q_next = const_cast<double*>(io->q);
  tick.mem_out = const_cast<double*>(io->mem_in);
  StepParams P;
  if (N < 0 || !tick.qdot || !tick.status || !tick.iters) return fail(WBC_ERR_INVALID_ARG, "qdot / status / iters are required%s");
  if (K == 0 || N == 0) return WBC_OK;
  tick.targets = targets_traj;
  tick.imu_quat = imu_traj;
  int rc = check_cfg(model, cfg, &tick, &P);
  if (rc != WBC_OK) return rc;
  P.N = N;
  // The whole horizon in ONE launch where the reduced-front instantiation applies (four foot constraints, <= 16 rows, any
  // joint-task mode): a robot stays with one warp for all K ticks, so nothing separates the ticks but that warp's own
  // program order -- no relaunch, no drain tail per tick.  WBC_B200_ROLLOUT_LAUNCHES=1 forces one launch per tick (A/B runs).
  static const bool per_tick = [] { const char* e = getenv("WBC_B200_ROLLOUT_LAUNCHES"); return e && e[0] == '1'; }();
  const bool locked3 = model->host.nv - (P.cfg.gripper_joint_id - 2 + 6) == 3;
  if (K > 1 && !per_tick && P.red_ok && P.nC <= 16 && locked3 && (model->host.nv == 25 || model->host.nv == 26)) {
    P.K = K;
    return launch_step<false>(model, P, (cudaStream_t)stream, nullptr);
  }
  for (int k = 0; k < K; ++k) {
    P.io.targets = targets_traj + (size_t)k * N * WBC_TARGETS_STRIDE;
    P.io.imu_quat = imu_traj ? imu_traj + (size_t)k * N * 4 : nullptr;
    rc = launch_step<false>(model, P, (cudaStream_t)stream, nullptr);
    if (rc != WBC_OK) return rc;
  }
  return WBC_OK;
}

int wbc_step_launch_info(const WbcModel* model, int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* regs) {
  if (!model) return fail(WBC_ERR_INVALID_ARG, "null model%s");
  StepParams P;
  memset(&P, 0, sizeof(P));
  P.N = 1 << 20;
  P.cfg.gripper_joint_id = model->host.nv - 7;   // the shipped arms: gripper + two fingers locked (the NF = 3 instantiation)
  // the usual constraint set (sim3.py:145-148, BASELINE P2 / P3): trunk box + four feet -> the reduced-front instantiation
  P.cfg.constraint_mask = WBC_CON_TRUNK | (WBC_CON_FR << 0) | (WBC_CON_FR << 1) | (WBC_CON_FR << 2) | (WBC_CON_FR << 3);
  P.nC = cfg_nc(P.cfg);
  set_rows(&P);
  set_reduced(model->host, &P);
  int info[4] = {0, 0, 0, 0};
  int rc = launch_step<false>(model, P, nullptr, info);
  if (rc != WBC_OK) return rc;
  if (grid) *grid = info[0];
  if (block) *block = info[1];
  if (smem_bytes) *smem_bytes = info[2];
  if (regs) *regs = info[3];
  return WBC_OK;
}

int wbc_qp_solve(int64_t N, int32_t nv, int32_t m, int32_t nC, const double* A, const double* b, const double* H,
                 const double* g, const double* lb, const double* ub, const double* C, const double* Clb,
                 const double* Cub, int32_t max_iter, double* x, int32_t* status, int32_t* iters,
                 uint64_t* active_set, void* stream) {
  if (N < 0 || nv < 1 || nv > WBC_MAX_NV) return fail(WBC_ERR_INVALID_ARG, "need N >= 0 and 1 <= nv <= 32%s");
  if (nC < 0 || nC > WBC_MAX_NC) return fail(WBC_ERR_UNSUPPORTED, "need 0 <= nC <= 32%s");
  const bool haveA = A && b, haveH = H && g;
  if (haveA == haveH) return fail(WBC_ERR_INVALID_ARG, "give exactly one of (A, b) or (H, g)%s");
  if (haveA && m < 1) return fail(WBC_ERR_INVALID_ARG, "A needs m >= 1 rows%s");
  if (!lb || !ub || !x || !status || !iters) return fail(WBC_ERR_INVALID_ARG, "lb / ub / x / status / iters are required%s");
  if (nC > 0 && (!C || !Clb || !Cub)) return fail(WBC_ERR_INVALID_ARG, "C / Clb / Cub are required when nC > 0%s");
  if (N == 0) return WBC_OK;
  QpParams P;
  P.N = N; P.nv = nv; P.m = m; P.nC = nC; P.max_iter = max_iter > 0 ? max_iter : 200;
  P.A = haveA ? A : nullptr; P.b = haveA ? b : nullptr; P.H = haveH ? H : nullptr; P.g = haveH ? g : nullptr;
  P.lb = lb; P.ub = ub; P.C = C; P.Clb = Clb; P.Cub = Cub;
  P.x = x; P.status = status; P.iters = iters; P.active_set = (unsigned long long*)active_set;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // the robot sizes get the register-resident solver; any other 1 <= nv <= 32 the run-time-size one
  if (nv == 26) return launch_qp_reg<26>(P, sms, (cudaStream_t)stream);
  if (nv == 25) return launch_qp_reg<25>(P, sms, (cudaStream_t)stream);
  const int LD = nv | 1;
  const int per_warp = 2 * nv * LD + nC * LD + 96 + ((2 * nv * LD + nC * LD) & 1);
  const int wpc = 8;
  const size_t smem = (size_t)wpc * per_warp * sizeof(double);
  CUDA_TRY(cudaFuncSetAttribute(wbc_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long need = (N + wpc - 1) / wpc;
  const long long cap = (long long)sms * 2;
  wbc_qp_kernel<<<(int)(need < cap ? need : cap), wpc * 32, smem, (cudaStream_t)stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  return WBC_OK;
}

int wbc_measure_fp64_peak(double* flops_per_s, void* stream) {
  if (!flops_per_s) return fail(WBC_ERR_INVALID_ARG, "null output%s");
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 1 << 16;
  double* buf = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  float best = 1e30f;
  cudaError_t e = cudaMalloc(&buf, sizeof(double) * blocks * threads);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) {
    wbc_dfma_kernel<<<blocks, threads, 0, st>>>(buf, 1024);           // warm-up
    e = cudaGetLastError();
  }
  for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) {
    e = cudaEventRecord(e0, st);
    if (e != cudaSuccess) break;
    wbc_dfma_kernel<<<blocks, threads, 0, st>>>(buf, iters);
    e = cudaEventRecord(e1, st);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess && ms < best) best = ms;
  }
  if (e0) cudaEventDestroy(e0);                                       // one exit: nothing leaks on a failure
  if (e1) cudaEventDestroy(e1);
  if (buf) cudaFree(buf);
  if (e != cudaSuccess) return fail(WBC_ERR_CUDA, "wbc_measure_fp64_peak: %s", cudaGetErrorString(e));
  *flops_per_s = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3);
  return WBC_OK;
}

}  // extern "C"
