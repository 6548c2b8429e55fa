// wbc_device.cuh -- device-side model, small-matrix helpers, rotation conversions and the
// warp-level kinematics passes (FK, joint-Jacobian columns, frame placements).
//
// One robot state is processed by one warp.  Lane j owns joint j during the tree pass and
// velocity column k = lane during everything that is column-parallel (nv <= 32).
//
// Reference semantics followed (paths relative to the reference root):
//   pin.forwardKinematics / computeJointJacobians / updateFramePlacements / getFrameJacobian as
//   called from wrappers/Robot_Wrapper4.py:400-405, 458-488, 641-758  (restated in oracle/pin.py);
//   scipy Rotation.from_matrix / from_euler('xyz') / as_matrix / as_euler('xyz') as called from
//   wrappers/Robot_Wrapper4.py:363-367, 714-715, 964-970, 1101-1102 (restated in oracle/rotation_port.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wbc_b200.h"

#define WBC_FULL_MASK 0xffffffffu
#define WBC_T_STRIDE 13   // doubles per stored SE3 (9 R row-major + 3 p + 1 pad): odd stride => no 8-way bank conflicts

// Device copy of the kinematic tree (built on the host from WbcTreeTable, lives in global memory,
// staged into shared memory once per CTA).
#define WBC_FK_ROUNDS 5       // 2^5 >= WBC_MAX_JOINTS
struct DevModel {
  int32_t njoints, nq, nv, nframes, maxdepth, nrounds, pad1, pad2;
  int32_t parent[WBC_MAX_JOINTS];
  int32_t anc[WBC_FK_ROUNDS][WBC_MAX_JOINTS];   // anc[r][j]: the 2^r-th ancestor of joint j (0 = none): FK by pointer jumping
  int32_t jtype[WBC_MAX_JOINTS];
  int32_t idx_q[WBC_MAX_JOINTS];
  int32_t depth[WBC_MAX_JOINTS];
  int32_t col_joint[WBC_MAX_NV];     // joint that owns velocity column k
  int32_t col_ang[WBC_MAX_NV];       // 1: column is a rotation about col_axis, 0: a translation along it
  int32_t col_q[WBC_MAX_NV];         // q index a 1-DoF column integrates into (free-flyer columns: -1)
  uint32_t joint_supp[WBC_MAX_JOINTS];   // bit k: column k supports joint j
  uint32_t sub_joints[WBC_MAX_NV];       // bit j: joint j is in the subtree moved by column k
  int32_t frame_parent[WBC_MAX_FRAMES];
  uint32_t frame_supp[WBC_MAX_FRAMES];
  int32_t pl_ident[WBC_MAX_JOINTS];      // 1: the joint placement has an identity rotation (pure translation)
  int32_t fr_ident[WBC_MAX_FRAMES];      // 1: the frame offset has an identity rotation
  double plR[WBC_MAX_JOINTS][9];
  double plp[WBC_MAX_JOINTS][3];
  double axis[WBC_MAX_JOINTS][3];
  double col_axis[WBC_MAX_NV][3];    // joint-frame axis of column k (e_x/e_y/e_z for free-flyer columns)
  double frR[WBC_MAX_FRAMES][9];
  double frp[WBC_MAX_FRAMES][3];
  double lower[WBC_MAX_NQ];
  double upper[WBC_MAX_NQ];
  double velocity[WBC_MAX_NV];
  double mass[WBC_MAX_JOINTS];
  double com[WBC_MAX_JOINTS][3];
  double total_mass;
};

// ------------------------------------------------------------------------------------------------
// explicit shared-memory access through 32-bit shared-window addresses
// ------------------------------------------------------------------------------------------------
// ptxas rematerialises the address of a dynamic-shared-memory element (S2UR SR_CgaCtaId, ULEA, IMAD, LEA ...,
// up to 15 instructions) in front of every access inside switch-case blocks instead of keeping one register
// live (seen in the SASS of the first register-resident solver).  The hot loops therefore address shared
// memory as  [base + immediate]  with `base` a 32-bit register that went through a self-shuffle, which
// ptxas cannot recompute.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  return (uint32_t)__shfl_sync(WBC_FULL_MASK, (int)a, threadIdx.x & 31);
}
// (a constant added to `a` after unrolling is folded into the instruction's immediate offset by ptxas)
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// predicated store: `if (c) sts_f64(a, v)` compiles to a branch around the (volatile) store with a convergence
// barrier (BSSY / BRA / STS / BSYNC); this is a compare and one predicated STS
__device__ __forceinline__ void sts_f64_if(bool c, uint32_t a, double v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(a), "d"(v), "r"((int)c) : "memory");
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void lds_vec3(uint32_t a, double* v) {
  v[0] = lds_f64(a); v[1] = lds_f64(a + 8); v[2] = lds_f64(a + 16);
}
__device__ __forceinline__ void lds_mat3(uint32_t a, double* R) {
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = lds_f64(a + 8 * i);
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// six predicated 128-bit stores at a, a + 16, ..., a + 80 behind one compare (lanes with c == false take no part in the
// store at all: fewer shared-memory wavefronts than dumping their values somewhere)
__device__ __forceinline__ void sts_f64x12_if(bool c, uint32_t a, const double* v) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %13, 0;\n\t"
      "@p st.shared.v2.f64 [%0], {%1, %2};\n\t@p st.shared.v2.f64 [%0+16], {%3, %4};\n\t"
      "@p st.shared.v2.f64 [%0+32], {%5, %6};\n\t@p st.shared.v2.f64 [%0+48], {%7, %8};\n\t"
      "@p st.shared.v2.f64 [%0+64], {%9, %10};\n\t@p st.shared.v2.f64 [%0+80], {%11, %12};\n\t}"
      ::"r"(a), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]), "d"(v[8]), "d"(v[9]),
        "d"(v[10]), "d"(v[11]), "r"((int)c)
      : "memory");
}

// Phase barrier: the warps of a group re-align (named barrier per group of WBC_SYNC_GROUP warps; 0 = whole CTA).
#ifndef WBC_SYNC_GROUP
#define WBC_SYNC_GROUP -1      // -1: two groups of half the CTA's warps each: 3.4 % faster than one group (less waiting for the slowest QP); three groups thrash the I-cache (-21 %)
#endif
template <bool ON>
__device__ __forceinline__ void phase_sync() {
  if (ON) {
#if WBC_SYNC_GROUP < 0 && defined(WBC_SYNC_SMSP)
    // A/B: groups by scheduler -- warps 0,1,4,5,... (sub-partitions 0 and 1) against 2,3,6,7,... (sub-partitions 2 and 3), so that
    // each sub-partition's instruction cache sees one instruction stream
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = (wid >> 1) & 1;
    const int n1 = ((nw >> 2) << 1) + ((nw & 3) > 2 ? (nw & 3) - 2 : 0), n0 = nw - n1;
    asm volatile("bar.sync %0, %1;" ::"r"(g ? 2 : 1), "r"((g ? n1 : n0) * 32) : "memory");
#elif WBC_SYNC_GROUP < 0
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int half = (nw + 1) >> 1;
    const bool g = wid >= half;
    asm volatile("bar.sync %0, %1;" ::"r"(g ? 2 : 1), "r"((g ? nw - half : half) * 32) : "memory");
#elif WBC_SYNC_GROUP > 0
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = wid / WBC_SYNC_GROUP;
    const int first = g * WBC_SYNC_GROUP;
    const int cnt = (nw - first < WBC_SYNC_GROUP ? nw - first : WBC_SYNC_GROUP) * 32;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(cnt) : "memory");
#else
    __syncthreads();
#endif
  } else {
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 helpers (row-major)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_mul_bt(const double* A, const double* B, double* C) {  // C = A * B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
__device__ __forceinline__ void mat3_vec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
__device__ __forceinline__ void mat3t_vec(const double* A, const double* v, double* o) {  // o = A^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// Eigen::Quaternion::toRotationMatrix (no normalisation) -- what pin's free-flyer FK uses.
__device__ __forceinline__ void quat_to_matrix_eigen(double x, double y, double z, double w, double* R) {
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen Quaternion = Matrix3 (pin.integrate). q = (x, y, z, w).
__device__ __forceinline__ void matrix_to_quat_eigen(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

// Rotation about a unit axis; exact elementary forms for the aligned axes (JointModelRX/RY/RZ).
__device__ __forceinline__ void axis_angle_matrix(const double* a, double s, double c, double* R) {
  const double x = a[0], y = a[1], z = a[2];
  if (x == 1.0 && y == 0.0 && z == 0.0) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = c; R[5] = -s; R[6] = 0; R[7] = s; R[8] = c;
  } else if (x == 0.0 && y == 1.0 && z == 0.0) {
    R[0] = c; R[1] = 0; R[2] = s; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = -s; R[7] = 0; R[8] = c;
  } else if (x == 0.0 && y == 0.0 && z == 1.0) {
    R[0] = c; R[1] = -s; R[2] = 0; R[3] = s; R[4] = c; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  } else {
    const double t = 1 - c;
    R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
    R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
    R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
  }
}

// --- SciPy Rotation restatements (oracle/rotation_port.py has the same sequences) -----------------
// Rotation.from_matrix's pre-step (scipy _rotation_xp.py:51-88): a matrix whose Gramian R R^T is not close to I
// (isclose with atol 1e-12, rtol 1e-5: 1e-12 off the diagonal, 1e-5 on it) is replaced by its nearest orthogonal
// matrix U V^T.  The free-flyer FK does not normalise the quaternion (Eigen toRotationMatrix), so this triggers
// whenever a configuration's quaternion is not unit to ~1e-13 -- e.g. qpJointb "MANI" perturbing the quaternion
// entries.  U V^T is the polar factor; it is computed by Newton's iteration X <- (X + X^-T) / 2, which converges
// quadratically to the same matrix the SVD gives.
__device__ __forceinline__ void scipy_orthogonalize(double* R) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) {
      const double g = R[3 * i] * R[3 * j] + R[3 * i + 1] * R[3 * j + 1] + R[3 * i + 2] * R[3 * j + 2];
      ok = ok && ((i == j) ? (fabs(g - 1.0) <= 1e-12 + 1e-5) : (fabs(g) <= 1e-12));
    }
  if (ok) return;
#pragma unroll 1
  for (int it = 0; it < 30; ++it) {
    double C[9];                                   // cofactor matrix: X^-T = C / det
    C[0] = R[4] * R[8] - R[5] * R[7]; C[1] = R[5] * R[6] - R[3] * R[8]; C[2] = R[3] * R[7] - R[4] * R[6];
    C[3] = R[2] * R[7] - R[1] * R[8]; C[4] = R[0] * R[8] - R[2] * R[6]; C[5] = R[1] * R[6] - R[0] * R[7];
    C[6] = R[1] * R[5] - R[2] * R[4]; C[7] = R[2] * R[3] - R[0] * R[5]; C[8] = R[0] * R[4] - R[1] * R[3];
    const double det = R[0] * C[0] + R[1] * C[1] + R[2] * C[2];
    double change = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double x = 0.5 * (R[i] + C[i] / det);
      change = fmax(change, fabs(x - R[i]));
      R[i] = x;
    }
    if (change <= 1e-16) break;
  }
}

// Rotation.from_matrix(R).as_quat(), (x, y, z, w), no sign canonicalisation; a copy of R is orthogonalised first
// when SciPy would do so.
__device__ __forceinline__ void scipy_quat_from_matrix(const double* Rin, double* q) {
  double R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = Rin[i];
  scipy_orthogonalize(R);
  const double tr = R[0] + R[4] + R[8];
  int choice = 0;
  double best = R[0];
  if (R[4] > best) { best = R[4]; choice = 1; }
  if (R[8] > best) { best = R[8]; choice = 2; }
  if (tr > best) { choice = 3; }
  if (choice == 0) {
    q[0] = 1 - tr + 2 * R[0]; q[1] = R[3] + R[1]; q[2] = R[6] + R[2]; q[3] = R[7] - R[5];
  } else if (choice == 1) {
    q[0] = R[3] + R[1]; q[1] = 1 - tr + 2 * R[4]; q[2] = R[7] + R[5]; q[3] = R[2] - R[6];
  } else if (choice == 2) {
    q[0] = R[6] + R[2]; q[1] = R[7] + R[5]; q[2] = 1 - tr + 2 * R[8]; q[3] = R[3] - R[1];
  } else {
    q[0] = R[7] - R[5]; q[1] = R[2] - R[6]; q[2] = R[3] - R[1]; q[3] = 1 + tr;
  }
  // (SciPy divides by the norm; multiplying by rsqrt(|q|^2) differs by <= 2 ulp and saves a square root and four
  //  divisions, ~130 instructions of every tick)
  const double rn = rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= rn; q[1] *= rn; q[2] *= rn; q[3] *= rn;
}

__device__ __forceinline__ void quat_compose(const double* p, const double* q, double* o) {  // o = p * q
  const double cx = p[1] * q[2] - p[2] * q[1];
  const double cy = p[2] * q[0] - p[0] * q[2];
  const double cz = p[0] * q[1] - p[1] * q[0];
  o[0] = p[3] * q[0] + q[3] * p[0] + cx;
  o[1] = p[3] * q[1] + q[3] * p[1] + cy;
  o[2] = p[3] * q[2] + q[3] * p[2] + cz;
  o[3] = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
}

// Rotation.from_euler('xyz', e).as_quat(): extrinsic => qz * (qy * qx)
__device__ __forceinline__ void scipy_quat_from_euler_xyz(const double* e, double* q) {
  double s0, c0, s1, c1, s2, c2;
  sincos(e[0] / 2.0, &s0, &c0);
  sincos(e[1] / 2.0, &s1, &c1);
  sincos(e[2] / 2.0, &s2, &c2);
  const double qx[4] = {s0, 0.0, 0.0, c0};
  const double qy[4] = {0.0, s1, 0.0, c1};
  const double qz[4] = {0.0, 0.0, s2, c2};
  double t[4];
  quat_compose(qy, qx, t);
  quat_compose(qz, t, q);
}

// the same from the half-angle sines / cosines (computed one angle per lane by the caller)
__device__ __forceinline__ void scipy_quat_from_half_sincos(double s0, double c0, double s1, double c1, double s2, double c2,
                                                            double* q) {
  const double qx[4] = {s0, 0.0, 0.0, c0};
  const double qy[4] = {0.0, s1, 0.0, c1};
  const double qz[4] = {0.0, 0.0, s2, c2};
  double t[4];
  quat_compose(qy, qx, t);
  quat_compose(qz, t, q);
}

// Rotation.as_matrix()
__device__ __forceinline__ void scipy_matrix_from_quat(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);      R[2] = 2 * (xz + yw);
  R[3] = 2 * (xy + zw);     R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
  R[6] = 2 * (xz - yw);     R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}

__device__ __forceinline__ double wrap_pi(double a) {  // (a + pi) % (2 pi) - pi, floor-mod
  // |a| <= 2 pi for every caller (sums / differences of atan2 results), so one fold replaces fmod; r - 2 pi is exact
  // (Sterbenz), i.e. the result is bit-identical to the fmod form
  const double pi = 3.141592653589793, two_pi = 6.283185307179586;
  double r = a + pi;
  if (r >= two_pi) r -= two_pi;
  if (r < 0) r += two_pi;
  return r - pi;
}

// Rotation.as_euler('xyz') from a quaternion (extrinsic, asymmetric sequence, sign = +1)
__device__ __forceinline__ void scipy_euler_xyz_from_quat(const double* q, double* e) {
  const double pi = 3.141592653589793;
  const double a = q[3] - q[1], b = q[0] + q[2], c = q[1] + q[3], d = q[2] - q[0];
  const double half_sum = atan2(b, a), half_diff = atan2(d, c);
  double a1 = 2 * atan2(hypot(c, d), hypot(a, b));
  const bool case1 = fabs(a1) <= 1e-7, case2 = fabs(a1 - pi) <= 1e-7;
  double a0, a2;
  if (!(case1 || case2)) {
    a0 = half_sum - half_diff;
    a2 = half_sum + half_diff;
  } else {
    a2 = 0.0;
    a0 = case1 ? 2 * half_sum : -2 * half_diff;
  }
  a1 -= pi / 2;
  e[0] = wrap_pi(a0); e[1] = wrap_pi(a1); e[2] = wrap_pi(a2);
}

// The same conversion for ONE quaternion held by every lane of the warp (uniform input), with the transcendental calls
// spread over lanes: one hypot, one atan2 and one fmod are executed warp-wide instead of 2 + 3 + 3 by a single lane.
// Same functions on the same arguments, so the result is bit-identical to scipy_euler_xyz_from_quat.  All 32 lanes call;
// every lane returns the full triple.
__device__ __forceinline__ void warp_scipy_euler_xyz_from_quat(const double* q, int lane, double* e) {
  const double pi = 3.141592653589793;
  const double a = q[3] - q[1], b = q[0] + q[2], c = q[1] + q[3], d = q[2] - q[0];
  const int r = lane % 3;
  const double hx = (lane & 1) ? a : c, hy = (lane & 1) ? b : d;    // lane 0: hypot(c, d), lane 1: hypot(a, b)
  const double h = sqrt(fma(hx, hx, hy * hy));                      // (|q| = 1: no overflow to guard against; <= 1 ulp from hypot)
  const double hcd = __shfl_sync(WBC_FULL_MASK, h, 0), hab = __shfl_sync(WBC_FULL_MASK, h, 1);
  const double ty = (r == 0) ? b : ((r == 1) ? d : hcd);
  const double tx = (r == 0) ? a : ((r == 1) ? c : hab);
  const double t = atan2(ty, tx);
  const double half_sum = __shfl_sync(WBC_FULL_MASK, t, 0), half_diff = __shfl_sync(WBC_FULL_MASK, t, 1);
  double a1 = 2 * __shfl_sync(WBC_FULL_MASK, t, 2);
  const bool case1 = fabs(a1) <= 1e-7, case2 = fabs(a1 - pi) <= 1e-7;
  double a0, a2;
  if (!(case1 || case2)) {
    a0 = half_sum - half_diff;
    a2 = half_sum + half_diff;
  } else {
    a2 = 0.0;
    a0 = case1 ? 2 * half_sum : -2 * half_diff;
  }
  a1 -= pi / 2;
  const double w = wrap_pi((r == 0) ? a0 : ((r == 1) ? a1 : a2));
  e[0] = __shfl_sync(WBC_FULL_MASK, w, 0); e[1] = __shfl_sync(WBC_FULL_MASK, w, 1); e[2] = __shfl_sync(WBC_FULL_MASK, w, 2);
}

// ------------------------------------------------------------------------------------------------
// Warp-level kinematics
// ------------------------------------------------------------------------------------------------
// forwardKinematics: oMi for every joint into shared memory (stride WBC_T_STRIDE), level by level.
// qs: this state's configuration in shared memory.  All 32 lanes must call.
__device__ __forceinline__ void warp_fk(const DevModel* __restrict__ M, const double* __restrict__ qs,
                                        double* __restrict__ oMi, int lane) {
  double Rl[9], pl[3];
  int myDepth = 0, par = 0;
  const bool active = lane >= 1 && lane < M->njoints;
  if (active) {
    const int jt = M->jtype[lane];
    const int iq = M->idx_q[lane];
    par = M->parent[lane];
    myDepth = M->depth[lane];
    double Rj[9];
    if (jt == WBC_JT_FREEFLYER) {
      quat_to_matrix_eigen(qs[iq + 3], qs[iq + 4], qs[iq + 5], qs[iq + 6], Rj);
      const double pj[3] = {qs[iq], qs[iq + 1], qs[iq + 2]};
      mat3_mul(M->plR[lane], Rj, Rl);
      mat3_vec(M->plR[lane], pj, pl);
      pl[0] += M->plp[lane][0]; pl[1] += M->plp[lane][1]; pl[2] += M->plp[lane][2];
    } else if (jt == WBC_JT_REVOLUTE) {
      double s, c;
      sincos(qs[iq], &s, &c);
      axis_angle_matrix(M->axis[lane], s, c, Rj);
      mat3_mul(M->plR[lane], Rj, Rl);
      pl[0] = M->plp[lane][0]; pl[1] = M->plp[lane][1]; pl[2] = M->plp[lane][2];
    } else {  // prismatic
      const double d = qs[iq];
      const double pj[3] = {M->axis[lane][0] * d, M->axis[lane][1] * d, M->axis[lane][2] * d};
#pragma unroll
      for (int i = 0; i < 9; ++i) Rl[i] = M->plR[lane][i];
      mat3_vec(M->plR[lane], pj, pl);
      pl[0] += M->plp[lane][0]; pl[1] += M->plp[lane][1]; pl[2] += M->plp[lane][2];
    }
  }
  const int maxdepth = M->maxdepth;
  for (int d = 1; d <= maxdepth; ++d) {
    if (active && myDepth == d) {
      double* out = oMi + lane * WBC_T_STRIDE;
      if (par == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) out[i] = Rl[i];
        out[9] = pl[0]; out[10] = pl[1]; out[11] = pl[2];
      } else {
        const double* P = oMi + par * WBC_T_STRIDE;
        double Rp[9], R[9], p[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rp[i] = P[i];
        mat3_mul(Rp, Rl, R);
        mat3_vec(Rp, pl, p);
#pragma unroll
        for (int i = 0; i < 9; ++i) out[i] = R[i];
        out[9] = p[0] + P[9]; out[10] = p[1] + P[10]; out[11] = p[2] + P[11];
      }
    }
    __syncwarp();
  }
}

// updateFramePlacements for the uploaded frame slots: oMf[f] = oMi[parent] * offset.
__device__ __forceinline__ void warp_frames(const DevModel* __restrict__ M, const double* __restrict__ oMi,
                                            double* __restrict__ oMf, int lane) {
  if (lane < M->nframes) {
    const int par = M->frame_parent[lane];
    double* out = oMf + lane * WBC_T_STRIDE;
    if (par == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) out[i] = M->frR[lane][i];
      out[9] = M->frp[lane][0]; out[10] = M->frp[lane][1]; out[11] = M->frp[lane][2];
    } else {
      const double* P = oMi + par * WBC_T_STRIDE;
      double Rp[9], R[9], p[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rp[i] = P[i];
      mat3_mul(Rp, M->frR[lane], R);
      mat3_vec(Rp, M->frp[lane], p);
#pragma unroll
      for (int i = 0; i < 9; ++i) out[i] = R[i];
      out[9] = p[0] + P[9]; out[10] = p[1] + P[10]; out[11] = p[2] + P[11];
    }
  }
  __syncwarp();
}

// computeJointJacobians: column `lane` of data.J in the WORLD frame, S = [lin(3); ang(3)].
__device__ __forceinline__ void warp_jac_column(const DevModel* __restrict__ M, const double* __restrict__ oMi,
                                                int lane, double* S) {
#pragma unroll
  for (int i = 0; i < 6; ++i) S[i] = 0.0;
  if (lane < M->nv) {
    const double* T = oMi + M->col_joint[lane] * WBC_T_STRIDE;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = T[i];
    double a[3];
    mat3_vec(R, M->col_axis[lane], a);
    if (M->col_ang[lane]) {
      const double p[3] = {T[9], T[10], T[11]};
      cross3(p, a, S);
      S[3] = a[0]; S[4] = a[1]; S[5] = a[2];
    } else {
      S[0] = a[0]; S[1] = a[1]; S[2] = a[2];
    }
  }
}

// getFrameJacobian column for a placement T (R|p) whose supporting-column mask is `supp`.
__device__ __forceinline__ void frame_jac_column(const double* S, uint32_t supp, int lane, const double* __restrict__ T,
                                                 int rf, double* Jc) {
  if (!((supp >> lane) & 1u)) {
#pragma unroll
    for (int i = 0; i < 6; ++i) Jc[i] = 0.0;
    return;
  }
  if (rf == WBC_RF_WORLD) {
#pragma unroll
    for (int i = 0; i < 6; ++i) Jc[i] = S[i];
    return;
  }
  const double p[3] = {T[9], T[10], T[11]};
  double pxw[3];
  cross3(p, S + 3, pxw);
  const double lin[3] = {S[0] - pxw[0], S[1] - pxw[1], S[2] - pxw[2]};
  if (rf == WBC_RF_LOCAL_WORLD_ALIGNED) {
    Jc[0] = lin[0]; Jc[1] = lin[1]; Jc[2] = lin[2];
    Jc[3] = S[3]; Jc[4] = S[4]; Jc[5] = S[5];
  } else {  // LOCAL
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = T[i];
    mat3t_vec(R, lin, Jc);
    mat3t_vec(R, S + 3, Jc + 3);
  }
}

// Warp-wide sum of a double.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(WBC_FULL_MASK, v, o);
  return v;
}
