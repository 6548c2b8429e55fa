/* wbc_oracle.c -- plain-C CPU restatement of the WBC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * One call = one open-loop tick of wrappers/Robot_Wrapper4.py:runWBC (:1330-1394) for N states:
 *   updateState (:387-428)  -> FK, WORLD joint Jacobian, frame placements        (oracle/pin.py)
 *   qpA / qpb (:1271-1294)  -> task rows A (:474-490) and targets b (:948-1157)  (oracle/robot_wrapper4.py)
 *   velDamperJointConstraints (:572-637), findConstraints (:764-836)
 *   QP(A, b, lb, ub, C, Clb, Cub).solveQP() (wrappers/QP_Wrapper.py:10-53): H = A^T A and g = -A^T b as dense
 *   products, then a Goldfarb-Idnani dual active-set solve with the pivoting rules of oracle/qp_wrapper.py.
 *
 * It exists (a) as a second, independent checker of the CUDA path that is fast enough to compare every state
 * of BASELINE config 2 (4096 states) and (b) as the CPU baseline timed by bench.py ("what native Pinocchio +
 * qpOASES would roughly cost": neither library is installable here).  It is pinned against the
 * NumPy/SciPy oracle by tests/test_oracle_c_cpu.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * The input structs are the ones of include/wbc_b200.h (plain C), so the same marshalled WbcTreeTable /
 * WbcConfig drive both sides.  Threads: OpenMP over states.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC -Iinclude oracle/wbc_oracle.c -o oracle/_build/libwbc_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "wbc_b200.h"

#define NVMAX WBC_MAX_NV
#define MMAX (36 + WBC_MAX_NV)
#define NCMAX WBC_MAX_NC
#define FEAS_TOL 1e-10
#define TIE_REL 1e-9           /* entering constraint: candidates within TIE_ABS + TIE_REL |min| of the most violated one */
#define TIE_ABS 1e-12          /* are tied, the lowest index wins (oracle/qp_wrapper.py) */
#define DEP_TOL 1e-13
#define PIVOT_REL 1e-14
#define MEM_OFF_PREV_EE_POS(i) (3 * (i))          /* layout of the task-memory block, include/wbc_b200.h */
#define MEM_OFF_PREV_EE_ROT(i) (15 + 9 * (i))

typedef struct { double R[9], p[3]; } SE3;

static void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void mat3_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static void se3_mul(const SE3* a, const SE3* b, SE3* o) {   /* o = a * b */
  double t[3];
  mat3_mul(a->R, b->R, o->R);
  mat3_vec(a->R, b->p, t);
  o->p[0] = t[0] + a->p[0]; o->p[1] = t[1] + a->p[1]; o->p[2] = t[2] + a->p[2];
}

/* Eigen quaternion -> matrix without normalisation (free-flyer FK, oracle/pin.py: quat_to_matrix) */
static void quat_to_matrix(double x, double y, double z, double w, double* R) {
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* joint transform of a revolute joint: exact elementary forms for the unit axes (oracle/pin.py: axis_angle_matrix) */
static void axis_angle(const double* a, double s, double c, double* R) {
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

/* ---- SciPy Rotation restatements (oracle/rotation_port.py) -------------------------------------- */
/* from_matrix's pre-step: nearest orthogonal matrix when R R^T is not close to I (Newton's polar iteration) */
static void sp_orthogonalize(double* R) {
  int ok = 1;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      const double g = R[3 * i] * R[3 * j] + R[3 * i + 1] * R[3 * j + 1] + R[3 * i + 2] * R[3 * j + 2];
      ok = ok && ((i == j) ? (fabs(g - 1.0) <= 1e-12 + 1e-5) : (fabs(g) <= 1e-12));
    }
  if (ok) return;
  for (int it = 0; it < 30; ++it) {
    double C[9];
    C[0] = R[4] * R[8] - R[5] * R[7]; C[1] = R[5] * R[6] - R[3] * R[8]; C[2] = R[3] * R[7] - R[4] * R[6];
    C[3] = R[2] * R[7] - R[1] * R[8]; C[4] = R[0] * R[8] - R[2] * R[6]; C[5] = R[1] * R[6] - R[0] * R[7];
    C[6] = R[1] * R[5] - R[2] * R[4]; C[7] = R[2] * R[3] - R[0] * R[5]; C[8] = R[0] * R[4] - R[1] * R[3];
    const double det = R[0] * C[0] + R[1] * C[1] + R[2] * C[2];
    double change = 0.0;
    for (int i = 0; i < 9; ++i) {
      const double x = 0.5 * (R[i] + C[i] / det);
      if (fabs(x - R[i]) > change) change = fabs(x - R[i]);
      R[i] = x;
    }
    if (change <= 1e-16) break;
  }
}
static void sp_quat_from_matrix(const double* Rin, double* q) {
  double R[9];
  memcpy(R, Rin, sizeof(R));
  sp_orthogonalize(R);
  const double tr = R[0] + R[4] + R[8];
  const double dec[4] = {R[0], R[4], R[8], tr};
  int choice = 0;
  for (int i = 1; i < 4; ++i)
    if (dec[i] > dec[choice]) choice = i;
  if (choice == 0) { q[0] = 1 - tr + 2 * R[0]; q[1] = R[3] + R[1]; q[2] = R[6] + R[2]; q[3] = R[7] - R[5]; }
  else if (choice == 1) { q[0] = R[3] + R[1]; q[1] = 1 - tr + 2 * R[4]; q[2] = R[7] + R[5]; q[3] = R[2] - R[6]; }
  else if (choice == 2) { q[0] = R[6] + R[2]; q[1] = R[7] + R[5]; q[2] = 1 - tr + 2 * R[8]; q[3] = R[3] - R[1]; }
  else { q[0] = R[7] - R[5]; q[1] = R[2] - R[6]; q[2] = R[3] - R[1]; q[3] = 1 + tr; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void sp_compose(const double* p, const double* q, double* o) {
  const double cx = p[1] * q[2] - p[2] * q[1];
  const double cy = p[2] * q[0] - p[0] * q[2];
  const double cz = p[0] * q[1] - p[1] * q[0];
  o[0] = p[3] * q[0] + q[3] * p[0] + cx;
  o[1] = p[3] * q[1] + q[3] * p[1] + cy;
  o[2] = p[3] * q[2] + q[3] * p[2] + cz;
  o[3] = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
}
static void sp_quat_from_euler_xyz(const double* e, double* q) {
  const double qx[4] = {sin(e[0] / 2.0), 0.0, 0.0, cos(e[0] / 2.0)};
  const double qy[4] = {0.0, sin(e[1] / 2.0), 0.0, cos(e[1] / 2.0)};
  const double qz[4] = {0.0, 0.0, sin(e[2] / 2.0), cos(e[2] / 2.0)};
  double t[4];
  sp_compose(qy, qx, t);
  sp_compose(qz, t, q);
}
static void sp_matrix_from_quat(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);      R[2] = 2 * (xz + yw);
  R[3] = 2 * (xy + zw);     R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
  R[6] = 2 * (xz - yw);     R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}
static double sp_wrap(double a) {
  const double pi = 3.141592653589793, two_pi = 2 * 3.141592653589793;
  double r = fmod(a + pi, two_pi);
  if (r < 0) r += two_pi;
  return r - pi;
}
static void sp_euler_xyz_from_quat(const double* q, double* e) {
  const double pi = 3.141592653589793;
  const double a = q[3] - q[1], b = q[0] + q[2], c = q[1] + q[3], d = q[2] - q[0];
  const double half_sum = atan2(b, a), half_diff = atan2(d, c);
  double a1 = 2 * atan2(hypot(c, d), hypot(a, b));
  const int case1 = fabs(a1) <= 1e-7, case2 = fabs(a1 - pi) <= 1e-7;
  double a0, a2;
  if (!(case1 || case2)) { a0 = half_sum - half_diff; a2 = half_sum + half_diff; }
  else { a2 = 0.0; a0 = case1 ? 2 * half_sum : -2 * half_diff; }
  a1 -= pi / 2;
  e[0] = sp_wrap(a0); e[1] = sp_wrap(a1); e[2] = sp_wrap(a2);
}

/* ---- kinematics (oracle/pin.py: forwardKinematics, computeJointJacobians, updateFramePlacements) ---- */
typedef struct {
  SE3 oMi[WBC_MAX_JOINTS];
  SE3 oMf[WBC_MAX_FRAMES];
  double J[6][NVMAX];                 /* data.J, WORLD */
  uint32_t supp[WBC_MAX_JOINTS];      /* supporting columns of each joint */
} Kin;

static void kinematics(const WbcTreeTable* t, const double* q, Kin* k) {
  memset(k->J, 0, sizeof(k->J));
  for (int i = 0; i < 9; ++i) k->oMi[0].R[i] = (i % 4 == 0);
  k->oMi[0].p[0] = k->oMi[0].p[1] = k->oMi[0].p[2] = 0.0;
  k->supp[0] = 0;
  for (int j = 1; j < t->njoints; ++j) {
    SE3 pl, jt, tmp;
    memcpy(pl.R, t->placement_R[j], sizeof(pl.R));
    memcpy(pl.p, t->placement_p[j], sizeof(pl.p));
    const int iq = t->idx_q[j], iv = t->idx_v[j];
    jt.p[0] = jt.p[1] = jt.p[2] = 0.0;
    if (t->jtype[j] == WBC_JT_FREEFLYER) {
      quat_to_matrix(q[iq + 3], q[iq + 4], q[iq + 5], q[iq + 6], jt.R);
      jt.p[0] = q[iq]; jt.p[1] = q[iq + 1]; jt.p[2] = q[iq + 2];
    } else if (t->jtype[j] == WBC_JT_REVOLUTE) {
      axis_angle(t->axis[j], sin(q[iq]), cos(q[iq]), jt.R);
    } else {
      for (int i = 0; i < 9; ++i) jt.R[i] = (i % 4 == 0);
      for (int i = 0; i < 3; ++i) jt.p[i] = t->axis[j][i] * q[iq];
    }
    se3_mul(&pl, &jt, &tmp);
    se3_mul(&k->oMi[t->parent[j]], &tmp, &k->oMi[j]);
    const SE3* M = &k->oMi[j];
    k->supp[j] = k->supp[t->parent[j]];
    if (t->jtype[j] == WBC_JT_FREEFLYER) {
      for (int e = 0; e < 3; ++e) {
        const double ax[3] = {M->R[e], M->R[3 + e], M->R[6 + e]};     /* R e_k */
        double px[3];
        cross3(M->p, ax, px);
        for (int r = 0; r < 3; ++r) {
          k->J[r][iv + e] = ax[r];
          k->J[r][iv + 3 + e] = px[r];
          k->J[3 + r][iv + 3 + e] = ax[r];
        }
        k->supp[j] |= (1u << (iv + e)) | (1u << (iv + 3 + e));
      }
    } else {
      double aw[3];
      mat3_vec(M->R, t->axis[j], aw);
      if (t->jtype[j] == WBC_JT_REVOLUTE) {
        double px[3];
        cross3(M->p, aw, px);
        for (int r = 0; r < 3; ++r) { k->J[r][iv] = px[r]; k->J[3 + r][iv] = aw[r]; }
      } else {
        for (int r = 0; r < 3; ++r) k->J[r][iv] = aw[r];
      }
      k->supp[j] |= 1u << iv;
    }
  }
  for (int f = 0; f < t->nframes; ++f) {
    SE3 off;
    memcpy(off.R, t->frame_R[f], sizeof(off.R));
    memcpy(off.p, t->frame_p[f], sizeof(off.p));
    se3_mul(&k->oMi[t->frame_parent[f]], &off, &k->oMf[f]);
  }
}

/* pin.getFrameJacobian(frame, rf) for rf in {WORLD, LOCAL_WORLD_ALIGNED, LOCAL}: 6 x nv, zero outside the support */
static void frame_jacobian(const WbcTreeTable* t, const Kin* k, int f, int rf, double Jf[6][NVMAX]) {
  const uint32_t supp = k->supp[t->frame_parent[f]];
  const SE3* M = &k->oMf[f];
  for (int c = 0; c < t->nv; ++c) {
    if (!((supp >> c) & 1u)) { for (int r = 0; r < 6; ++r) Jf[r][c] = 0.0; continue; }
    const double lin[3] = {k->J[0][c], k->J[1][c], k->J[2][c]}, ang[3] = {k->J[3][c], k->J[4][c], k->J[5][c]};
    if (rf == WBC_RF_WORLD) {
      for (int r = 0; r < 3; ++r) { Jf[r][c] = lin[r]; Jf[3 + r][c] = ang[r]; }
      continue;
    }
    double pxw[3];
    cross3(M->p, ang, pxw);
    const double l2[3] = {lin[0] - pxw[0], lin[1] - pxw[1], lin[2] - pxw[2]};
    if (rf == WBC_RF_LOCAL_WORLD_ALIGNED) {
      for (int r = 0; r < 3; ++r) { Jf[r][c] = l2[r]; Jf[3 + r][c] = ang[r]; }
    } else {
      for (int r = 0; r < 3; ++r) {
        Jf[r][c] = M->R[r] * l2[0] + M->R[3 + r] * l2[1] + M->R[6 + r] * l2[2];
        Jf[3 + r][c] = M->R[r] * ang[0] + M->R[3 + r] * ang[1] + M->R[6 + r] * ang[2];
      }
    }
  }
}

/* ---- dual active-set QP with factor updating (same pivoting rules as oracle/qp_wrapper.py) ---------- */
typedef struct { int status, iters; uint64_t act_box, act_rows; } QpOut;

static void qp_solve(int n, int nC, double H[NVMAX][NVMAX], const double* g, const double* lb, const double* ub,
                     double C[NCMAX][NVMAX], const double* clb, const double* cub, int max_iter, double* x, QpOut* out) {
  double L[NVMAX][NVMAX], J[NVMAX][NVMAX], Rm[NVMAX][NVMAX];   /* J = L^-T Q, Rm = upper triangular */
  int status = 0, iters = 0;
  double hd = 0.0;
  for (int i = 0; i < n; ++i) if (H[i][i] > hd) hd = H[i][i];
  const double piv_min = PIVOT_REL * (hd > 0 ? hd : 0.0);
  memset(L, 0, sizeof(L));
  for (int k = 0; k < n; ++k) {
    double d = H[k][k];
    for (int j = 0; j < k; ++j) d -= L[k][j] * L[k][j];
    if (!(d > piv_min)) { d = piv_min > 0 ? piv_min : 1.0; status |= WBC_QP_NOT_PD; }
    L[k][k] = sqrt(d);
    for (int i = k + 1; i < n; ++i) {
      double s = H[i][k];
      for (int j = 0; j < k; ++j) s -= L[i][j] * L[k][j];
      L[i][k] = s / L[k][k];
    }
  }
  /* J = L^-T: column c of L^-1 by forward substitution, stored as row c of J */
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) {
      double a = (i == c) ? 1.0 : 0.0;
      for (int j = c; j < i; ++j) a -= L[i][j] * J[c][j];
      J[c][i] = (i < c) ? 0.0 : a / L[i][i];
    }
  }
  /* x = -H^-1 g = -J J^T g */
  {
    double w[NVMAX];
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += J[i][j] * g[i]; w[j] = s; }
    for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s -= J[i][j] * w[j]; x[i] = s; }
  }
  int iq = 0, p_eq = 0;
  int ws[NVMAX];                 /* constraint id at each working-set position */
  int wside[NVMAX];              /* -1 lower, +1 upper, 0 equality */
  double u[NVMAX];
  int bstat[NVMAX], cstat[NCMAX];
  memset(bstat, 0, sizeof(bstat));
  memset(cstat, 0, sizeof(cstat));
  const int m = n + nC;
  int eq_next = 0;               /* next constraint index to test for lo == up */
  int done = 0;
  while (!done) {
    int ip = -1, side = -1, is_eq = 0;
    while (eq_next < m) {
      const int c = eq_next++;
      const double lo = c < n ? lb[c] : clb[c - n], up = c < n ? ub[c] : cub[c - n];
      if (lo == up) { ip = c; is_eq = 1; break; }
    }
    if (ip < 0) {
      double best = 0.0, viol[NVMAX + NCMAX];
      int vside[NVMAX + NCMAX];
      for (int c = 0; c < m; ++c) {
        viol[c] = 0.0; vside[c] = -1;
        if ((c < n ? bstat[c] : cstat[c - n]) != 0) continue;
        double ax, lo, up;
        if (c < n) { ax = x[c]; lo = lb[c]; up = ub[c]; }
        else { ax = 0; for (int j = 0; j < n; ++j) ax += C[c - n][j] * x[j]; lo = clb[c - n]; up = cub[c - n]; }
        const double slo = ax - lo, sup = up - ax;
        viol[c] = slo < sup ? slo : sup;
        vside[c] = (slo <= sup) ? -1 : +1;
        if (viol[c] < best) best = viol[c];
      }
      if (!(best < -FEAS_TOL)) break;
      const double thr = best + (TIE_ABS + TIE_REL * fabs(best));       /* lowest index inside the tie window */
      for (int c = 0; c < m; ++c)
        if ((c < n ? bstat[c] : cstat[c - n]) == 0 && viol[c] <= thr) { ip = c; side = vside[c]; break; }
      if (ip < 0) break;
    }
    const double sgn = (side > 0) ? -1.0 : 1.0;
    double nrm_v[NVMAX];
    for (int j = 0; j < n; ++j) nrm_v[j] = sgn * (ip < n ? (j == ip ? 1.0 : 0.0) : C[ip - n][j]);
    double u_new = 0.0;
    for (;;) {
      if (!is_eq && iters >= max_iter) { status |= WBC_QP_MAXITER; done = 1; break; }
      iters++;
      double d[NVMAX], z[NVMAX], r[NVMAX];
      for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += J[i][j] * nrm_v[i]; d[j] = s; }
      double dd = 0, dd2 = 0;
      for (int j = 0; j < n; ++j) { dd += d[j] * d[j]; if (j >= iq) dd2 += d[j] * d[j]; }
      for (int i = 0; i < n; ++i) { double s = 0; for (int j = iq; j < n; ++j) s += J[i][j] * d[j]; z[i] = s; }
      for (int k = iq - 1; k >= 0; --k) {
        double s = d[k];
        for (int j = k + 1; j < iq; ++j) s -= Rm[k][j] * r[j];
        r[k] = s / Rm[k][k];
      }
      double nx = 0;
      for (int j = 0; j < n; ++j) nx += nrm_v[j] * x[j];
      const double lo = ip < n ? lb[ip] : clb[ip - n], up = ip < n ? ub[ip] : cub[ip - n];
      const double s_ip = (side > 0) ? (up + nx) : (nx - lo);          /* n.x - bnd, n = sgn a */
      const int dependent = dd2 <= DEP_TOL * dd;
      if (is_eq) {
        if (dependent) { if (fabs(s_ip) > 1e-8) status |= WBC_QP_INFEASIBLE; break; }
        const double t = -s_ip / dd2;
        for (int i = 0; i < n; ++i) x[i] += t * z[i];
        u_new = t;
      } else {
        double t1 = INFINITY; int l = -1;
        for (int k = p_eq; k < iq; ++k)
          if (r[k] > 0.0) { const double ratio = u[k] / r[k]; if (ratio < t1) { t1 = ratio; l = k; } }
        const double t2 = dependent ? INFINITY : -s_ip / dd2;
        const double t = t1 < t2 ? t1 : t2;
        if (!(t < INFINITY)) { status |= WBC_QP_INFEASIBLE; done = 1; break; }
        for (int k = p_eq; k < iq; ++k) u[k] -= t * r[k];
        u_new += t;
        if (!dependent) for (int i = 0; i < n; ++i) x[i] += t * z[i];
        if (dependent || !(t2 <= t1)) {
          /* drop position l: Givens rotations restore the triangular R, applied to the columns of J */
          const int c_drop = ws[l];
          for (int k = l; k < iq - 1; ++k) {
            for (int i = 0; i <= k + 1; ++i) Rm[i][k] = Rm[i][k + 1];       /* shift column k+1 into k */
            const double a = Rm[k][k], b = Rm[k + 1][k];
            const double rho = sqrt(a * a + b * b);
            const double cg = rho > 0 ? a / rho : 1.0, sg = rho > 0 ? b / rho : 0.0;
            Rm[k][k] = rho; Rm[k + 1][k] = 0.0;
            for (int j = k + 1; j < iq - 1; ++j) {                          /* remaining (already shifted later) columns */
              const double r0 = Rm[k][j + 1], r1 = Rm[k + 1][j + 1];
              Rm[k][j + 1] = cg * r0 + sg * r1;
              Rm[k + 1][j + 1] = -sg * r0 + cg * r1;
            }
            for (int i = 0; i < n; ++i) {
              const double j0 = J[i][k], j1 = J[i][k + 1];
              J[i][k] = cg * j0 + sg * j1;
              J[i][k + 1] = -sg * j0 + cg * j1;
            }
            ws[k] = ws[k + 1]; wside[k] = wside[k + 1]; u[k] = u[k + 1];
          }
          if (c_drop < n) bstat[c_drop] = 0; else cstat[c_drop - n] = 0;
          iq--;
          continue;
        }
      }
      /* full step: constraint enters at position iq; Householder reflection maps d2 onto -sigma e_iq */
      {
        const double nrm = sqrt(dd2);
        const double sigma = d[iq] >= 0.0 ? nrm : -nrm;
        const double v_iq = d[iq] + sigma;
        const double beta = 1.0 / (sigma * v_iq);
        for (int i = 0; i < n; ++i) {
          const double bw = beta * (z[i] + sigma * J[i][iq]);
          J[i][iq] -= bw * v_iq;
          for (int j = iq + 1; j < n; ++j) J[i][j] -= bw * d[j];
        }
        for (int i = 0; i < iq; ++i) Rm[i][iq] = d[i];
        Rm[iq][iq] = -sigma;
        ws[iq] = ip; wside[iq] = is_eq ? 0 : side; u[iq] = u_new;
        const int st = is_eq ? 3 : (side > 0 ? 2 : 1);
        if (ip < n) bstat[ip] = st; else cstat[ip - n] = st;
        iq++;
        if (is_eq) p_eq = iq;
      }
      break;
    }
  }
  (void)wside;
  uint64_t ab = 0, ar = 0;
  for (int c = 0; c < n; ++c) ab |= ((uint64_t)(bstat[c] & 1) << (2 * c)) | ((uint64_t)((bstat[c] >> 1) & 1) << (2 * c + 1));
  for (int c = 0; c < nC; ++c) ar |= ((uint64_t)(cstat[c] & 1) << (2 * c)) | ((uint64_t)((cstat[c] >> 1) & 1) << (2 * c + 1));
  out->status = status; out->iters = iters; out->act_box = ab; out->act_rows = ar;
}

/* ---- one tick for one state --------------------------------------------------------------------- */
static int count_nc(const WbcConfig* c) {
  int n = 0;
  if (c->constraint_mask & WBC_CON_COM) n += 2;
  if (c->constraint_mask & WBC_CON_TRUNK) n += 4;
  for (int i = 0; i < 5; ++i) if (c->constraint_mask & (WBC_CON_FR << i)) n += 3;
  return n + c->n_extra_rows;
}

static int tick(const WbcTreeTable* t, const WbcConfig* cfg, const double* q, const double* targets, const double* mem_in,
                const double* ref, double dt, double* qdot, QpOut* qo, double* mem_out, double* A_out, double* b_out) {
  const int nv = t->nv;
  Kin k;
  kinematics(t, q, &k);
  double A[MMAX][NVMAX], b[MMAX], mem[WBC_MEM_STRIDE];
  memcpy(mem, mem_in, sizeof(mem));
  int m = 0;
  double Jf[6][NVMAX];
  /* qpCartesianA / qpCartesianB: FR FL RR RL GRIP, then Trunk (:839-876, :1160-1196) */
  for (int i = 0; i < 5; ++i) {
    if (!(cfg->task_mask & (1 << i))) continue;
    frame_jacobian(t, &k, i, WBC_RF_LOCAL_WORLD_ALIGNED, Jf);
    const double w = cfg->cart_task_weight[i];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < nv; ++c) {
        double s = 0;
        for (int e = 0; e < 6; ++e) s += cfg->ee_weight[i][6 * r + e] * (Jf[e][c] * w);   /* W (J w) (:480-482) */
        A[m + r][c] = s;
      }
    /* calcTargetVelEE3 (:1052-1157) */
    const double* target = targets + 3 * i;
    double* prev = mem + MEM_OFF_PREV_EE_POS(i);
    double* prevR = mem + MEM_OFF_PREV_EE_ROT(i);
    double ref_vel[3], err[3], ge[3], qr[4], Rref[9], dR[9], sk[9];
    for (int e = 0; e < 3; ++e) { ref_vel[e] = (target[e] - prev[e]) / dt; err[e] = (target[e] - k.oMf[i].p[e]) / dt; }
    mat3_vec(cfg->ee_gain_pos[i], err, ge);
    sp_quat_from_euler_xyz(ref + 3 * i, qr);
    sp_matrix_from_quat(qr, Rref);
    for (int e = 0; e < 9; ++e) dR[e] = (Rref[e] - prevR[e]) / dt;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) sk[3 * r + c] = dR[3 * r] * Rref[3 * c] + dR[3 * r + 1] * Rref[3 * c + 1] + dR[3 * r + 2] * Rref[3 * c + 2];
    b[m + 0] = (ref_vel[0] + ge[0]) * w; b[m + 1] = (ref_vel[1] + ge[1]) * w; b[m + 2] = (ref_vel[2] + ge[2]) * w;
    b[m + 3] = sk[7] * w; b[m + 4] = sk[2] * w; b[m + 5] = sk[3] * w;
    for (int e = 0; e < 3; ++e) prev[e] = target[e];
    for (int e = 0; e < 9; ++e) prevR[e] = Rref[e];
    m += 6;
  }
  double fkq[4] = {0, 0, 0, 1};
  sp_quat_from_matrix(k.oMf[WBC_FRAME_TRUNK].R, fkq);
  if (cfg->task_mask & WBC_TASK_TRUNK) {
    frame_jacobian(t, &k, WBC_FRAME_TRUNK, WBC_RF_WORLD, Jf);
    const double w = cfg->cart_task_weight[5];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < nv; ++c) {
        double s = 0;
        for (int e = 0; e < 6; ++e) s += cfg->trunk_weight[6 * r + e] * Jf[e][c];
        A[m + r][c] = s * w;                                                              /* (W J) w (:488-490) */
      }
    /* calcTargetVelTrunk2 (:948-1015) */
    const double* target = targets + 15;
    double* prev = mem + 60;
    double* oldR = mem + 63;
    double ref_vel[3], err[3], ge[3], r[4], Rref[9], dR[9], sk[9];
    for (int e = 0; e < 3; ++e) { ref_vel[e] = (target[e] - prev[e]) / dt; err[e] = (target[e] - k.oMf[WBC_FRAME_TRUNK].p[e]) / dt; }
    mat3_vec(cfg->trunk_gain_pos, err, ge);
    sp_quat_from_euler_xyz(ref + 15, r);
    sp_matrix_from_quat(r, Rref);
    const double* f = fkq;
    double qe[3];
    qe[0] = (f[3] * r[0]) - (f[0] * r[3]) + (f[1] * r[2]) - (f[2] * r[1]);
    qe[1] = (f[3] * r[1]) - (f[1] * r[3]) - (f[0] * r[2]) + (f[2] * r[0]);
    qe[2] = (f[3] * r[2]) - (f[3] * r[2]) + (f[0] * r[1]) - (f[1] * r[0]);               /* :976 */
    for (int e = 0; e < 9; ++e) dR[e] = (Rref[e] - oldR[e]) / dt;
    mat3_mul(dR, Rref, sk);                                                               /* no transpose (:984) */
    b[m + 0] = (ref_vel[0] + ge[0]) * w; b[m + 1] = (ref_vel[1] + ge[1]) * w; b[m + 2] = (ref_vel[2] + ge[2]) * w;
    b[m + 3] = (sk[7] + cfg->trunk_gain_ori[0] * qe[0]) * w;
    b[m + 4] = (sk[2] + cfg->trunk_gain_ori[1] * qe[1]) * w;
    b[m + 5] = (sk[3] + cfg->trunk_gain_ori[2] * qe[2]) * w;
    for (int e = 0; e < 3; ++e) prev[e] = target[e];
    for (int e = 0; e < 9; ++e) oldR[e] = Rref[e];
    m += 6;
  }
  if (cfg->task_mask & WBC_TASK_JOINT) {                                                 /* qpJointA / qpJointb (:1199-1217) */
    if (cfg->joint_mode != WBC_JOINT_ZERO && cfg->joint_mode != WBC_JOINT_PREV) return WBC_ERR_UNSUPPORTED;
    const double aj = (1.0 / nv) * cfg->joint_task_weight;
    for (int r = 0; r < nv; ++r) {
      for (int c = 0; c < nv; ++c) A[m + r][c] = (r == c) ? aj : 0.0;
      b[m + r] = (cfg->joint_mode == WBC_JOINT_PREV) ? ((1.0 / nv) * q[r < 6 ? r : r + 1]) * cfg->joint_task_weight : 0.0;
    }
    m += nv;
  }
  /* velDamperJointConstraints (:572-637), including the off-by-one quirk when the compat flag is set */
  double lb[NVMAX], ub[NVMAX];
  const int grip = cfg->gripper_joint_id;
  for (int i = 0; i < nv; ++i) {
    const int qidx = i < 6 ? i : i + 1;
    double lo, up;
    if (qidx < 7) { lo = -5; up = 5; }
    else if (qidx >= grip - 2 + 7) { lo = 0; up = 0; }
    else { lo = t->lower[qidx]; up = t->upper[qidx]; }
    const double vel = i < 7 ? 5.0 : t->velocity[i];
    const double c = (cfg->compat_flags & WBC_COMPAT_DAMPER_OFF_BY_ONE) ? q[i] : q[qidx];
    const double coef = cfg->damper_coef, qi = cfg->damper_qi, qs = cfg->damper_qs;
    if (c <= lo + qi) {
      lb[i] = -coef * (c - lo - qs) / (qi - qs);
      if (lb[i] > vel) lb[i] = vel;
      if (lb[i] < -vel) lb[i] = -vel;
    } else lb[i] = -vel;
    if (c >= up - qi) {
      ub[i] = coef * (up - c - qs) / (qi - qs);
      if (ub[i] < -vel) ub[i] = -vel;
      if (ub[i] > vel) ub[i] = vel;
    } else ub[i] = vel;
    if (lb[i] > 0) lb[i] = lb[i] * -1;
    if (ub[i] < 0) ub[i] = ub[i] * -1;
    if (i >= grip - 2 + 6) { lb[i] = 0; ub[i] = 0; }
  }
  /* findConstraints (:764-836): CoM, Trunk, FR, FL, RR, RL, GRIP, then extension rows */
  double C[NCMAX][NVMAX], clb[NCMAX], cub[NCMAX];
  int nC = 0;
  if (cfg->constraint_mask & WBC_CON_COM) {
    double Mt = 0, com[3] = {0, 0, 0};
    double mc[WBC_MAX_JOINTS][3];
    for (int j = 1; j < t->njoints; ++j) {
      double c[3];
      mat3_vec(k.oMi[j].R, t->com[j], c);
      for (int e = 0; e < 3; ++e) { mc[j][e] = c[e] + k.oMi[j].p[e]; com[e] += t->mass[j] * mc[j][e]; }
      Mt += t->mass[j];
    }
    for (int e = 0; e < 3; ++e) com[e] /= Mt;
    for (int c = 0; c < nv; ++c) {
      /* column c moves every joint whose support contains c */
      double sm = 0, s[3] = {0, 0, 0};
      for (int j = 1; j < t->njoints; ++j)
        if ((k.supp[j] >> c) & 1u) { sm += t->mass[j]; for (int e = 0; e < 3; ++e) s[e] += t->mass[j] * mc[j][e]; }
      const double ang[3] = {k.J[3][c], k.J[4][c], k.J[5][c]};
      double cx[3];
      cross3(s, ang, cx);
      C[nC][c] = (sm * k.J[0][c] - cx[0]) / Mt;
      C[nC + 1][c] = (sm * k.J[1][c] - cx[1]) / Mt;
    }
    for (int r = 0; r < 2; ++r) {
      clb[nC + r] = ((k.oMf[2].p[r] - com[r]) / dt) * 0.8;      /* RR */
      cub[nC + r] = ((k.oMf[1].p[r] - com[r]) / dt) * 0.8;      /* FL */
    }
    nC += 2;
  }
  if (cfg->constraint_mask & WBC_CON_TRUNK) {
    frame_jacobian(t, &k, WBC_FRAME_TRUNK, WBC_RF_LOCAL_WORLD_ALIGNED, Jf);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < nv; ++c) C[nC + r][c] = Jf[2 + r][c];
    double eul[3];
    sp_euler_xyz_from_quat(fkq, eul);
    const double* ip = ref + 18;
    const double* ie = ref + 21;
    const double cur[4] = {k.oMf[WBC_FRAME_TRUNK].p[2], eul[0], eul[1], eul[2]};
    const double z_var = ip[2] * 0.25, var = 1.5 * 0.1;
    const double lo[4] = {ip[2] - z_var, ie[0] - var, ie[1] - var, ie[2] - var};
    const double up[4] = {ip[2] + z_var, ie[0] + var, ie[1] + var, ie[2] + var};
    for (int r = 0; r < 4; ++r) { clb[nC + r] = ((lo[r] - cur[r]) / dt) * 0.5; cub[nC + r] = ((up[r] - cur[r]) / dt) * 0.5; }
    nC += 4;
  }
  for (int i = 0; i < 5; ++i) {
    if (!(cfg->constraint_mask & (WBC_CON_FR << i))) continue;
    frame_jacobian(t, &k, i, WBC_RF_WORLD, Jf);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < nv; ++c) C[nC + r][c] = Jf[r][c]; clb[nC + r] = 0; cub[nC + r] = 0; }
    nC += 3;
  }
  for (int e = 0; e < cfg->n_extra_rows; ++e) {
    frame_jacobian(t, &k, cfg->extra_frame[e], cfg->extra_rf[e], Jf);
    for (int c = 0; c < nv; ++c) {
      double s = 0;
      for (int r = 0; r < 6; ++r) s += cfg->extra_coeff[e][r] * Jf[r][c];
      C[nC][c] = s;
    }
    clb[nC] = cfg->extra_lo[e]; cub[nC] = cfg->extra_hi[e];
    nC++;
  }
  /* QP.__init__: H = A^T A, g = -A^T b as dense products (QP_Wrapper.py:17-18) */
  double H[NVMAX][NVMAX], g[NVMAX];
  for (int i = 0; i < nv; ++i) {
    for (int j = 0; j < nv; ++j) { double s = 0; for (int r = 0; r < m; ++r) s += A[r][i] * A[r][j]; H[i][j] = s; }
    double s = 0;
    for (int r = 0; r < m; ++r) s -= A[r][i] * b[r];
    g[i] = s;
  }
  qp_solve(nv, nC, H, g, lb, ub, C, clb, cub, cfg->max_iter > 0 ? cfg->max_iter : 200, qdot, qo);
  if (mem_out) memcpy(mem_out, mem, sizeof(mem));
  if (A_out) for (int r = 0; r < m; ++r) memcpy(A_out + (size_t)r * nv, A[r], sizeof(double) * nv);
  if (b_out) memcpy(b_out, b, sizeof(double) * m);
  return WBC_OK;
}

/* Batched open-loop tick on the CPU.  Optional outputs may be NULL.  nthreads <= 0: all OpenMP threads.
 * Returns WBC_OK or WBC_ERR_*; A_out [N, m, nv] / b_out [N, m] are debug outputs for the parity tests. */
int wbc_oracle_step(const WbcTreeTable* table, const WbcConfig* cfg, const double* q, const double* targets,
                    const double* mem_in, const double* ref, double dt, int64_t N, int nthreads, double* qdot,
                    int32_t* status, int32_t* iters, uint64_t* active_set, double* mem_out, double* A_out, double* b_out) {
  if (!table || !cfg || !q || !targets || !mem_in || !ref || !qdot || !status || !iters || !(dt > 0)) return WBC_ERR_INVALID_ARG;
  if (table->nv > NVMAX || count_nc(cfg) > NCMAX) return WBC_ERR_UNSUPPORTED;
  const int nq = table->nq, nv = table->nv;
  int m = 0;
  for (int t = 0; t < 6; ++t) if (cfg->task_mask & (1 << t)) m += 6;
  if (cfg->task_mask & WBC_TASK_JOINT) m += nv;
  int rc_all = WBC_OK;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < N; ++s) {
    QpOut qo;
    const int rc = tick(table, cfg, q + s * nq, targets + s * WBC_TARGETS_STRIDE, mem_in + s * WBC_MEM_STRIDE,
                        ref + s * WBC_REF_STRIDE, dt, qdot + s * nv, &qo, mem_out ? mem_out + s * WBC_MEM_STRIDE : 0,
                        A_out ? A_out + (size_t)s * m * nv : 0, b_out ? b_out + (size_t)s * m : 0);
    if (rc != WBC_OK) { rc_all = rc; continue; }
    status[s] = qo.status; iters[s] = qo.iters;
    if (active_set) { active_set[2 * s] = qo.act_box; active_set[2 * s + 1] = qo.act_rows; }
  }
  return rc_all;
}

int wbc_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
