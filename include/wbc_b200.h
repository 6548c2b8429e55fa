/* wbc_b200.h -- C ABI of libwbc_b200.so: batched whole-body-control hot path on B200 (sm_100a).
 *
 * The reference (joey156/MECH5845M-WBC-for-Legged-Manipulator) has no FFI layer of its own: its
 * hot path is Python sitting on Pinocchio (Boost.Python) and qpOASES (Cython).  The entry points
 * below are what a binding for that path has to reach; each cites the reference interface it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns an int status (WBC_OK == 0); wbc_last_error() gives the message of
 *     the last failure on the calling thread.  No exceptions cross the ABI.
 *   - every array pointer is a BORROWED DEVICE pointer (row-major, float64 unless noted);
 *     NULL means "not requested" for optional outputs.  Nothing is allocated inside the calls.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it.
 *   - a WbcModel is immutable after creation and may be shared by streams/threads (wbc_step_host excepted: it owns
 *     a copy / compute pipeline inside the handle).  It lives on the CUDA device that was current when it was created;
 *     every later call must be made with that device current (WBC_ERR_INVALID_ARG otherwise).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Batched layout (N = number of robot states, leading dimension of every tensor):
 *   q        [N, nq]   x y z qx qy qz qw | joints in Pinocchio order   (current_joint_config, Robot_Wrapper4.py:402)
 *   targets  [N, 18]   ee_target[5][3] (FR FL RR RL GRIP) | trunk_target[3]          (runWBC args, :1330)
 *   mem      [N, 72]   prev_EE_pos[5][3] | prev_EE_CoM_rot[5][9] | prev_trunk_ref[3] | old_ref_trunk_rot[9]
 *                                                                              (:133-140, :995-996, :1151-1152)
 *   ref      [N, 24]   default_EE_ori[5][3] | default_trunk_ori[3] | initial_trunk_pos[3] | initial_trunk_ori_euler[3]
 *                                                                              (:363-367, :379-383)
 */
#ifndef WBC_B200_H
#define WBC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WBC_ABI_VERSION 2

#define WBC_MAX_JOINTS 32
#define WBC_MAX_NV 32
#define WBC_MAX_NQ 33
#define WBC_MAX_FRAMES 16
#define WBC_NUM_EE 5
#define WBC_FRAME_TRUNK 5      /* frame slots 0..4 = EE frames FR FL RR RL GRIP, slot 5 = trunk (imu) frame */
#define WBC_MAX_NC 32          /* rows of C handled by the QP kernels */
#define WBC_MAX_EXTRA_ROWS 24  /* extension rows (friction pyramid / torque proxy); not in the reference */

#define WBC_TARGETS_STRIDE 18
#define WBC_MEM_STRIDE 72
#define WBC_REF_STRIDE 24

/* status codes */
#define WBC_OK 0
#define WBC_ERR_INVALID_ARG 1
#define WBC_ERR_CUDA 2
#define WBC_ERR_UNSUPPORTED 3

/* joint types in WbcTreeTable.jtype */
#define WBC_JT_UNIVERSE 0
#define WBC_JT_FREEFLYER 1
#define WBC_JT_REVOLUTE 2
#define WBC_JT_PRISMATIC 3

/* pin.ReferenceFrame */
#define WBC_RF_WORLD 0
#define WBC_RF_LOCAL 1
#define WBC_RF_LOCAL_WORLD_ALIGNED 2

/* task bits: setTasks(Trunk, FR, FL, RR, RL, Grip, Joint)  (Robot_Wrapper4.py:176); row order of A is
 * FR FL RR RL GRIP Trunk Joint (:839-876, :1271-1280) */
#define WBC_TASK_FR 1
#define WBC_TASK_FL 2
#define WBC_TASK_RR 4
#define WBC_TASK_RL 8
#define WBC_TASK_GRIP 16
#define WBC_TASK_TRUNK 32
#define WBC_TASK_JOINT 64

/* joint-task mode: True / "PREV" / "MANI" / "HYBRID"  (Robot_Wrapper4.py:1209-1268) */
#define WBC_JOINT_ZERO 0
#define WBC_JOINT_PREV 1
#define WBC_JOINT_MANI 2
#define WBC_JOINT_HYBRID 3

/* constraint bits: setConstraints(CoM, Trunk, FR, FL, RR, RL, Grip)  (:186); row order of C is
 * CoM(2) Trunk(4) FR(3) FL(3) RR(3) RL(3) GRIP(3)  (:764-836) */
#define WBC_CON_COM 1
#define WBC_CON_TRUNK 2
#define WBC_CON_FR 4
#define WBC_CON_FL 8
#define WBC_CON_RR 16
#define WBC_CON_RL 32
#define WBC_CON_GRIP 64

/* compat_flags: reference quirks reproduced by default */
#define WBC_COMPAT_DAMPER_OFF_BY_ONE 1   /* velDamper compares joint j's limits with q[j-1]  (:597-613) */

/* WbcStepIO.flags */
#define WBC_STEP_FLAG_PLAIN_INTEGRATE 1  /* q_next = integrate(q, qdot dt) only: the bootstrap loop's update (:325, :446-447) */

/* QP status (per state); NOT_PD is a flag or-ed onto the others */
#define WBC_QP_SOLVED 0
#define WBC_QP_MAXITER 1
#define WBC_QP_INFEASIBLE 2
#define WBC_QP_NOT_PD 4

/* Flat kinematic tree: replaces pin.Model built by pin.buildModelFromUrdf(urdf, JointModelFreeFlyer())
 * (Robot_Wrapper4.py:21) plus the name->index look-ups of RobotModel.__init__ (:30-52).
 * Joint 0 = universe, joint 1 = free-flyer root.  Frames are (parent joint, SE3 offset). */
typedef struct WbcTreeTable {
  int32_t njoints, nq, nv, nframes;
  int32_t parent[WBC_MAX_JOINTS];
  int32_t jtype[WBC_MAX_JOINTS];
  int32_t idx_q[WBC_MAX_JOINTS];
  int32_t idx_v[WBC_MAX_JOINTS];
  double placement_R[WBC_MAX_JOINTS][9];   /* row-major */
  double placement_p[WBC_MAX_JOINTS][3];
  double axis[WBC_MAX_JOINTS][3];
  int32_t frame_parent[WBC_MAX_FRAMES];
  double frame_R[WBC_MAX_FRAMES][9];
  double frame_p[WBC_MAX_FRAMES][3];
  double lower[WBC_MAX_NQ];                /* model.lowerPositionLimit (nq) */
  double upper[WBC_MAX_NQ];                /* model.upperPositionLimit (nq) */
  double velocity[WBC_MAX_NV];             /* model.velocityLimit (nv) */
  double mass[WBC_MAX_JOINTS];             /* lumped body mass per joint (fixed links merged) */
  double com[WBC_MAX_JOINTS][3];           /* its centre of mass in the joint frame */
} WbcTreeTable;

/* Controller configuration: the attributes RobotModel keeps as Python state (:72-125, :176-193,
 * :574-576, :1415-1464).  Passed by value with every call, so setTasks / setConstraints /
 * staticReachMode are just edits of the caller's struct. */
typedef struct WbcConfig {
  int32_t task_mask;
  int32_t joint_mode;
  int32_t constraint_mask;
  int32_t compat_flags;
  int32_t gripper_joint_id;                /* end_effector_index_list_joint[4]: locks v >= id-2+6 (:628) */
  int32_t arm_base_id;                     /* getJointId(G_base) (:37), used by "HYBRID" */
  int32_t max_iter;                        /* QP working-set iteration cap (reference nWSR = 100000, QP_Wrapper.py:20) */
  int32_t n_extra_rows;                    /* extension rows appended to C (0 = reference behaviour) */
  double ee_weight[WBC_NUM_EE][36];        /* EE_weight[i], 6x6 row-major (:80) */
  double trunk_weight[36];                 /* :74 */
  double cart_task_weight[6];              /* cart_task_weight_EE_list[0..4], cart_task_weight_Trunk (:86-92) */
  double joint_task_weight;                /* :93 */
  double ee_gain_pos[WBC_NUM_EE][9];       /* EE_gains[i][0:3,0:3] -- list order quirk resolved by the host (:125) */
  double trunk_gain_pos[9];                /* trunk_gain[0:3,0:3] (:119) */
  double trunk_gain_ori[3];                /* diag(trunk_gain[3:,3:]) (:982) */
  double damper_coef, damper_qi, damper_qs; /* velocity damper constants (:574-576) */
  /* extension rows (NOT in the reference): row = coeff . J_frame(ref_frame), lo <= row . qdot <= hi */
  int32_t extra_frame[WBC_MAX_EXTRA_ROWS];
  int32_t extra_rf[WBC_MAX_EXTRA_ROWS];
  double extra_coeff[WBC_MAX_EXTRA_ROWS][6];
  double extra_lo[WBC_MAX_EXTRA_ROWS];
  double extra_hi[WBC_MAX_EXTRA_ROWS];
} WbcConfig;

typedef struct WbcModel WbcModel;

/* One batched WBC tick: everything RobotModel.runWBC does between reading its arguments and
 * returning joint targets (Robot_Wrapper4.py:1330-1412), for N states at once. */
typedef struct WbcStepIO {
  /* inputs */
  const double* q;          /* [N, nq] */
  const double* targets;    /* [N, 18] */
  const double* mem_in;     /* [N, 72] */
  const double* ref;        /* [N, 24] */
  const double* imu_quat;   /* [N, 4] or NULL: base orientation fed back after the step (runWBC base_config, :1402) */
  double dt;                /* the reference measures this with a busy-wait (:1338-1342); explicit here */
  int64_t flags;            /* WBC_STEP_FLAG_* */
  /* outputs */
  double* qdot;             /* [N, nv]  QP solution */
  int32_t* status;          /* [N] WBC_QP_* */
  int32_t* iters;           /* [N] working-set iterations */
  uint64_t* active_set;     /* [N, 2] or NULL: word0 = box bounds (bit 2k lower, 2k+1 upper of x_k), word1 = rows of C */
  double* mem_out;          /* [N, 72] or NULL: task memory after the tick (may alias mem_in) */
  double* q_next;           /* [N, nq] or NULL: integrate + base estimate (:1397-1402, :1297-1327); may alias q */
  double* joint_targets;    /* [N, nq - 7] or NULL: q_next[7:], the joint position targets runWBC returns as its five
                               slices FL, FR, RL, RR, grip (:1405-1412); needs q_next */
} WbcStepIO;

/* Accessor / debug outputs of the assembly stage: what qpA, qpb, velDamperJointConstraints and
 * findConstraints return (Robot_Wrapper4.py:1271, 1283, 572, 764) plus H, g of QP.__init__
 * (QP_Wrapper.py:17-18).  Any pointer may be NULL. */
typedef struct WbcAssembleOut {
  double* A;      /* [N, m, nv], m = 6 * (#cartesian tasks) + (joint task ? nv : 0) */
  double* b;      /* [N, m] */
  double* lb;     /* [N, nv] */
  double* ub;     /* [N, nv] */
  double* C;      /* [N, nC, nv]  (rows = constraints; the reference returns C.T, :836) */
  double* Clb;    /* [N, nC] */
  double* Cub;    /* [N, nC] */
  double* H;      /* [N, nv, nv] */
  double* g;      /* [N, nv] */
} WbcAssembleOut;

int wbc_abi_version(void);
const char* wbc_last_error(void);

/* pin.buildModelFromUrdf + createData (Robot_Wrapper4.py:21-23): uploads the table to the current device. */
int wbc_model_create(const WbcTreeTable* table, WbcModel** out_model);
void wbc_model_destroy(WbcModel* model);

/* rows of A / C implied by a config (host helper; m and nC of the layouts above) */
int wbc_config_rows(const WbcConfig* cfg, int32_t nv, int32_t* m_rows, int32_t* nc_rows);

/* updateState + getFrameJacobian (Robot_Wrapper4.py:387-428, 458-488, 641-758):
 * forwardKinematics / computeJointJacobians / updateFramePlacements for N states, then for each of
 * the `nsel` frame slots: oMf (R row-major 9 | p 3) and the 6 x nv frame Jacobian in `ref_frame`.
 * out_oMf [N, nsel, 12], out_J [N, nsel, 6, nv]; either may be NULL. */
int wbc_fk_jac(const WbcModel* model, const double* q, int64_t N, const int32_t* frame_slots, int32_t nsel,
               int32_t ref_frame, double* out_oMf, double* out_J, void* stream);

/* data.oMi and data.J (WORLD joint Jacobian, computeJointJacobians :403): out_oMi [N, njoints, 12], out_J [N, 6, nv]. */
int wbc_joint_jacobians(const WbcModel* model, const double* q, int64_t N, double* out_oMi, double* out_J, void* stream);

/* initialiseWBC (Robot_Wrapper4.py:354-383): FK at q, then snapshot task memory and references. */
int wbc_init_memory(const WbcModel* model, const double* q, int64_t N, double* mem_out, double* ref_out, void* stream);

/* jointVelocitiestoConfig (Robot_Wrapper4.py:440-441): q_out = pin.integrate(model, q, v * scale); q_out may alias q. */
int wbc_integrate(const WbcModel* model, const double* q, const double* v, int64_t N, double scale, double* q_out,
                  void* stream);

/* updateState(joint_config, imu_data, running=True) (Robot_Wrapper4.py:387-428) without the accessor refresh: the base
 * orientation of q is replaced by imu_quat (NULL: kept), FK, trunkWorldPos (:1297-1327: base xyz re-estimated from the
 * four foot targets, targets [N, 18] as in WbcStepIO), and q_out = [estimated xyz, quaternion, joints].  base_out [N, 3]
 * (or NULL) receives just the estimate, i.e. what trunkWorldPos() returns.  q_out may be NULL or alias q. */
int wbc_base_estimate(const WbcModel* model, const double* q, const double* imu_quat, const double* targets, int64_t N,
                      double* q_out, double* base_out, void* stream);

/* qpA / qpb / velDamperJointConstraints / findConstraints (+ H, g) without solving. */
int wbc_assemble(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, int64_t N,
                 const WbcAssembleOut* out, void* stream);

/* QP(A, b, lb, ub, C, Clb, Cub).solveQP() / solveQPHotstart (QP_Wrapper.py:10-73), batched.
 * Either (A [N, m, nv], b [N, m]) or (H [N, nv, nv], g [N, nv]) is given (the other pair NULL).
 * C [N, nC, nv] rows = constraints, may be NULL with nC = 0 (QProblemB path). */
int wbc_qp_solve(int64_t N, int32_t nv, int32_t m, int32_t nC, const double* A, const double* b, const double* H,
                 const double* g, const double* lb, const double* ub, const double* C, const double* Clb,
                 const double* Cub, int32_t max_iter, double* x, int32_t* status, int32_t* iters,
                 uint64_t* active_set, void* stream);

/* the fused tick */
int wbc_step(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, int64_t N, void* stream);

/* element type of the floating-point HOST arrays of wbc_step_host */
#define WBC_HOST_F64 0
#define WBC_HOST_F32 1         /* optional FP32 I/O mode: float32 arrays on the host side (half the PCIe bytes); the tick
                                  itself still computes in float64.  Agreement with the float64 call: <= 1e-4 */

/* WbcHostIO.flags.  DELTA_INPUTS (float32, closed loop only): `targets` holds the INCREMENT of every target over the
 * previous tick's target (the task memory's prev_EE_pos / prev_trunk_ref, Robot_Wrapper4.py:995, :1151) and `imu_quat`
 * the increment of the base quaternion over the resident q[3:7]; the tick adds them back in float64.  A float32
 * increment of ~1e-3 carries an absolute error of ~1e-10, a float32 position of ~0.5 m one of ~3e-8 -- which the target
 * laws multiply by 1 / dt = 500.  With increments the FP32 I/O mode agrees with the float64 call to < 1e-4 in qdot;
 * with absolute float32 targets only the joint position targets do (qdot: ~1e-5 typical, ~3e-3 worst case). */
#define WBC_HOST_FLAG_DELTA_INPUTS 1

/* Host-side arrays of one tick for wbc_step_host: what a caller of the reference holds as NumPy arrays
 * (runWBC's arguments and return values, Robot_Wrapper4.py:1330-1412).  Page-locked memory makes the copies
 * asynchronous; pageable memory works but serialises them.  An input left NULL is not copied: it is resident in the
 * device buffer of `io` already (the configuration, the task memory and the per-robot references, which the reference
 * keeps as attributes of the controller object).  An output left NULL is not copied back. */
typedef struct WbcHostIO {
  const void* q;            /* [N, nq] or NULL */
  const void* targets;      /* [N, 18] */
  const void* mem_in;       /* [N, 72] or NULL */
  const void* ref;          /* [N, 24] or NULL */
  const void* imu_quat;     /* [N, 4] or NULL: runWBC's base_config (:1330, :1402); needs io->q_next */
  void* qdot;               /* [N, nv] or NULL */
  int32_t* status;          /* [N] or NULL */
  int32_t* iters;           /* [N] or NULL */
  void* joint_targets;      /* [N, nq - 7] or NULL: what runWBC returns (:1405-1412); needs io->q_next */
  int32_t dtype;            /* WBC_HOST_F64 / WBC_HOST_F32: element type of q, targets, mem_in, ref, imu_quat, qdot, joint_targets */
  int32_t flags;            /* WBC_HOST_FLAG_* */
} WbcHostIO;

/* One tick with host buffers: runWBC as its callers see it (NumPy arrays in, NumPy arrays out).  `io` names the device
 * buffers: all of wbc_step's required members; for every travelling array its device twin is staging space whose
 * contents are unspecified afterwards (io->imu_quat / io->joint_targets are needed as staging only when chunks >= 1).
 *   open loop    io->q_next == NULL, io->mem_out == NULL: nothing on the device is advanced;
 *   closed loop  io->q_next == io->q and io->mem_out == io->mem_in (both resident: host->q and host->mem_in NULL): the
 *                configuration and the task memory are advanced in place exactly as runWBC mutates its object
 *                (:995-996, :1151-1152, :1397-1402); per tick only the IMU quaternion and the targets travel in and the
 *                joint targets / status out.  K calls equal wbc_rollout over the same K ticks.
 *   chunks < 0   if every host array is page-locked the kernel reads the inputs from and writes the outputs to host memory
 *                directly (zero-copy, one launch on `stream`); otherwise as chunks = 8;
 *   chunks == 0  self-tuning: as chunks < 0, but with page-locked arrays and N >= 4096 the model handle measures both host
 *                paths on the first four calls of a problem shape (two zero-copy launches, two calls with 8 staged slices;
 *                the second of each is timed with events on `stream`) and runs the faster one from then on -- one GPU
 *                alone is faster zero-copy, eight GPUs pulling on one host NUMA node are faster staged.  Results do not
 *                depend on the path.  wbc_step_host_path() reports the decision;
 *   chunks >= 1  staged: the batch is cut into `chunks` wave-aligned slices whose host->device copies, kernel and
 *                device->host copies overlap on three streams owned by the model (created on first use).
 * io->active_set must be NULL.  Asynchronous: `stream` is ordered before the first access and after the last;
 * synchronise it before reading the host outputs.  Not re-entrant per model handle. */
int wbc_step_host(WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, const WbcHostIO* host, int64_t N,
                  int32_t chunks, void* stream);

/* the host path wbc_step_host's self-tuning mode (chunks == 0) settled on for the last problem shape: -1 undecided (fewer
 * than five calls, or the mode has not been used), 0 zero-copy, 8 staged slices; -2: null handle */
int wbc_step_host_path(const WbcModel* model);

/* Closed loop: K consecutive ticks (the tick loop of sim3.py:287-327 around runWBC, Robot_Wrapper4.py:1330-1412) for
 * N robots, configuration and task memory advanced in place on the device.  With the usual constraint set (trunk box + four
 * feet, <= 16 rows, any joint-task mode) the whole horizon is ONE persistent launch: a robot stays with one warp for
 * all K ticks, so the ticks need no grid-wide synchronisation and only the last one has a drain tail; every other
 * configuration runs one fused launch per tick.  Results are identical either way.
 * io->q is read and overwritten (io->q_next must be NULL or equal to io->q), io->mem_in likewise (io->mem_out NULL or
 * equal); io->targets is ignored: tick k reads targets_traj[k] ([K, N, 18]).  imu_traj [K, N, 4] or NULL (base
 * orientation fed back after each tick).  qdot / status / iters / active_set hold the last tick. */
int wbc_rollout(const WbcModel* model, const WbcConfig* cfg, const WbcStepIO* io, const double* targets_traj,
                const double* imu_traj, int32_t K, int64_t N, void* stream);

/* launch geometry the step kernel uses on the current device (for bench reporting) */
int wbc_step_launch_info(const WbcModel* model, int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* regs);

/* DFMA-saturating microkernel: returns achieved FP64 FLOP/s (the roofline denominator MEASURED_PEAKS.json lacks) */
int wbc_measure_fp64_peak(double* flops_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WBC_B200_H */
