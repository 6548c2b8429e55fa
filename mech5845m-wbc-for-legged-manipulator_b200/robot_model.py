"""Batched drop-in for ``wrappers/Robot_Wrapper4.py: class RobotModel`` (reference :18-1500).

Same constructor arguments, same method names, same argument meaning -- but every array carries
a leading batch dimension N and lives in a float64 CUDA tensor; all arithmetic runs in the
hand-written kernels of ``libwbc_b200.so`` through the C ABI (``_cabi.py``).  PyTorch is used for
device memory and streams only.  There is no CPU fallback: constructing a model without a CUDA
device (or without the built library) raises.

Departures from the reference, all forced:
  * ``dt`` is an explicit attribute (the reference busy-waits on the wall clock, :1338-1342);
  * nothing is printed from the hot path (:1075-1085);
  * the 2000-tick bootstrap the reference runs inside its constructor (:161) is opt-in
    (``run_bootstrap=True`` or ``setInitialState()``);
  * QP status, iteration counts and active sets are kept (``last_status`` ...) instead of discarded.
Reference quirks are reproduced by default (``compat_damper_off_by_one`` etc.).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi as cabi
from .tree_table import TreeTable

EE_FRAME_NAMES = ["FR_foot_fixed", "FL_foot_fixed", "RR_foot_fixed", "RL_foot_fixed", "gripper_bar"]
EE_JOINT_NAMES = ["FR_calf_joint", "FL_calf_joint", "RR_calf_joint", "RL_calf_joint", "gripper"]
HIP_WAIST_JOINT_NAMES = ["FR_hip_joint", "FL_hip_joint", "RR_hip_joint", "RL_hip_joint", "waist"]

_JOINT_MODES = {True: cabi.JOINT_ZERO, "PREV": cabi.JOINT_PREV, "MANI": cabi.JOINT_MANI, "HYBRID": cabi.JOINT_HYBRID}


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class _SE3View:
    """``robot_data.oMf[i]``: ``.translation`` [N, 3], ``.rotation`` [N, 3, 3] (views into the FK output)."""

    def __init__(self, block):
        self.rotation = block[:, :9].reshape(-1, 3, 3)
        self.translation = block[:, 9:12]


class _FrameAccessor:
    def __init__(self, owner):
        self._o = owner

    def __getitem__(self, frame_id):
        slot = self._o._frame_slot.get(int(frame_id))
        if slot is None:
            raise KeyError(f"frame {frame_id} was not uploaded to the device model (hot frames only)")
        return _SE3View(self._o._oMf[:, slot])


class _JointAccessor:
    def __init__(self, owner):
        self._o = owner

    def __getitem__(self, joint_id):
        return _SE3View(self._o._oMi[:, int(joint_id)])


class _Data:
    def __init__(self, owner):
        self.oMf = _FrameAccessor(owner)
        self.oMi = _JointAccessor(owner)


class LinearTrajectory:
    """Batched ``klampt.model.trajectory.Trajectory(milestones=...)`` as the drivers use it (``sim3.py:207-228``,
    Robot_Wrapper4.py:264-283): piecewise linear, knot i at t = i, clamped at both ends.  ``milestones`` [K, N, 3]."""

    def __init__(self, milestones):
        self.m = milestones

    def eval(self, t):
        K = self.m.shape[0]
        if t <= 0:
            return self.m[0].clone()
        if t >= K - 1:
            return self.m[-1].clone()
        i = int(np.floor(t))
        u = t - i
        return self.m[i] + u * (self.m[i + 1] - self.m[i])


class HostDeltaEncoder:
    """Host-side twin of the FP32 increment inputs of ``RobotModel.step_host(..., delta_inputs=True)``: keeps float64
    mirrors of what the device holds as "previous targets" (task memory ``prev_EE_pos`` / ``prev_trunk_ref``) and as base
    quaternion, hands out float32 increments and advances the mirrors with exactly the float64 additions the kernel
    performs, so host and device never drift apart."""

    def __init__(self, robot):
        m = robot._mem
        self.prev_targets = torch.cat((m[:, 0:15], m[:, 60:63]), dim=1).cpu().clone()
        self.quat = robot.current_joint_config[:, 3:7].cpu().clone()

    def encode(self, targets, imu=None, out_targets=None, out_imu=None):
        """targets [N, 18] (and imu [N, 4]) float64 CPU tensors -> float32 increments (written into the given pinned
        buffers when supplied)."""
        dt = (targets - self.prev_targets).to(torch.float32)
        self.prev_targets += dt.double()
        if out_targets is not None:
            out_targets.copy_(dt); dt = out_targets
        if imu is None:
            return dt, None
        di = (imu - self.quat).to(torch.float32)
        self.quat += di.double()
        if out_imu is not None:
            out_imu.copy_(di); di = out_imu
        return dt, di


class RobotModel:
    def __init__(self, urdf_path, mesh_dir_path=None, EE_frame_names=EE_FRAME_NAMES, EE_joint_names=EE_JOINT_NAMES,
                 G_base="waist", imu="imu_joint", FR_hip_joint="FR_hip_joint",
                 hip_waist_joint_names=HIP_WAIST_JOINT_NAMES, foot_offset=False, *,
                 batch=1, device=None, dt=0.002, run_bootstrap=False):
        """Robot_Wrapper4.py:19-173.  ``urdf_path``: URDF file, tree-table JSON, table name or ``TreeTable``."""
        if not torch.cuda.is_available():
            raise cabi.WbcError("RobotModel needs a CUDA device (B200): the hot path has no CPU fallback")
        self._lib = cabi.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.robot_model = urdf_path if isinstance(urdf_path, TreeTable) else TreeTable.load(urdf_path)
        t = self.robot_model
        self.N = int(batch)
        self.joint_names = t.joint_names
        self.foot_radius = 0

        self.trunk_frame_index = t.getFrameId(imu, "FIXED_JOINT")                        # :30
        self.EE_frame_names = list(EE_frame_names)
        self.EE_joint_names = list(EE_joint_names)
        self.hip_waist_joint_names = list(hip_waist_joint_names)
        self.arm_base_id = t.getJointId(G_base)                                          # :37
        self.arm_base_frame_id = t.getFrameId(G_base, "JOINT")
        self.FR_hip_joint = t.getJointId(FR_hip_joint)
        self.n_velocity_dimensions = t.nv
        self.n_configuration_dimensions = t.nq
        self.n_of_EE = 5
        self.end_effector_index_list_frame = [t.getFrameId(n, "FIXED_JOINT") for n in self.EE_frame_names]   # :46-52
        self.end_effector_index_list_joint = [t.getJointId(n) for n in self.EE_joint_names]
        self.hip_waist_joint_index_list_frame = [t.getFrameId(n, "JOINT") for n in self.hip_waist_joint_names]
        for name, fid in zip(self.EE_frame_names + [imu], self.end_effector_index_list_frame + [self.trunk_frame_index]):
            if fid >= t.nframes:
                raise ValueError(f"frame {name!r} not found in the model")
        if foot_offset is True:                                                          # :55-58
            self.foot_radius = t.collision_geoms[8][2]

        # device model: slots 0..4 EE frames, 5 trunk, then hip/waist JOINT frames and the arm base frame
        slots = self.end_effector_index_list_frame + [self.trunk_frame_index]
        for fid in self.hip_waist_joint_index_list_frame + [self.arm_base_frame_id]:
            if fid < t.nframes and fid not in slots:
                slots.append(fid)
        self._slot_frames = slots
        self._frame_slot = {fid: s for s, fid in enumerate(slots)}
        self._table = self._make_table(t, slots)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_model_create(C.byref(self._table), C.byref(handle)))
        self._model = handle

        # weights :72-93
        self.trunk_weight = np.identity(6) * 1
        self.FR_weight = np.identity(6) * 1
        self.FL_weight = np.identity(6) * 1
        self.RR_weight = np.identity(6) * 1
        self.RL_weight = np.identity(6) * 1
        self.grip_weight = np.identity(6) * 1
        self.EE_weight = [self.FR_weight, self.FL_weight, self.RR_weight, self.RL_weight, self.grip_weight]
        self.cart_task_weight_EE_list = [1, 1, 1, 1, 1]
        self.cart_task_weight_Trunk = 1
        self.joint_task_weight = 0.05
        self.setTasks()
        self.setConstraints()
        self.step_time = dt
        self.dt = dt
        # gains :119-125 -- list order FL, FR, RL, RR, GRIP, indexed with the FR, FL, RR, RL, GRIP index (quirk D.8)
        self.trunk_gain = np.identity(6) * 0.5
        self.FL_gain = np.identity(6) * 0.5
        self.FR_gain = np.identity(6) * 0.5
        self.RL_gain = np.identity(6) * 0.5
        self.RR_gain = np.identity(6) * 0.5
        self.GRIP_gain = np.identity(6) * 0.5
        self.EE_gains = [self.FL_gain, self.FR_gain, self.RL_gain, self.RR_gain, self.GRIP_gain]
        self.damper = (0.01, 0.026, 0.015)                                               # :574-576
        self.compat_damper_off_by_one = True
        self.max_qp_iterations = 200
        self.extra_rows = []        # extension rows (frame_slot, ref_frame, coeff[6], lo, hi); not in the reference

        # batched state
        N, nq, nv = self.N, t.nq, t.nv
        f64 = dict(dtype=torch.float64, device=self.device)
        self.current_joint_config = torch.zeros(N, nq, **f64)
        self.current_joint_config[:, 6] = 1.0                                             # pin.neutral (:66)
        self.previous_joint_config = self.current_joint_config.clone()
        self._mem = torch.zeros(N, cabi.MEM_STRIDE, **f64)
        self._ref = torch.zeros(N, cabi.REF_STRIDE, **f64)
        self._oMf = torch.zeros(N, len(slots), 12, **f64)
        self._oMi = torch.zeros(N, t.njoints, 12, **f64)
        self.J = torch.zeros(N, 6, nv, **f64)
        self._slot_sel = (C.c_int32 * len(slots))(*range(len(slots)))
        self.robot_data = _Data(self)
        self.qdot = torch.zeros(N, nv, **f64)
        self.last_status = torch.zeros(N, dtype=torch.int32, device=self.device)
        self.last_iters = torch.zeros(N, dtype=torch.int32, device=self.device)
        self.last_active_set = torch.zeros(N, 2, dtype=torch.int64, device=self.device)
        self._targets = torch.zeros(N, cabi.TARGETS_STRIDE, **f64)
        self.EE_A_list = [0, 0, 0, 0, 0]
        self.EE_b_list = [0, 0, 0, 0, 0]
        self.trunk_A = 0
        self.trunk_B = 0
        self.firstQP = True
        self.initialised = False
        self.updateState(self.current_joint_config, feedback=False)
        arm_base = self.robot_data.oMf[self.arm_base_frame_id].translation if self.arm_base_frame_id in self._frame_slot else None
        grip = self.robot_data.oMf[self.end_effector_index_list_frame[4]].translation
        self.arm_reach = float((grip - arm_base)[0].sum()) if arm_base is not None else 0.0   # :70
        if run_bootstrap:
            self.setInitialState()
        self.initialised = True
        mem, _ = self._log_previous_states()                                             # :168-173
        self._mem.copy_(mem)

    def __del__(self):
        try:
            if getattr(self, "_model", None):
                self._lib.wbc_model_destroy(self._model)
                self._model = None
        except Exception:
            pass

    # ---------------------------------------------------------------------------------- table / config marshalling
    @staticmethod
    def _make_table(t, slots):
        if t.njoints > cabi.MAX_JOINTS or t.nv > cabi.MAX_NV or len(slots) > cabi.MAX_FRAMES:
            raise ValueError("model exceeds the device table limits (32 joints, nv <= 32, 16 frames)")
        T = cabi.WbcTreeTable()
        T.njoints, T.nq, T.nv, T.nframes = t.njoints, t.nq, t.nv, len(slots)
        for j in range(t.njoints):
            T.parent[j], T.jtype[j], T.idx_q[j], T.idx_v[j] = t.parent[j], t.jtype[j], t.idx_q[j], t.idx_v[j]
            for k in range(9):
                T.placement_R[j][k] = float(t.placement_R[j].reshape(-1)[k])
            for k in range(3):
                T.placement_p[j][k] = float(t.placement_p[j][k])
                T.axis[j][k] = float(t.axis[j][k])
                T.com[j][k] = float(t.com[j][k])
            T.mass[j] = float(t.mass[j])
        for s, fid in enumerate(slots):
            T.frame_parent[s] = t.frame_parent[fid]
            for k in range(9):
                T.frame_R[s][k] = float(t.frame_R[fid].reshape(-1)[k])
            for k in range(3):
                T.frame_p[s][k] = float(t.frame_p[fid][k])
        for i in range(t.nq):
            T.lower[i], T.upper[i] = float(t.lower[i]), float(t.upper[i])
        for i in range(t.nv):
            T.velocity[i] = float(t.velocity[i])
        return T

    def _task_mask(self):
        m = 0
        for flag, bit in ((self.task_active_FR_foot, cabi.TASK_FR), (self.task_active_FL_foot, cabi.TASK_FL),
                          (self.task_active_RR_foot, cabi.TASK_RR), (self.task_active_RL_foot, cabi.TASK_RL),
                          (self.task_active_GRIP, cabi.TASK_GRIP), (self.task_active_Trunk, cabi.TASK_TRUNK)):
            if flag is True:
                m |= bit
        j = self.task_active_Joint
        if j is True or (isinstance(j, str) and j in ("PREV", "MANI", "HYBRID")):
            m |= cabi.TASK_JOINT
        return m

    def _constraint_mask(self):
        m = 0
        for flag, bit in ((self.const_active_CoM, cabi.CON_COM), (self.const_active_Trunk, cabi.CON_TRUNK),
                          (self.const_active_FR_foot, cabi.CON_FR), (self.const_active_FL_foot, cabi.CON_FL),
                          (self.const_active_RR_foot, cabi.CON_RR), (self.const_active_RL_foot, cabi.CON_RL),
                          (self.const_active_GRIP, cabi.CON_GRIP)):
            if flag is True:
                m |= bit
        return m

    def _config(self, task_mask=None, constraint_mask=None):
        """The controller's settings as the WbcConfig struct that travels with every call (bulk copies through NumPy
        views of the ctypes arrays: this runs once per tick)."""
        c = cabi.WbcConfig()
        view = np.ctypeslib.as_array
        c.task_mask = self._task_mask() if task_mask is None else task_mask
        j = self.task_active_Joint
        c.joint_mode = _JOINT_MODES.get(j if isinstance(j, str) else True, cabi.JOINT_ZERO)
        c.constraint_mask = self._constraint_mask() if constraint_mask is None else constraint_mask
        c.compat_flags = cabi.COMPAT_DAMPER_OFF_BY_ONE if self.compat_damper_off_by_one else 0
        c.gripper_joint_id = self.end_effector_index_list_joint[4]
        c.arm_base_id = self.arm_base_id
        c.max_iter = int(self.max_qp_iterations)
        view(c.ee_weight)[:] = np.asarray(self.EE_weight, dtype=np.float64).reshape(5, 36)
        # EE_gains[frame_index] (:908): the list is built FL, FR, RL, RR, GRIP (:125) -- indexed as given (quirk D.8)
        view(c.ee_gain_pos)[:] = np.asarray(self.EE_gains, dtype=np.float64)[:, 0:3, 0:3].reshape(5, 9)
        view(c.trunk_weight)[:] = np.asarray(self.trunk_weight, dtype=np.float64).reshape(36)
        ctw = view(c.cart_task_weight)
        ctw[:5] = np.asarray(self.cart_task_weight_EE_list, dtype=np.float64)
        ctw[5] = float(self.cart_task_weight_Trunk)
        c.joint_task_weight = float(self.joint_task_weight)
        G = np.asarray(self.trunk_gain, dtype=np.float64)
        view(c.trunk_gain_pos)[:] = G[0:3, 0:3].reshape(9)
        view(c.trunk_gain_ori)[:] = np.diagonal(G)[3:6]
        c.damper_coef, c.damper_qi, c.damper_qs = self.damper
        if len(self.extra_rows) > cabi.MAX_EXTRA:
            raise ValueError(f"at most {cabi.MAX_EXTRA} extension rows")
        c.n_extra_rows = len(self.extra_rows)
        for e, (slot, rf, coeff, lo, hi) in enumerate(self.extra_rows):
            c.extra_frame[e], c.extra_rf[e], c.extra_lo[e], c.extra_hi[e] = int(slot), int(rf), float(lo), float(hi)
            for k in range(6):
                c.extra_coeff[e][k] = float(coeff[k])
        return c

    def _rows(self, cfg):
        m, nc = C.c_int32(), C.c_int32()
        cabi.check(self._lib.wbc_config_rows(C.byref(cfg), self.n_velocity_dimensions, C.byref(m), C.byref(nc)))
        return m.value, nc.value

    def _as_batch(self, x, width):
        """numpy / list / tensor -> contiguous float64 CUDA tensor [N, width] (broadcast from one row)."""
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x, dtype=np.float64))
        x = x.to(device=self.device, dtype=torch.float64).reshape(-1, width)
        if x.shape[0] == 1 and self.N != 1:
            x = x.expand(self.N, width)
        if x.shape[0] != self.N:
            raise ValueError(f"expected batch {self.N}, got {x.shape[0]}")
        return x.contiguous()

    def _pack_targets(self, target_cartesian_pos_EE, target_cartesian_pos_trunk):
        """runWBC arguments (list of 5 EE targets + trunk target, :1330) -> [N, 18]."""
        T = self._targets
        if torch.is_tensor(target_cartesian_pos_EE) and target_cartesian_pos_EE.dim() == 3:
            T[:, :15] = target_cartesian_pos_EE.to(self.device, torch.float64).reshape(self.N, 15)
        else:
            for i in range(5):
                T[:, 3 * i:3 * i + 3] = self._as_batch(target_cartesian_pos_EE[i], 3)
        if target_cartesian_pos_trunk is not None:
            T[:, 15:18] = self._as_batch(target_cartesian_pos_trunk, 3)
        return T

    def _io(self, targets=None, **out):
        io = cabi.WbcStepIO()
        io.q = self.current_joint_config.data_ptr()
        io.targets = (targets if targets is not None else self._targets).data_ptr()
        io.mem_in = self._mem.data_ptr()
        io.ref = self._ref.data_ptr()
        io.dt = float(self.dt)
        for k, v in out.items():
            setattr(io, k, v.data_ptr() if v is not None else None)
        return io

    # ---------------------------------------------------------------------------------- named views of the packed state
    @property
    def prev_EE_pos(self):
        return self._mem[:, 0:15].view(self.N, 5, 3)

    @property
    def prev_EE_CoM_rot(self):
        return self._mem[:, 15:60].view(self.N, 5, 3, 3)

    @property
    def prev_trunk_ref(self):
        return self._mem[:, 60:63]

    @property
    def old_ref_trunk_rot_matrix(self):
        return self._mem[:, 63:72].view(self.N, 3, 3)

    @property
    def default_EE_ori_list(self):
        return self._ref[:, 0:15].view(self.N, 5, 3)

    @property
    def default_trunk_ori(self):
        return self._ref[:, 15:18]

    @property
    def initial_trunk_pos(self):
        return self._ref[:, 18:21]

    @property
    def initial_trunk_ori_euler(self):
        return self._ref[:, 21:24]

    @property
    def EE_frame_pos(self):
        return [self._oMf[:, i, 9:12] for i in range(5)]

    @property
    def trunk_frame_pos(self):
        return self._oMf[:, cabi.FRAME_TRUNK, 9:12]

    # ---------------------------------------------------------------------------------- switches :176-193, :1415-1464
    def setTasks(self, Trunk=False, FR=False, FL=False, RR=False, RL=False, Grip=False, Joint=False):
        self.task_active_Trunk = Trunk
        self.task_active_FR_foot = FR
        self.task_active_FL_foot = FL
        self.task_active_RR_foot = RR
        self.task_active_RL_foot = RL
        self.task_active_GRIP = Grip
        self.task_active_Joint = Joint

    def setConstraints(self, CoM=False, Trunk=False, FR=False, FL=False, RR=False, RL=False, Grip=False):
        self.const_active_CoM = CoM
        self.const_active_Trunk = Trunk
        self.const_active_FR_foot = FR
        self.const_active_FL_foot = FL
        self.const_active_RR_foot = RR
        self.const_active_RL_foot = RL
        self.const_active_GRIP = Grip

    def staticReachMode(self):
        self.trunk_weight = np.identity(6) * 1
        self.EE_weight = [np.identity(6) * 1 for _ in range(5)]
        self.cart_task_weight_EE_list = [100, 100, 100, 100, 1]
        self.cart_task_weight_Trunk = 1
        self.joint_task_weight = 0.001
        self.trunk_gain = np.identity(6) * 0.8
        self.FL_gain = np.identity(6) * 0.8
        self.FR_gain = np.identity(6) * 0.8
        self.RL_gain = np.identity(6) * 0.8
        self.RR_gain = np.identity(6) * 0.8
        self.GRIP_gain = np.identity(6) * 0.05
        self.EE_gains = [self.FL_gain, self.FR_gain, self.RL_gain, self.RR_gain, self.GRIP_gain]

    # ---------------------------------------------------------------------------------- FK / Jacobian refresh :387-428
    def _refresh(self, config):
        self.previous_joint_config = self.current_joint_config
        self.current_joint_config = config
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_fk_jac(self._model, _ptr(config), self.N, self._slot_sel, len(self._slot_frames),
                                            cabi.RF_WORLD, _ptr(self._oMf), None, _stream_ptr()))
            cabi.check(self._lib.wbc_joint_jacobians(self._model, _ptr(config), self.N, _ptr(self._oMi), _ptr(self.J),
                                                     _stream_ptr()))

    def updateState(self, joint_config, imu_data=0, feedback=True, running=False):
        nq = self.n_configuration_dimensions
        if feedback is True and running is True:
            joints = self._as_batch(joint_config, nq - 7)
            imu = self._as_batch(imu_data, 4)
            config = torch.cat((self.current_joint_config[:, :3], imu, joints), dim=1).contiguous()
        else:
            config = self._as_batch(joint_config, nq).clone()
        self._refresh(config)
        if running is True:
            config = self._base_estimate(config, want_config=True)                       # trunkWorldPos (:414)
            self._refresh(config)                                                        # :418-428

    def getFrameJacobian(self, frame_id, reference_frame):
        """pin.getFrameJacobian(model, data, frame_id, rf) -> [N, 6, nv] (fresh tensor)."""
        slot = self._frame_slot[int(frame_id)]
        out = torch.empty(self.N, 1, 6, self.n_velocity_dimensions, dtype=torch.float64, device=self.device)
        sel = (C.c_int32 * 1)(slot)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_fk_jac(self._model, _ptr(self.current_joint_config), self.N, sel, 1,
                                            int(reference_frame), None, _ptr(out), _stream_ptr()))
        return out[:, 0]

    def frameJacobians(self, reference_frame, slots=None):
        """All hot-frame Jacobians at once: ([N, F, 12] placements, [N, F, 6, nv])."""
        slots = list(range(6)) if slots is None else list(slots)
        sel = (C.c_int32 * len(slots))(*slots)
        oMf = torch.empty(self.N, len(slots), 12, dtype=torch.float64, device=self.device)
        J = torch.empty(self.N, len(slots), 6, self.n_velocity_dimensions, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_fk_jac(self._model, _ptr(self.current_joint_config), self.N, sel, len(slots),
                                            int(reference_frame), _ptr(oMf), _ptr(J), _stream_ptr()))
        return oMf, J

    # ---------------------------------------------------------------------------------- bootstrap :196-351
    def setInitialState(self, bootstrap_steps=None):
        """The reference's constructor bootstrap (Robot_Wrapper4.py:196-351), batched: from ``pin.neutral`` every robot
        follows linear end-effector trajectories into the crouched start pose with 2000 bounds-only QP ticks (task
        stack P1, cold QP each tick, plain ``integrate``), one fused launch per tick for the whole batch."""
        t, N, nq, nv = self.robot_model, self.N, self.n_configuration_dimensions, self.n_velocity_dimensions
        f64 = dict(dtype=torch.float64, device=self.device)
        q = torch.zeros(N, nq, **f64)
        q[:, 6] = 1.0                                                                    # pin.neutral (:199)
        upper = torch.as_tensor(np.asarray(t.upper, dtype=np.float64)[:nq], **f64)
        q[:, :nv] = torch.minimum(q[:, :nv], upper[:nv])                                 # :201-208 (quirk D.12: range(nv), upper side only)
        self.updateState(q, feedback=False)
        mem, ref = self._log_previous_states()                                           # :214-219
        self._mem.copy_(mem)
        self._ref[:, :18] = ref[:, :18]                                                  # default orientations :222-226
        trunk_target = self.trunk_frame_pos.clone()                                      # :229
        ee0 = [self.EE_frame_pos[i].clone() for i in range(5)]
        scale_F = torch.tensor([1.0, 1.0, 0.9], **f64)
        scale_G = torch.tensor([1.1, 1.0, 1.5], **f64)
        second = []
        for i in range(4):                                                               # :247-250
            p2 = ee0[i].clone()
            p2[:, 0] = self.robot_data.oMf[self.hip_waist_joint_index_list_frame[i]].translation[:, 0]
            second.append(p2 * scale_F)
        g2 = ee0[4].clone()
        g2[:, 2] = self.robot_data.oMi[self.arm_base_id].translation[:, 2]                # :253
        g2[:, 0] = self.robot_data.oMi[self.FR_hip_joint].translation[:, 0]               # :254
        second.append(g2 * scale_G)
        traj = [LinearTrajectory(torch.stack((ee0[i], second[i]))) for i in range(5)]    # order FR, FL, RR, RL, G (:269)
        self.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)   # :272
        ticks = np.arange(0, 2, 0.001).tolist()                                          # :275
        if bootstrap_steps is not None:
            ticks = ticks[:bootstrap_steps]
        was_initialised, self.initialised = self.initialised, False
        for tt in ticks:                                                                 # :278-325
            ee = torch.stack([traj[i].eval(tt) for i in range(5)], dim=1)
            self.step(ee, trunk_target, advance=True, plain_integrate=True, constraint_mask=0)
        q = self.current_joint_config.clone()
        q[:, 3:6] = 0.0                                                                  # :328-330
        self.updateState(q, feedback=False)
        feet_z = sum(self.EE_frame_pos[i][:, 2] for i in range(4))
        q = self.current_joint_config.clone()
        q[:, 2] = -feet_z / 4 + self.foot_radius                                         # :336-338
        self.updateState(q, feedback=False)
        jc = self.current_joint_config[:, 7:]
        self.FL_leg, self.FR_leg, self.RL_leg, self.RR_leg, self.grip = jc[:, 0:3], jc[:, 3:6], jc[:, 6:9], jc[:, 9:12], jc[:, 12:]
        self.dt = self.step_time
        self.initialised = was_initialised
        return self.current_joint_config

    # ---------------------------------------------------------------------------------- :354-383
    def _log_previous_states(self):
        mem = torch.empty_like(self._mem)
        ref = torch.empty_like(self._ref)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_init_memory(self._model, _ptr(self.current_joint_config), self.N, _ptr(mem),
                                                 _ptr(ref), _stream_ptr()))
        return mem, ref

    def initialiseWBC(self, imu_data):
        self.updateState(self.current_joint_config, imu_data, running=False)
        mem, ref = self._log_previous_states()
        self._mem.copy_(mem)
        self._ref.copy_(ref)

    # ---------------------------------------------------------------------------------- :440-449
    def jointVelocitiestoConfig(self, joint_vel, update_model=False):
        v = self._as_batch(joint_vel, self.n_velocity_dimensions)
        new_config = torch.empty_like(self.current_joint_config)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_integrate(self._model, _ptr(self.current_joint_config), _ptr(v), self.N,
                                               float(self.dt), _ptr(new_config), _stream_ptr()))
        if update_model is True:
            self.updateState(new_config, feedback=False, running=bool(self.initialised))
            return None
        return new_config

    # ---------------------------------------------------------------------------------- assembly accessors
    def _assemble(self, cfg, want, targets=None, mem_out=None):
        m, nc = self._rows(cfg)
        N, nv = self.N, self.n_velocity_dimensions
        f64 = dict(dtype=torch.float64, device=self.device)
        shapes = {"A": (N, m, nv), "b": (N, m), "lb": (N, nv), "ub": (N, nv), "C": (N, nc, nv), "Clb": (N, nc),
                  "Cub": (N, nc), "H": (N, nv, nv), "g": (N, nv)}
        out = cabi.WbcAssembleOut()
        res = {}
        for k in want:
            res[k] = torch.zeros(*shapes[k], **f64)
            setattr(out, k, res[k].data_ptr())
        io = self._io(targets=targets, mem_out=mem_out)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_assemble(self._model, C.byref(cfg), C.byref(io), N, C.byref(out), _stream_ptr()))
        return res

    def endEffectorA2(self, frame_index):                                                # :474-484
        cfg = self._config(task_mask=cabi.TASK_FR << frame_index, constraint_mask=0)
        self.EE_A_list[frame_index] = self._assemble(cfg, ["A"])["A"]

    def trunkA(self):                                                                    # :487-490
        cfg = self._config(task_mask=cabi.TASK_TRUNK, constraint_mask=0)
        self.trunk_A = self._assemble(cfg, ["A"])["A"]

    def qpA(self):                                                                       # :1271-1280
        cfg = self._config(constraint_mask=0)
        return self._assemble(cfg, ["A"])["A"]

    def qpb(self, target_cartesian_pos_EE, target_cartesian_pos_trunk):                  # :1283-1294 (mutates task memory)
        T = self._pack_targets(target_cartesian_pos_EE, target_cartesian_pos_trunk)
        cfg = self._config(constraint_mask=0)
        return self._assemble(cfg, ["b"], targets=T, mem_out=self._mem)["b"].unsqueeze(-1)

    def velDamperJointConstraints(self):                                                 # :572-637
        cfg = self._config(constraint_mask=0, task_mask=self._task_mask() or cabi.TASK_JOINT)
        r = self._assemble(cfg, ["lb", "ub"])
        return r["lb"], r["ub"]

    def _constraint(self, mask):
        cfg = self._config(task_mask=cabi.TASK_JOINT, constraint_mask=mask)
        r = self._assemble(cfg, ["C", "Clb", "Cub"])
        return r["C"], r["Clb"], r["Cub"]

    def EEConstraint(self, frame_index):                                                 # :757-761
        return self._constraint(cabi.CON_FR << frame_index)

    def trunkConstraint(self):                                                           # :707-754
        return self._constraint(cabi.CON_TRUNK)

    def CoMConstraint(self):                                                             # :669-694
        return self._constraint(cabi.CON_COM)

    def findConstraints(self):                                                           # :764-836 (returns C.T)
        C_, Clb, Cub = self._constraint(self._constraint_mask())
        return C_.transpose(1, 2), Clb, Cub

    def assemble(self, target_cartesian_pos_EE=None, target_cartesian_pos_trunk=None, want=("A", "b", "lb", "ub", "C", "Clb", "Cub", "H", "g"),
                 update_memory=False):
        """Everything runWBC hands to the QP, in one launch (debug / parity accessor)."""
        T = self._pack_targets(target_cartesian_pos_EE, target_cartesian_pos_trunk) if target_cartesian_pos_EE is not None else None
        return self._assemble(self._config(), list(want), targets=T, mem_out=self._mem if update_memory else None)

    # ---------------------------------------------------------------------------------- :1297-1327
    def _base_estimate(self, config, want_config=False):
        """``wbc_base_estimate``: FK at ``config`` and the base position estimate from the four foot targets."""
        out = torch.empty_like(config) if want_config else torch.empty(self.N, 3, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_base_estimate(self._model, _ptr(config), None, _ptr(self._targets), self.N,
                                                   _ptr(out) if want_config else None,
                                                   None if want_config else _ptr(out), _stream_ptr()))
        return out

    def trunkWorldPos(self):
        """[N, 3]: mean(foot targets) - R_trunk mean(foot - trunk) at the current configuration; the foot targets are the
        ones of the last ``runWBC`` / ``qpb`` / ``step`` call (``FR_target_cartesian_pos`` ..., :1332-1335)."""
        return self._base_estimate(self.current_joint_config.contiguous())

    # ---------------------------------------------------------------------------------- the fused tick :1330-1412
    def step(self, target_cartesian_pos_EE, target_cartesian_pos_trunk, imu_quat=None, advance=True,
             plain_integrate=False, constraint_mask=None, report_active_set=True):
        """One fused launch: FK + Jacobians + task stack + bounds + constraints + QP (+ integrate / base estimate).

        Returns qdot [N, nv]; ``last_status`` / ``last_iters`` / ``last_active_set`` hold the per-state QP report.
        With ``advance`` the task memory and ``current_joint_config`` move on in place, as runWBC does.
        ``report_active_set=False``: only what the reference's ``solveQP`` returns (the primal solution) plus status and
        iteration count; the active-set bit masks (an extension, 16 B per state) are not packed and not written.
        """
        T = self._pack_targets(target_cartesian_pos_EE, target_cartesian_pos_trunk)
        cfg = self._config(constraint_mask=constraint_mask)
        imu = self._as_batch(imu_quat, 4) if imu_quat is not None else None
        q_next = torch.empty_like(self.current_joint_config) if advance else None
        io = self._io(targets=T, qdot=self.qdot, status=self.last_status, iters=self.last_iters,
                      active_set=self.last_active_set if report_active_set else None,
                      mem_out=self._mem if advance else None, q_next=q_next, imu_quat=imu)
        io.flags = cabi.STEP_FLAG_PLAIN_INTEGRATE if plain_integrate else 0
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_step(self._model, C.byref(cfg), C.byref(io), self.N, _stream_ptr()))
        if advance:
            self.previous_joint_config = self.current_joint_config
            self.current_joint_config = q_next
        return self.qdot

    def step_host(self, host_in, host_out, chunks=0, resident_state=False, closed_loop=False, delta_inputs=False, cfg=None):
        """One tick with HOST buffers (what a caller holding NumPy arrays pays end to end): one C-ABI call,
        ``wbc_step_host``.

        ``closed_loop=True`` -- the tick ``runWBC`` actually is (Robot_Wrapper4.py:1330-1412): ``host_in`` carries this
        tick's arguments, ``targets`` [N, 18] (5 EE targets + trunk target) and optionally ``imu`` [N, 4] (``base_config``);
        the configuration, the task memory and the per-robot references are the controller's state and stay on the
        device, advanced in place exactly as runWBC mutates its object (``prev_EE_pos`` ... :995-996, :1151-1152,
        ``current_joint_config`` :1397-1402).  ``host_out`` receives ``joint_targets`` [N, nq - 7] (the five slices runWBC
        returns, :1405-1412), ``status`` / ``iters`` [N] (int32) and, if present, ``qdot`` [N, nv].  K calls equal
        ``rollout`` over the same K ticks.

        ``closed_loop=False`` -- the open-loop tick: ``host_in`` = q [N, nq], targets [N, 18] (+ mem [N, 72], ref [N, 24]
        unless ``resident_state``), ``host_out`` = qdot, status, iters; nothing on the device is advanced.

        Float arrays are float64, or ALL float32 (the optional FP32 I/O mode: half the PCIe bytes, float64 arithmetic
        inside).  ``delta_inputs`` (float32, closed loop): ``targets`` / ``imu`` hold INCREMENTS over the previous tick's
        targets / the resident base quaternion (``HostDeltaEncoder`` produces them): a float32 increment is exact to
        ~1e-10, so the mode agrees with the float64 call to < 1e-4 in qdot, where absolute float32 positions (error
        ~3e-8, times 1 / dt = 500 in the target laws) only guarantee it for the joint position targets.  ``chunks=-1``: with page-locked tensors the kernel reads the
        inputs from and writes the outputs to host memory directly (zero-copy, one launch); with pageable tensors, or
        ``chunks >= 1``, the batch is cut into slices that go host -> device, through the fused kernel and back on three
        streams owned by the model; the current stream waits for all of them.  ``chunks=0`` (default): self-tuning -- the
        first four calls of a problem shape try both (zero-copy twice, 8 slices twice), the faster one runs from then on
        (``host_path()`` tells which); the results do not depend on the path.  ``cfg``: a ``WbcConfig`` built earlier with
        ``_config()`` (settings that do not change from tick to tick need not be marshalled again).
        Returns (h2d_bytes, d2h_bytes).
        """
        N, nq, nv = self.N, self.n_configuration_dimensions, self.n_velocity_dimensions
        if closed_loop:
            moved = ("targets",) + (("imu",) if host_in.get("imu") is not None else ())
            outs = tuple(k for k in ("joint_targets", "qdot") if host_out.get(k) is not None)
        else:
            moved = ("q", "targets") if resident_state else ("q", "targets", "mem", "ref")
            outs = ("qdot",)
        fdt = host_in["targets"].dtype
        if fdt not in (torch.float64, torch.float32):
            raise ValueError("host arrays must be float64 or float32")
        for k in moved:
            t = host_in[k]
            if t.device.type != "cpu" or t.dtype != fdt or not t.is_contiguous() or t.shape[0] != N:
                raise ValueError(f"host_in[{k!r}] must be a contiguous {fdt} CPU tensor with {N} rows")
        reports = tuple(k for k in ("status", "iters") if host_out.get(k) is not None)
        if "status" not in reports:
            raise ValueError("host_out['status'] is required")
        for k in outs + reports:
            t, dt_ = host_out[k], (fdt if k in outs else torch.int32)
            if t.device.type != "cpu" or t.dtype != dt_ or not t.is_contiguous() or t.shape[0] != N:
                raise ValueError(f"host_out[{k!r}] must be a contiguous {dt_} CPU tensor with {N} rows")
        io = cabi.WbcStepIO()
        self.current_joint_config = self.current_joint_config.contiguous()
        io.q = self.current_joint_config.data_ptr()
        io.targets = self._targets.data_ptr()
        io.mem_in = self._mem.data_ptr()
        io.ref = self._ref.data_ptr()
        io.dt = float(self.dt)
        io.qdot = self.qdot.data_ptr()
        io.status = self.last_status.data_ptr()
        io.iters = self.last_iters.data_ptr()
        host = cabi.WbcHostIO()
        host.dtype = cabi.HOST_F32 if fdt == torch.float32 else cabi.HOST_F64
        host.flags = cabi.HOST_FLAG_DELTA_INPUTS if delta_inputs else 0
        host.targets = host_in["targets"].data_ptr()
        if closed_loop:
            if getattr(self, "_stage_imu", None) is None:                               # staging space of the sliced path
                self._stage_imu = torch.empty(N, 4, dtype=torch.float64, device=self.device)
                self._stage_joints = torch.empty(N, nq - 7, dtype=torch.float64, device=self.device)
            io.q_next = io.q
            io.mem_out = io.mem_in
            io.joint_targets = self._stage_joints.data_ptr()
            if "imu" in moved:
                io.imu_quat = self._stage_imu.data_ptr()
                host.imu_quat = host_in["imu"].data_ptr()
            if "joint_targets" in outs:
                host.joint_targets = host_out["joint_targets"].data_ptr()
            if "qdot" in outs:
                host.qdot = host_out["qdot"].data_ptr()
        else:
            host.q = host_in["q"].data_ptr()
            if not resident_state:
                host.mem_in = host_in["mem"].data_ptr()
                host.ref = host_in["ref"].data_ptr()
            host.qdot = host_out["qdot"].data_ptr()
        host.status = host_out["status"].data_ptr()
        if "iters" in reports:
            host.iters = host_out["iters"].data_ptr()
        with torch.cuda.device(self.device):
            cabi.check(self._lib.wbc_step_host(self._model, C.byref(cfg if cfg is not None else self._config()), C.byref(io),
                                               C.byref(host), N, int(chunks), _stream_ptr()))
        if closed_loop:
            self.firstQP = False
        esz = 4 if fdt == torch.float32 else 8
        h2d = sum(host_in[k].numel() * esz for k in moved)
        d2h = sum(host_out[k].numel() * esz for k in outs) + sum(host_out[k].numel() * 4 for k in reports)
        return h2d, d2h

    def host_path(self):
        """What the self-tuning ``step_host(chunks=0)`` settled on: "undecided", "zero_copy" or "staged_8"."""
        c = int(self._lib.wbc_step_host_path(self._model))
        return {0: "zero_copy", -1: "undecided"}.get(c, f"staged_{c}")

    def runWBC(self, base_config, target_cartesian_pos_EE=None, target_cartesian_pos_trunk=None):
        """Robot_Wrapper4.py:1330-1412, batched.  Like the reference, the QP's report is not acted upon: a state whose QP
        hit the iteration cap or was infeasible (``last_status`` != 0) is integrated with whatever the solver held, as
        qpOASES' primal vector is used unchecked (QP_Wrapper.py:45-51).  ``check_status()`` raises on such states."""
        self.step(target_cartesian_pos_EE, target_cartesian_pos_trunk, imu_quat=base_config, advance=True)
        self.firstQP = False
        self._refresh(self.current_joint_config)           # keep the accessor caches (oMf, J) on the new state
        joint_config = self.current_joint_config[:, 7:]
        FL_leg = joint_config[:, 0:3]                                                     # :1405-1409
        FR_leg = joint_config[:, 3:6]
        RL_leg = joint_config[:, 6:9]
        RR_leg = joint_config[:, 9:12]
        grip = joint_config[:, 12:]
        return FL_leg, FR_leg, RL_leg, RR_leg, grip

    def rollout(self, target_EE_traj, target_trunk_traj, imu_quat_traj=None, record=False, report_active_set=True):
        """Closed-loop horizon (BASELINE config 5): K consecutive runWBC ticks (:1330-1412) for all N robots; task memory and
        configuration stay resident on the device between ticks.  One persistent launch for the whole horizon where the
        reduced-front kernel applies (wbc_rollout), else one fused launch per tick.

        ``target_EE_traj`` [K, N, 5, 3], ``target_trunk_traj`` [K, N, 3], optional ``imu_quat_traj`` [K, N, 4] (base
        orientation fed back each tick; default: the integrated orientation).  Returns qdot of the last tick, or with
        ``record`` the tuple (q history [K, N, nq], qdot history [K, N, nv], status history [K, N]).
        ``report_active_set=False``: the active-set bit masks of the last tick are not packed (as in ``step``).
        """
        K = int(target_EE_traj.shape[0])
        if not record:
            # one C-ABI call: q and task memory advanced in place on the device
            traj = torch.cat((target_EE_traj.to(self.device, torch.float64).reshape(K, self.N, 15),
                              target_trunk_traj.to(self.device, torch.float64).reshape(K, self.N, 3)), dim=2).contiguous()
            imu = imu_quat_traj.to(self.device, torch.float64).reshape(K, self.N, 4).contiguous() if imu_quat_traj is not None else None
            self.current_joint_config = self.current_joint_config.contiguous().clone()
            io = self._io(targets=traj, qdot=self.qdot, status=self.last_status, iters=self.last_iters,
                          active_set=self.last_active_set if report_active_set else None)
            with torch.cuda.device(self.device):
                cabi.check(self._lib.wbc_rollout(self._model, C.byref(self._config()), C.byref(io), _ptr(traj), _ptr(imu), K,
                                                 self.N, _stream_ptr()))
            self._targets.copy_(traj[-1])
            self.firstQP = False
            return self.qdot
        qs, vs, st = [], [], []
        for k in range(K):
            imu = imu_quat_traj[k] if imu_quat_traj is not None else None
            self.step(target_EE_traj[k], target_trunk_traj[k], imu_quat=imu, advance=True)
            qs.append(self.current_joint_config.clone()); vs.append(self.qdot.clone()); st.append(self.last_status.clone())
        self.firstQP = False
        return torch.stack(qs), torch.stack(vs), torch.stack(st)

    def check_status(self):
        """Raise if any state's last QP did not end with WBC_QP_SOLVED (the reference silently ignores qpOASES' return
        value; callers that integrate ``runWBC`` / ``rollout`` results blindly may want this).  Synchronises."""
        bad = int((self.last_status != 0).sum().item())
        if bad:
            raise cabi.WbcError(f"{bad} of {self.N} states: QP not solved (status bits: 1 iteration cap, 2 infeasible, 4 not PD)")

    def launch_info(self):
        g, b, s, r = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        cabi.check(self._lib.wbc_step_launch_info(self._model, C.byref(g), C.byref(b), C.byref(s), C.byref(r)))
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "regs": r.value}
