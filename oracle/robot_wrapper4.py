"""oracle.robot_wrapper4 -- CPU restatement of ``wrappers/Robot_Wrapper4.py`` (class ``RobotModel``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  One robot state per object, NumPy + SciPy
``Rotation`` (called verbatim, like the reference) on top of ``oracle.pin`` / ``oracle.qp_wrapper``.
Bug-compatible with the reference: the velocity-damper off-by-one, the
untransposed trunk orientation law, the discarded EE quaternion error, ``C.T`` from
``findConstraints`` ... are all reproduced on purpose.  Two deliberate departures, both forced:

  * ``dt`` is an explicit attribute (reference: wall-clock busy-wait, Robot_Wrapper4.py:292-293,
    1338-1342, nominal 0.002 s) -- a wall-clock value cannot be compared against anything;
  * nothing is printed (reference prints from inside the hot path, :1075-1085).

Every method cites the reference lines it follows.
"""
import math

import numpy as np
from scipy.spatial.transform import Rotation as R

from . import pin
from .qp_wrapper import QP

EE_FRAME_NAMES = ["FR_foot_fixed", "FL_foot_fixed", "RR_foot_fixed", "RL_foot_fixed", "gripper_bar"]
EE_JOINT_NAMES = ["FR_calf_joint", "FL_calf_joint", "RR_calf_joint", "RL_calf_joint", "gripper"]
HIP_WAIST_JOINT_NAMES = ["FR_hip_joint", "FL_hip_joint", "RR_hip_joint", "RL_hip_joint", "waist"]


class _LinearTrajectory:
    """klampt.model.trajectory.Trajectory(milestones=...).eval(t): knot i at t = i, clamped ends."""

    def __init__(self, milestones):
        self.m = [np.asarray(x, dtype=float) for x in milestones]

    def eval(self, t):
        if t <= 0:
            return self.m[0].copy()
        if t >= len(self.m) - 1:
            return self.m[-1].copy()
        i = int(math.floor(t))
        u = t - i
        return self.m[i] + u * (self.m[i + 1] - self.m[i])


class RobotModel:
    def __init__(self, urdf_path, mesh_dir_path=None, EE_frame_names=EE_FRAME_NAMES, EE_joint_names=EE_JOINT_NAMES,
                 G_base="waist", imu="imu_joint", FR_hip_joint="FR_hip_joint",
                 hip_waist_joint_names=HIP_WAIST_JOINT_NAMES, foot_offset=False,
                 dt=0.002, run_bootstrap=False, bootstrap_steps=None):
        """Robot_Wrapper4.py:19-173.  ``urdf_path`` may also be an ``oracle.pin.Model`` or a tree-table dict.

        ``run_bootstrap=False`` skips the 2000-tick ``setInitialState`` loop the reference runs in its
        constructor (:161); call ``setInitialState()`` explicitly to get it.
        """
        if isinstance(urdf_path, pin.Model):
            self.robot_model = urdf_path
        elif isinstance(urdf_path, dict):
            self.robot_model = pin.Model.from_dict(urdf_path)
        elif str(urdf_path).endswith(".json"):
            self.robot_model = pin.Model.from_json(urdf_path)
        else:
            self.robot_model = pin.buildModelFromUrdf(urdf_path)                       # :21
        self.robot_data = self.robot_model.createData()                                # :23
        self.joint_names = self.robot_model.names
        self.foot_radius = 0

        self.trunk_frame_index = self.robot_model.getFrameId(imu, pin.FIXED_JOINT)      # :30
        self.current_joint_config = 0
        self.EE_frame_names = EE_frame_names
        self.EE_joint_names = EE_joint_names
        self.hip_waist_joint_names = hip_waist_joint_names
        self.arm_base_id = self.robot_model.getJointId(G_base)                         # :37
        self.arm_base_frame_id = self.robot_model.getFrameId(G_base, pin.JOINT)
        self.FR_hip_joint = self.robot_model.getJointId(FR_hip_joint)
        self.n_velocity_dimensions = self.robot_model.nv
        self.n_configuration_dimensions = self.robot_model.nq
        self.n_of_EE = 5
        self.end_effector_index_list_frame = []
        self.end_effector_index_list_joint = []
        self.hip_waist_joint_index_list_frame = []
        for i in range(len(self.EE_joint_names)):                                      # :46-52
            self.end_effector_index_list_frame.append(self.robot_model.getFrameId(self.EE_frame_names[i], pin.FIXED_JOINT))
            self.end_effector_index_list_joint.append(self.robot_model.getJointId(self.EE_joint_names[i]))
            self.hip_waist_joint_index_list_frame.append(self.robot_model.getFrameId(self.hip_waist_joint_names[i], pin.JOINT))

        if foot_offset is True:                                                        # :55-58
            self.foot_radius = self.robot_model.collision_geoms[8][2]

        self.initialised = False
        self.EE_frame_pos = [0, 0, 0, 0, 0]
        self.default_trunk_ori = np.array([[0, 0, 0]]).T
        self.default_EE_ori_list = [np.array([[0, 0, 0]]).T for _ in range(5)]
        q = pin.neutral(self.robot_model)                                              # :66
        self.updateState(q, feedback=False)
        arm_base_placement = np.copy(self.robot_data.oMf[self.arm_base_frame_id].translation)
        gripper_placement = np.copy(self.robot_data.oMf[self.end_effector_index_list_frame[4]].translation)
        self.arm_reach = np.sum(gripper_placement - arm_base_placement)                # :70

        # weights :72-93
        self.trunk_weight = np.identity(6) * 1
        self.FR_weight = np.identity(6) * 1
        self.FL_weight = np.identity(6) * 1
        self.RR_weight = np.identity(6) * 1
        self.RL_weight = np.identity(6) * 1
        self.grip_weight = np.identity(6) * 1
        self.EE_weight = [self.FR_weight, self.FL_weight, self.RR_weight, self.RL_weight, self.grip_weight]
        self.cart_task_weight_FR = 1
        self.cart_task_weight_FL = 1
        self.cart_task_weight_RR = 1
        self.cart_task_weight_RL = 1
        self.cart_task_weight_GRIP = 1
        self.cart_task_weight_Trunk = 1
        self.cart_task_weight_EE_list = [self.cart_task_weight_FR, self.cart_task_weight_FL, self.cart_task_weight_RR,
                                         self.cart_task_weight_RL, self.cart_task_weight_GRIP]
        self.joint_task_weight = 0.05

        # task / constraint switches :96-111
        self.setTasks()
        self.setConstraints()

        # timing :114-116 (dt explicit here)
        self.previous_time = 0
        self.step_time = dt
        self.dt = dt

        # gains :119-125 (list order FL, FR, RL, RR, GRIP -- Appendix D.8)
        self.trunk_gain = np.identity(6) * 0.5
        self.FL_gain = np.identity(6) * 0.5
        self.FR_gain = np.identity(6) * 0.5
        self.RL_gain = np.identity(6) * 0.5
        self.RR_gain = np.identity(6) * 0.5
        self.GRIP_gain = np.identity(6) * 0.5
        self.EE_gains = [self.FL_gain, self.FR_gain, self.RL_gain, self.RR_gain, self.GRIP_gain]

        # memory :129-158
        self.prev_trunk_ref = np.array([0, 0, 0])
        self.old_ref_trunk_rot_matrix = np.zeros((3, 3))
        self.prev_EE_pos = [0, 0, 0, 0, 0]
        self.prev_EE_CoM_rot = [0, 0, 0, 0, 0]
        self.trunk_frame_pos = np.copy(self.robot_data.oMf[self.trunk_frame_index].translation)
        self.EE_A_list = [0, 0, 0, 0, 0]
        self.EE_b_list = [0, 0, 0, 0, 0]
        self.firstQP = True
        self.qp = None
        self.trunk_A = 0
        self.trunk_B = 0
        self.last = {}
        self.bootstrap_steps = bootstrap_steps

        if run_bootstrap:
            self.setInitialState()                                                     # :161
        self.initialised = True
        self._log_previous_states()                                                    # :168-173

    def _log_previous_states(self):
        """Robot_Wrapper4.py:168-173 / :214-219 / :370-376."""
        d = self.robot_data
        self.prev_trunk_ref = np.copy(d.oMf[self.trunk_frame_index].translation)
        for i in range(len(self.prev_EE_pos)):
            self.prev_EE_pos[i] = np.copy(d.oMf[self.end_effector_index_list_frame[i]].translation)
            EE_rot = np.copy(d.oMf[self.end_effector_index_list_frame[i]].rotation)
            hip_waist_rot = np.copy(d.oMf[self.trunk_frame_index].rotation)
            self.prev_EE_CoM_rot[i] = np.dot(hip_waist_rot.T, EE_rot)

    # ------------------------------------------------------------------ switches :176-193
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

    # ------------------------------------------------------------------ bootstrap :196-351
    def _set_default_orientations(self):
        """:222-226 / :363-367."""
        d = self.robot_data
        self.default_trunk_ori = R.from_matrix(d.oMf[self.trunk_frame_index].rotation).as_euler('xyz').reshape(3, 1)
        for i in range(len(self.default_EE_ori_list)):
            self.default_EE_ori_list[i] = R.from_matrix(
                d.oMf[self.end_effector_index_list_frame[i]].rotation).as_euler('xyz').reshape(3, 1)

    def setInitialState(self, record=None):
        """Robot_Wrapper4.py:196-351: 2000 bounds-only QP ticks along linear EE trajectories (pattern P1).

        ``record`` (optional list) receives a dict per tick with the QP inputs/outputs.
        """
        model, data = self.robot_model, self.robot_data
        q = pin.neutral(model)
        for i in range(self.n_velocity_dimensions):                                    # :201-208 (quirk D.12)
            if q[i] > model.upperPositionLimit[i]:
                q[i] = model.upperPositionLimit[i]
        self.updateState(q, feedback=False)
        self._log_previous_states()                                                    # :214-219
        self._set_default_orientations()                                               # :222-226

        Trunk_target_pos = self.trunk_frame_pos.T                                      # :229
        EE_target_pos = [self.EE_frame_pos[i].T for i in range(5)]
        multiplier_F = np.identity(3)
        multiplier_R = np.identity(3)
        multiplier_G = np.identity(3)
        multiplier_F[2, 2] = 0.9
        multiplier_R[2, 2] = 0.9
        multiplier_G[2, 2] = 1.5
        multiplier_G[0, 0] = 1.1
        pos2 = [np.copy(EE_target_pos[i]) for i in range(4)]
        for i in range(4):                                                             # :247-250
            pos2[i][0] = data.oMf[self.hip_waist_joint_index_list_frame[i]].translation[0]
        EE_G_pos_2 = EE_target_pos[4].reshape((3,)).tolist()
        EE_G_pos_2[2] = data.oMi[self.arm_base_id].translation[2]                      # :253
        EE_G_pos_2[0] = data.oMi[self.FR_hip_joint].translation[0]                     # :254
        mult = [multiplier_F, multiplier_F, multiplier_R, multiplier_R]
        milestones = [[EE_target_pos[i].reshape((3,)).tolist(), np.dot(pos2[i].reshape((3,)), mult[i]).tolist()]
                      for i in range(4)]
        milestones.append([EE_target_pos[4].reshape((3,)).tolist(), np.dot(EE_G_pos_2, multiplier_G).tolist()])
        EE_traj = [_LinearTrajectory(m) for m in milestones]                           # order FR, FL, RR, RL, G :269

        self.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)  # :272
        trajectory_interval = np.arange(0, len(milestones[0]), 0.001).tolist()         # :275
        if self.bootstrap_steps is not None:
            trajectory_interval = trajectory_interval[:self.bootstrap_steps]

        for t in trajectory_interval:                                                  # :278-325
            for ii in range(len(EE_traj)):
                EE_target_pos[ii] = np.array(EE_traj[ii].eval(t)).reshape(3, 1)
            self.FR_target_cartesian_pos = EE_target_pos[0]
            self.FL_target_cartesian_pos = EE_target_pos[1]
            self.RR_target_cartesian_pos = EE_target_pos[2]
            self.RL_target_cartesian_pos = EE_target_pos[3]
            lb, ub = self.velDamperJointConstraints()                                  # :300
            A = self.qpA()                                                             # :313
            b = self.qpb(EE_target_pos, Trunk_target_pos).reshape((A.shape[0],))       # :316
            qp = QP(A, b, lb, ub, n_of_velocity_dimensions=self.n_velocity_dimensions)  # :320
            q_vel = qp.solveQP()
            if record is not None:
                record.append({"q": np.copy(self.current_joint_config), "A": A, "b": b, "lb": lb, "ub": ub,
                               "qdot": np.copy(q_vel), "result": qp.result})
            self.jointVelocitiestoConfig(q_vel, True)                                  # :325

        for i in range(len(self.current_joint_config)):                                # :328-330
            if 2 < i < 6:
                self.current_joint_config[i] = 0
        joint_config = self.current_joint_config
        self.updateState(joint_config, feedback=False)
        height_offset = (-self.EE_frame_pos[0][2] - self.EE_frame_pos[1][2] - self.EE_frame_pos[2][2]
                         - self.EE_frame_pos[3][2]) / 4                                # :336
        joint_config[2] = height_offset + self.foot_radius
        self.updateState(joint_config, feedback=False)
        joint_config = self.current_joint_config[7:]
        self.FL_leg = joint_config[0:3]                                                # :341-345
        self.FR_leg = joint_config[3:6]
        self.RL_leg = joint_config[6:9]
        self.RR_leg = joint_config[9:12]
        self.grip = joint_config[12:]
        self.fristQP = False                                                           # :347 (typo kept, D.11)
        self.dt = self.step_time

    # ------------------------------------------------------------------ :354-383
    def initialiseWBC(self, imu_data):
        joint_config = self.current_joint_config
        self.updateState(joint_config, imu_data, running=False)
        self._set_default_orientations()
        d = self.robot_data
        self.prev_trunk_ref = np.copy(d.oMf[self.trunk_frame_index].translation)
        self.old_ref_trunk_rot_matrix = np.copy(d.oMf[self.trunk_frame_index].rotation)
        for i in range(len(self.prev_EE_pos)):
            self.prev_EE_pos[i] = np.copy(d.oMf[self.end_effector_index_list_frame[i]].translation)
            EE_rot = np.copy(d.oMf[self.end_effector_index_list_frame[i]].rotation)
            hip_waist_rot = np.copy(d.oMf[self.trunk_frame_index].rotation)
            self.prev_EE_CoM_rot[i] = np.dot(hip_waist_rot.T, EE_rot)
        self.initial_trunk_pos = np.copy(d.oMf[self.trunk_frame_index].translation)
        self.initial_trunk_ori = np.copy(d.oMf[self.trunk_frame_index].rotation)
        self.initial_trunk_ori_euler = R.from_matrix(self.initial_trunk_ori).as_euler('xyz').reshape(3, 1)

    # ------------------------------------------------------------------ :387-428
    def _refresh(self, config):
        model, data = self.robot_model, self.robot_data
        pin.forwardKinematics(model, data, config)
        self.previous_joint_config = self.current_joint_config
        self.current_joint_config = config
        self.J = pin.computeJointJacobians(model, data, config)
        pin.framesForwardKinematics(model, data, config)
        pin.updateFramePlacements(model, data)
        self.trunk_frame_pos = np.copy(data.oMf[self.trunk_frame_index].translation)
        for i in range(len(self.EE_frame_pos)):
            self.EE_frame_pos[i] = np.copy(data.oMf[self.end_effector_index_list_frame[i]].translation)

    def updateState(self, joint_config, imu_data=0, feedback=True, running=False):
        if feedback is True and running is True:
            base_config = np.concatenate((self.current_joint_config[:3], imu_data), axis=0)
            config = np.concatenate((base_config, joint_config), axis=0)
        else:
            config = np.asarray(joint_config, dtype=float)
        self._refresh(config)
        if running is True:
            base_pos = self.trunkWorldPos()                                            # :414
            config = np.concatenate((base_pos, self.current_joint_config[3:]), axis=0)
            self._refresh(config)                                                      # :418-428

    # ------------------------------------------------------------------ :440-449
    def jointVelocitiestoConfig(self, joint_vel, update_model=False):
        new_config = pin.integrate(self.robot_model, self.current_joint_config, joint_vel * self.dt)
        if update_model is True:
            if self.initialised is True:
                self.updateState(new_config, feedback=False, running=True)
            if self.initialised is False:
                self.updateState(new_config, feedback=False, running=False)
        if update_model is False:
            return new_config

    # ------------------------------------------------------------------ A rows :474-490
    def endEffectorA2(self, frame_index):
        frame = pin.ReferenceFrame.LOCAL_WORLD_ALIGNED
        A = pin.getFrameJacobian(self.robot_model, self.robot_data,
                                 self.end_effector_index_list_frame[frame_index], frame).T
        A = A * self.cart_task_weight_EE_list[frame_index]
        A = np.dot(self.EE_weight[frame_index], A.T)
        self.EE_A_list[frame_index] = A

    def trunkA(self):
        self.trunk_A = pin.getFrameJacobian(self.robot_model, self.robot_data, self.trunk_frame_index,
                                            pin.ReferenceFrame.WORLD)
        self.trunk_A = np.dot(self.trunk_weight, self.trunk_A)
        self.trunk_A = self.trunk_A * self.cart_task_weight_Trunk

    # ------------------------------------------------------------------ bounds :572-637
    def velDamperJointConstraints(self):
        damping_coef = 0.01
        qi = 0.026
        qs = 0.015
        nv = self.n_velocity_dimensions
        lb = np.zeros((nv,))
        ub = np.zeros((nv,))
        lower_pos_lim = np.copy(self.robot_model.lowerPositionLimit)
        upper_pos_lim = np.copy(self.robot_model.upperPositionLimit)
        vel_lim = np.copy(self.robot_model.velocityLimit)
        for i in range(len(lower_pos_lim)):                                            # :589-596
            if i < 7:
                lower_pos_lim[i] = -5
                upper_pos_lim[i] = 5
                vel_lim[i] = 5
            if i >= (self.end_effector_index_list_joint[4] - 2 + 7):
                lower_pos_lim[i] = 0
                upper_pos_lim[i] = 0
        lower_pos_lim = np.delete(lower_pos_lim, 6)
        upper_pos_lim = np.delete(upper_pos_lim, 6)
        # :603-619 index joint i's limits with current_joint_config[i], but the limit arrays had entry 6 deleted and the
        # configuration did not (off-by-one from the first revolute joint on): reproduced by default;
        # compat_damper_off_by_one = False compares joint i with its own coordinate
        cfg = self.current_joint_config
        if not getattr(self, "compat_damper_off_by_one", True):
            cfg = np.delete(cfg, 6)
        for i in range(len(lower_pos_lim)):                                            # :603-619 (off-by-one, D.2)
            if cfg[i] <= (lower_pos_lim[i] + qi):
                lb[i] = -damping_coef * (cfg[i] - lower_pos_lim[i] - qs) / (qi - qs)
                if lb[i] > vel_lim[i]:
                    lb[i] = vel_lim[i]
                if lb[i] < -vel_lim[i]:
                    lb[i] = -vel_lim[i]
            else:
                lb[i] = -vel_lim[i]
            if cfg[i] >= (upper_pos_lim[i] - qi):
                ub[i] = damping_coef * (upper_pos_lim[i] - cfg[i] - qs) / (qi - qs)
                if ub[i] < -vel_lim[i]:
                    ub[i] = -vel_lim[i]
                if ub[i] > vel_lim[i]:
                    ub[i] = vel_lim[i]
            else:
                ub[i] = vel_lim[i]
        for i in range(len(lb)):                                                       # :621-625
            if lb[i] > 0:
                lb[i] = lb[i] * -1
            if ub[i] < 0:
                ub[i] = ub[i] * -1
        for i in range(len(lb)):                                                       # :627-630
            if i >= (self.end_effector_index_list_joint[4] - 2 + 6):
                lb[i] = 0
                ub[i] = 0
        return lb, ub

    # ------------------------------------------------------------------ constraints :640-836
    def footConstraint(self):
        C = pin.getFrameJacobian(self.robot_model, self.robot_data, self.end_effector_index_list_frame[0],
                                 pin.ReferenceFrame.WORLD)[:3]
        for i in range(len(self.end_effector_index_list_frame) - 2):
            Jtmp = pin.getFrameJacobian(self.robot_model, self.robot_data,
                                        self.end_effector_index_list_frame[i + 1], pin.ReferenceFrame.WORLD)[:3]
            C = np.concatenate((C, Jtmp), axis=0)
        return C, np.zeros(C.shape[0]), np.zeros(C.shape[0])

    def CoMConstraint(self):
        C = pin.jacobianCenterOfMass(self.robot_model, self.robot_data, self.current_joint_config)[:2]
        CoM_pos = self.robot_data.com[0][:2]
        FL_pos = self.EE_frame_pos[1][:2]
        RR_pos = self.EE_frame_pos[2][:2]
        Clb = ((RR_pos - CoM_pos) / self.dt).reshape(C.shape[0]) * 0.8
        Cub = ((FL_pos - CoM_pos) / self.dt).reshape(C.shape[0]) * 0.8
        return C, Clb, Cub

    def gripperOriConstraint(self):
        C = pin.getFrameJacobian(self.robot_model, self.robot_data, self.end_effector_index_list_frame[4],
                                 pin.ReferenceFrame.LOCAL)[3:]
        return C, np.zeros(C.shape[0]), np.zeros(C.shape[0])

    def trunkConstraint(self):
        C = pin.getFrameJacobian(self.robot_model, self.robot_data, self.trunk_frame_index,
                                 pin.ReferenceFrame.LOCAL_WORLD_ALIGNED)[2:]
        trunk_pos = self.trunk_frame_pos[2:]
        trunk_ori = R.from_matrix(np.copy(self.robot_data.oMf[self.trunk_frame_index].rotation))
        trunk_ori_euler = trunk_ori.as_euler('xyz')
        current_trunk_state = np.concatenate((trunk_pos, trunk_ori_euler), axis=0).reshape(4,)
        z_var = self.initial_trunk_pos[2] * 0.25
        roll_var = 1.5 * 0.1
        pitch_var = 1.5 * 0.1
        yaw_var = 1.5 * 0.1
        lb = np.zeros(C.shape[0])
        ub = np.zeros(C.shape[0])
        lb[0] = self.initial_trunk_pos[2] - z_var
        ub[0] = self.initial_trunk_pos[2] + z_var
        lb[1] = self.initial_trunk_ori_euler[0, 0] - roll_var
        ub[1] = self.initial_trunk_ori_euler[0, 0] + roll_var
        lb[2] = self.initial_trunk_ori_euler[1, 0] - pitch_var
        ub[2] = self.initial_trunk_ori_euler[1, 0] + pitch_var
        lb[3] = self.initial_trunk_ori_euler[2, 0] - yaw_var
        ub[3] = self.initial_trunk_ori_euler[2, 0] + yaw_var
        Clb = ((lb - current_trunk_state) / self.dt).reshape(C.shape[0]) * 0.5
        Cub = ((ub - current_trunk_state) / self.dt).reshape(C.shape[0]) * 0.5
        return C, Clb, Cub

    def EEConstraint(self, frame_index):
        C = pin.getFrameJacobian(self.robot_model, self.robot_data, self.end_effector_index_list_frame[frame_index],
                                 pin.ReferenceFrame.WORLD)[:3]
        return C, np.zeros(C.shape[0]), np.zeros(C.shape[0])

    def findConstraints(self):
        Cs, Clbs, Cubs = [], [], []
        if self.const_active_CoM is True:
            Cs_, l_, u_ = self.CoMConstraint(); Cs.append(Cs_); Clbs.append(l_); Cubs.append(u_)
        if self.const_active_Trunk is True:
            Cs_, l_, u_ = self.trunkConstraint(); Cs.append(Cs_); Clbs.append(l_); Cubs.append(u_)
        for flag, idx in ((self.const_active_FR_foot, 0), (self.const_active_FL_foot, 1),
                          (self.const_active_RR_foot, 2), (self.const_active_RL_foot, 3),
                          (self.const_active_GRIP, 4)):
            if flag is True:
                Cs_, l_, u_ = self.EEConstraint(idx); Cs.append(Cs_); Clbs.append(l_); Cubs.append(u_)
        C = np.concatenate(Cs, axis=0)
        Clb = np.concatenate(Clbs, axis=0)
        Cub = np.concatenate(Cubs, axis=0)
        return C.T, Clb, Cub                                                           # :836 (C.T, D.9)

    # ------------------------------------------------------------------ A stack :839-876, 1199-1206, 1271-1280
    def _ee_task_flags(self):
        return [self.task_active_FR_foot, self.task_active_FL_foot, self.task_active_RR_foot,
                self.task_active_RL_foot, self.task_active_GRIP]

    def qpCartesianA(self):
        A_list = []
        for i, flag in enumerate(self._ee_task_flags()):
            if flag is True:
                self.endEffectorA2(i)
                A_list.append(self.EE_A_list[i])
        if self.task_active_Trunk is True:
            self.trunkA()
            A_list.append(self.trunk_A)
        return np.concatenate(A_list, axis=0)

    def qpJointA(self):
        nv = self.n_velocity_dimensions
        U = np.identity(nv)
        a = np.ones(nv) * (1 / nv)
        A = U * a
        A = A * self.joint_task_weight
        return A

    def _joint_task_on(self):
        j = self.task_active_Joint
        return j is True or (isinstance(j, str) and j in ("PREV", "MANI", "HYBRID"))

    def qpA(self):
        A = self.qpCartesianA()
        if self._joint_task_on():
            A = np.concatenate((A, self.qpJointA()), axis=0)
        return A

    # ------------------------------------------------------------------ b targets :907-1196, 1209-1294
    def EndEffectorB2(self, target_cartesian_pos, frame_index):
        target = self.calcTargetVelEE3(target_cartesian_pos, self.default_EE_ori_list[frame_index], frame_index,
                                       self.EE_gains[frame_index])
        target = target * self.cart_task_weight_EE_list[frame_index]
        self.EE_b_list[frame_index] = target

    def TrunkB(self, target_cartesian_pos):
        self.trunk_B = self.calcTargetVelTrunk2(target_cartesian_pos, self.default_trunk_ori, self.trunk_frame_index,
                                                self.trunk_gain)
        self.trunk_B = self.trunk_B * self.cart_task_weight_Trunk

    def calcTargetVelTrunk2(self, target_pos, target_rot, frame_id, gain):
        """Robot_Wrapper4.py:948-1015."""
        gain_pos = gain[0:3, 0:3]
        gain_ori = gain[3:, 3:]
        target_pos = np.asarray(target_pos, dtype=float)
        ref_trunk_vel = (target_pos.reshape(3, 1) - self.prev_trunk_ref.reshape(3, 1)) / self.dt
        fk_trunk_pos = np.copy(self.robot_data.oMf[frame_id].translation)
        pos_vel = ref_trunk_vel + (np.dot(gain_pos, (target_pos.reshape(3, 1) - fk_trunk_pos.reshape(3, 1)).reshape(3, 1) / self.dt))
        fk_trunk_rot = R.from_matrix(np.copy(self.robot_data.oMf[frame_id].rotation))
        fk_trunk_quat = fk_trunk_rot.as_quat()
        ref_trunk_rotation = R.from_euler('xyz', target_rot.reshape(3,))
        ref_trunk_rot_matrix = ref_trunk_rotation.as_matrix()
        ref_trunk_quat = ref_trunk_rotation.as_quat()
        f, r = fk_trunk_quat, ref_trunk_quat
        quat_error = np.zeros((4,))
        quat_error[0] = (f[3] * r[0]) - (f[0] * r[3]) + (f[1] * r[2]) - (f[2] * r[1])
        quat_error[1] = (f[3] * r[1]) - (f[1] * r[3]) - (f[0] * r[2]) + (f[2] * r[0])
        quat_error[2] = (f[3] * r[2]) - (f[3] * r[2]) + (f[0] * r[1]) - (f[1] * r[0])        # :976 (cancelling terms kept)
        quat_error[3] = (f[3] * r[3]) + (f[0] * r[0]) + (f[1] * r[1]) + (f[2] * r[2])
        w = np.zeros((3,))
        for ii in range(3):
            w[ii] = gain_ori[ii, ii] * quat_error[ii]
        skew = np.dot(((ref_trunk_rot_matrix - self.old_ref_trunk_rot_matrix) / self.dt), ref_trunk_rot_matrix)  # :984 no transpose
        trunk_to_CoM_ori_ref = np.array([skew[2, 1], skew[0, 2], skew[1, 0]])
        ori_vel = (trunk_to_CoM_ori_ref + w).reshape(3, 1)
        self.prev_trunk_ref = target_pos                                               # :995-996
        self.old_ref_trunk_rot_matrix = ref_trunk_rot_matrix
        return np.concatenate((pos_vel, ori_vel), axis=0)

    def calcTargetVelEE3(self, target_pos, target_rot, i, gain):
        """Robot_Wrapper4.py:1052-1157."""
        gain_pos = gain[0:3, 0:3]
        target_pos = np.asarray(target_pos, dtype=float).reshape(3, 1)
        ref_EE_vel = (target_pos - self.prev_EE_pos[i].reshape(3, 1)) / self.dt
        fk_EE_pos = self.EE_frame_pos[i]
        pos_vel = ref_EE_vel + (np.dot(gain_pos, ((target_pos - fk_EE_pos.reshape(3, 1)) / self.dt)))   # :1070
        ref_EE_rotation = R.from_euler('xyz', target_rot.reshape(3,))
        ref_EE_rot = ref_EE_rotation.as_matrix()
        EE_CoM_Rot = ref_EE_rot                                                        # :1123
        skew = np.dot(((EE_CoM_Rot - self.prev_EE_CoM_rot[i]) / self.dt), EE_CoM_Rot.T)   # :1125
        w_ori_to_CoM_ref = np.array([skew[2, 1], skew[0, 2], skew[1, 0]])
        ori_vel = w_ori_to_CoM_ref.reshape(3, 1)                                       # :1133 (quaternion error discarded)
        self.prev_EE_pos[i] = target_pos                                               # :1151-1152
        self.prev_EE_CoM_rot[i] = EE_CoM_Rot
        return np.concatenate((pos_vel, ori_vel), axis=0)

    def qpCartesianB(self, target_cartesian_pos_EE, target_cartesian_pos_trunk):
        target_list = []
        for i, flag in enumerate(self._ee_task_flags()):
            if flag is True:
                self.EndEffectorB2(target_cartesian_pos_EE[i], i)
                target_list.append(self.EE_b_list[i])
        if self.task_active_Trunk is True:
            self.TrunkB(target_cartesian_pos_trunk)
            target_list.append(self.trunk_B)
        return np.concatenate(target_list, axis=0)

    def _manip(self, joint_id):
        J = pin.getJointJacobian(self.robot_model, self.robot_data, joint_id, pin.ReferenceFrame.LOCAL_WORLD_ALIGNED)
        return math.sqrt(np.linalg.det(np.dot(J, J.T)))

    def qpJointb(self):
        """Robot_Wrapper4.py:1209-1268 (modes True / "PREV" / "MANI" / "HYBRID", quirks D.4 kept)."""
        nv = self.n_velocity_dimensions
        mode = self.task_active_Joint
        if mode is True:
            u = np.zeros((nv, 1))
        if isinstance(mode, str) and mode == "PREV":
            u = np.delete(self.current_joint_config, 6).reshape((nv, 1))
        if isinstance(mode, str) and mode == "MANI":
            u = []
            q = np.copy(self.current_joint_config)
            deltaq = 0.0002
            for i in range(nv):
                joint_id = 1 if i < 6 else i + 1 - 5
                q[i] = q[i] + deltaq
                self.updateState(q, feedback=False)
                f1 = self._manip(joint_id)
                q[i] = q[i] - (deltaq * 2)
                self.updateState(q, feedback=False)
                f2 = self._manip(joint_id)
                u.append(0.5 * (f1 - f2) / deltaq)
            u = np.array(u).reshape((nv, 1))
            self.updateState(self.current_joint_config, feedback=False)
        if isinstance(mode, str) and mode == "HYBRID":
            u = np.delete(self.current_joint_config, 6).reshape((nv, 1))
            q = np.copy(self.current_joint_config)
            deltaq = 0.0002
            for i in range(len(u)):
                joint_id = i - 6
                if joint_id >= self.arm_base_id:
                    q[i] = q[i] + deltaq
                    self.updateState(q, feedback=False)
                    f1 = self._manip(joint_id)
                    q[i] = q[i] - (deltaq * 2)
                    self.updateState(q, feedback=False)
                    f2 = self._manip(joint_id)
                    u[i] = (0.5 * (f1 - f2) / deltaq)
        a = np.ones((nv, 1)) * (1 / nv)
        b = a * u
        b = b * self.joint_task_weight
        return b

    def qpb(self, target_cartesian_pos_EE, target_cartesian_pos_trunk):
        b = self.qpCartesianB(target_cartesian_pos_EE, target_cartesian_pos_trunk)
        if self._joint_task_on():
            b = np.concatenate((b, self.qpJointb()), axis=0)
        return b

    # ------------------------------------------------------------------ :1297-1327
    def trunkWorldPos(self):
        d = self.robot_data
        WRB = np.copy(d.oMf[self.trunk_frame_index].rotation)
        trunk_pos = np.copy(d.oMf[self.trunk_frame_index].translation)
        BPA = np.zeros((3, 1))
        for k in range(4):
            BPA = BPA + (np.copy(d.oMf[self.end_effector_index_list_frame[k]].translation).reshape(3, 1)
                         - trunk_pos.reshape(3, 1))
        BPA = BPA / 4
        WPA = (np.reshape(self.FR_target_cartesian_pos, (3, 1)) + np.reshape(self.FL_target_cartesian_pos, (3, 1))
               + np.reshape(self.RR_target_cartesian_pos, (3, 1)) + np.reshape(self.RL_target_cartesian_pos, (3, 1))) / 4
        return (WPA - np.dot(WRB, BPA)).reshape(3,)                                    # :1324 (D.10)

    # ------------------------------------------------------------------ :1330-1412
    def runWBC(self, base_config, target_cartesian_pos_EE=None, target_cartesian_pos_trunk=None):
        self.FR_target_cartesian_pos = target_cartesian_pos_EE[0]
        self.FL_target_cartesian_pos = target_cartesian_pos_EE[1]
        self.RR_target_cartesian_pos = target_cartesian_pos_EE[2]
        self.RL_target_cartesian_pos = target_cartesian_pos_EE[3]
        A = self.qpA()                                                                 # :1348
        b = self.qpb(target_cartesian_pos_EE, target_cartesian_pos_trunk).reshape((A.shape[0],))
        C, Clb, Cub = self.findConstraints()                                           # :1355
        lb, ub = self.velDamperJointConstraints()                                      # :1361
        if self.firstQP is True:                                                       # :1389-1394
            self.qp = QP(A, b, lb, ub, C, Clb, Cub, n_of_velocity_dimensions=self.n_velocity_dimensions)
            self.qp.max_iter = getattr(self, "max_qp_iterations", 200)
            q_vel = self.qp.solveQP()
            self.firstQP = False
        else:
            q_vel = self.qp.solveQPHotstart(A, b, lb, ub, C, Clb, Cub)
        self.last = {"A": A, "b": b, "C": C.T, "Clb": Clb, "Cub": Cub, "lb": lb, "ub": ub,
                     "qdot": np.copy(q_vel), "result": self.qp.result}
        joint_config = self.jointVelocitiestoConfig(q_vel, False)[7:]                  # :1397
        self.updateState(joint_config, base_config, running=True)                     # :1402
        FL_leg = joint_config[0:3]
        FR_leg = joint_config[3:6]
        RL_leg = joint_config[6:9]
        RR_leg = joint_config[9:12]
        grip = joint_config[12:]
        return FL_leg, FR_leg, RL_leg, RR_leg, grip

    # ------------------------------------------------------------------ :1415-1464
    def staticReachMode(self):
        self.trunk_weight = np.identity(6) * 1
        self.EE_weight = [np.identity(6) * 1 for _ in range(5)]
        self.cart_task_weight_FR = 100
        self.cart_task_weight_FL = 100
        self.cart_task_weight_RR = 100
        self.cart_task_weight_RL = 100
        self.cart_task_weight_GRIP = 1
        self.cart_task_weight_Trunk = 1
        self.cart_task_weight_EE_list = [100, 100, 100, 100, 1]
        self.joint_task_weight = 0.001
        self.trunk_gain = np.identity(6) * 0.8
        self.FL_gain = np.identity(6) * 0.8
        self.FR_gain = np.identity(6) * 0.8
        self.RL_gain = np.identity(6) * 0.8
        self.RR_gain = np.identity(6) * 0.8
        self.GRIP_gain = np.identity(6) * 0.05
        self.EE_gains = [self.FL_gain, self.FR_gain, self.RL_gain, self.RR_gain, self.GRIP_gain]
