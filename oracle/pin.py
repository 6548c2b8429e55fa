"""oracle.pin -- restatement of the slice of Pinocchio 2.x the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinocchio itself is a third-party
dependency that is NOT vendored under /root/reference and is not installable here
(version unpinned by the reference: no requirements file; Python 3.8 era => 2.x).  This
module restates its published algorithms for the calls made at
``wrappers/Robot_Wrapper4.py:21-24, 30, 37-39, 47-51, 66, 400-405, 441, 458-488, 641-758,
670, 1233-1258``.  Pinned by ``tests_NOT_FOR_USE/Jacobians.py:1-24`` (WORLD joint
Jacobians at neutral); everything else is parity-unpinned and checked by invariants.

API mirrors the pinocchio python module closely enough that ``oracle/robot_wrapper4.py``
reads like the reference:  ``buildModelFromUrdf``, ``Model.createData``, ``neutral``,
``forwardKinematics``, ``computeJointJacobians``, ``framesForwardKinematics``,
``updateFramePlacements``, ``getFrameJacobian``, ``getJointJacobian``, ``integrate``,
``jacobianCenterOfMass``, ``ReferenceFrame``, ``FrameType``.
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET

import numpy as np

DBL_MAX = 1.7976931348623157e308


class ReferenceFrame:
    WORLD = 0
    LOCAL = 1
    LOCAL_WORLD_ALIGNED = 2


WORLD = ReferenceFrame.WORLD
LOCAL = ReferenceFrame.LOCAL
LOCAL_WORLD_ALIGNED = ReferenceFrame.LOCAL_WORLD_ALIGNED


class FrameType:
    OP_FRAME = "OP_FRAME"
    JOINT = "JOINT"
    FIXED_JOINT = "FIXED_JOINT"
    BODY = "BODY"


JOINT = FrameType.JOINT
FIXED_JOINT = FrameType.FIXED_JOINT
BODY = FrameType.BODY


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def rpy_to_matrix(r, p, y):
    """URDF fixed-axis roll/pitch/yaw -> R = Rz(y) Ry(p) Rx(r) (urdfdom Rotation::setFromRPY)."""
    cr, sr = math.cos(r), math.sin(r)
    cp, sp = math.cos(p), math.sin(p)
    cy, sy = math.cos(y), math.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


class SE3:
    __slots__ = ("rotation", "translation")

    def __init__(self, R=None, p=None):
        self.rotation = np.eye(3) if R is None else np.array(R, dtype=float).reshape(3, 3)
        self.translation = np.zeros(3) if p is None else np.array(p, dtype=float).reshape(3)

    def __mul__(self, o):
        return SE3(self.rotation @ o.rotation, self.translation + self.rotation @ o.translation)

    def act(self, v):
        return self.rotation @ v + self.translation

    def copy(self):
        return SE3(self.rotation.copy(), self.translation.copy())


class Frame:
    def __init__(self, name, parent, placement, ftype):
        self.name, self.parent, self.placement, self.type = name, parent, placement, ftype


class Model:
    """Flat kinematic tree with a free-flyer root (pin.buildModelFromUrdf(urdf, JointModelFreeFlyer()))."""

    def __init__(self):
        self.name = ""
        self.names = ["universe"]
        self.jtypes = ["universe"]          # 'universe' | 'freeflyer' | 'revolute' | 'prismatic'
        self.axes = [np.zeros(3)]
        self.parents = [0]
        self.jointPlacements = [SE3()]
        self.idx_qs = [-1]
        self.idx_vs = [-1]
        self.nqs = [0]
        self.nvs = [0]
        self.masses = [0.0]                 # mass lumped on each joint (fixed links merged)
        self.coms = [np.zeros(3)]           # its centre of mass in the joint frame
        self.frames = [Frame("universe", 0, SE3(), FrameType.FIXED_JOINT)]
        self.collision_geoms = []           # pin.buildGeomFromUrdf(..., COLLISION) order: (link, type, radius)
        self.nq = 0
        self.nv = 0
        self._lo, self._up, self._vel, self._eff = [], [], [], []

    # -- construction -----------------------------------------------------------------
    def addJoint(self, parent, jtype, axis, placement, name, lo, up, vel, eff):
        nq, nv = (7, 6) if jtype == "freeflyer" else (1, 1)
        self.names.append(name)
        self.jtypes.append(jtype)
        self.axes.append(np.array(axis, dtype=float))
        self.parents.append(parent)
        self.jointPlacements.append(placement)
        self.idx_qs.append(self.nq)
        self.idx_vs.append(self.nv)
        self.nqs.append(nq)
        self.nvs.append(nv)
        self.masses.append(0.0)
        self.coms.append(np.zeros(3))
        self.nq += nq
        self.nv += nv
        self._lo += list(lo)
        self._up += list(up)
        self._vel += list(vel)
        self._eff += list(eff)
        return len(self.names) - 1

    def appendBody(self, joint_id, mass, com_in_joint):
        m0, c0 = self.masses[joint_id], self.coms[joint_id]
        m1 = m0 + mass
        if m1 > 0:
            self.coms[joint_id] = (m0 * c0 + mass * np.asarray(com_in_joint, dtype=float)) / m1
        self.masses[joint_id] = m1

    def finalize(self):
        self.njoints = len(self.names)
        self.nframes = len(self.frames)
        self.lowerPositionLimit = np.array(self._lo, dtype=float)
        self.upperPositionLimit = np.array(self._up, dtype=float)
        self.velocityLimit = np.array(self._vel, dtype=float)
        self.effortLimit = np.array(self._eff, dtype=float)
        # supports[j] = joints on the path root..j (inclusive); col_support[j] = their v-columns
        self.supports = []
        for j in range(self.njoints):
            chain, k = [], j
            while k > 0:
                chain.append(k)
                k = self.parents[k]
            self.supports.append(chain[::-1])
        return self

    # -- lookups (pin.Model.getJointId / getFrameId) -------------------------------------
    def getJointId(self, name):
        return self.names.index(name) if name in self.names else self.njoints

    def getFrameId(self, name, ftype=None):
        for i, f in enumerate(self.frames):
            if f.name == name and (ftype is None or f.type == ftype):
                return i
        return self.nframes

    def existFrame(self, name, ftype=None):
        return self.getFrameId(name, ftype) < self.nframes

    def createData(self):
        return Data(self)

    # -- (de)serialisation: same schema as the product's tree table JSON ------------------
    def to_dict(self):
        return {
            "name": self.name, "nq": self.nq, "nv": self.nv, "njoints": self.njoints,
            "joints": [{
                "name": self.names[j], "type": self.jtypes[j], "parent": self.parents[j],
                "axis": [float(x) for x in self.axes[j]],
                "R": [float(x) for x in self.jointPlacements[j].rotation.reshape(-1)],
                "p": [float(x) for x in self.jointPlacements[j].translation],
                "idx_q": self.idx_qs[j], "idx_v": self.idx_vs[j],
                "mass": float(self.masses[j]), "com": [float(x) for x in self.coms[j]],
            } for j in range(self.njoints)],
            "frames": [{
                "name": f.name, "type": f.type, "parent": f.parent,
                "R": [float(x) for x in f.placement.rotation.reshape(-1)],
                "p": [float(x) for x in f.placement.translation],
            } for f in self.frames],
            "lower": [float(x) for x in self.lowerPositionLimit],
            "upper": [float(x) for x in self.upperPositionLimit],
            "velocity": [float(x) for x in self.velocityLimit],
            "effort": [float(x) for x in self.effortLimit],
            "collision_geoms": [list(g) for g in self.collision_geoms],
        }

    @staticmethod
    def from_dict(d):
        m = Model()
        m.name = d.get("name", "")
        lo, up, vel, eff = d["lower"], d["upper"], d["velocity"], d["effort"]
        for j, jd in enumerate(d["joints"]):
            if j == 0:
                continue
            nq, nv = (7, 6) if jd["type"] == "freeflyer" else (1, 1)
            iq, iv = jd["idx_q"], jd["idx_v"]
            jid = m.addJoint(jd["parent"], jd["type"], jd["axis"], SE3(jd["R"], jd["p"]), jd["name"],
                             lo[iq:iq + nq], up[iq:iq + nq], vel[iv:iv + nv], eff[iv:iv + nv])
            m.masses[jid] = jd.get("mass", 0.0)
            m.coms[jid] = np.array(jd.get("com", [0, 0, 0]), dtype=float)
        m.frames = [Frame(f["name"], f["parent"], SE3(f["R"], f["p"]), f["type"]) for f in d["frames"]]
        m.collision_geoms = [tuple(g) for g in d.get("collision_geoms", [])]
        return m.finalize()

    @staticmethod
    def from_json(path):
        with open(path) as fh:
            return Model.from_dict(json.load(fh))


class Data:
    def __init__(self, model):
        self.oMi = [SE3() for _ in range(model.njoints)]
        self.oMf = [SE3() for _ in range(model.nframes)]
        self.J = np.zeros((6, model.nv))
        self.com = [np.zeros(3)]
        self.Jcom = np.zeros((3, model.nv))


# ---------------------------------------------------------------------------------------
# URDF -> Model, following pinocchio::urdf::buildModel + urdfdom's tree construction
# ---------------------------------------------------------------------------------------
def _floats(s, n=3, default=0.0):
    if s is None:
        return [default] * n
    return [float(x) for x in s.split()]


def buildModelFromUrdf(urdf_path, root_joint="freeflyer"):
    """pin.buildModelFromUrdf(urdf_path, pin.JointModelFreeFlyer())  (Robot_Wrapper4.py:21).

    urdfdom keeps joints in a std::map keyed by joint name, so each link's children are
    visited in ASCII order of the connecting joint's name; pinocchio walks that tree
    depth-first, adds one joint per non-fixed URDF joint (name = URDF joint name) and
    folds fixed joints into the parent joint (their offset accumulates into descendant
    placements, their inertia merges into the parent body, and they become FIXED_JOINT
    frames).  Every link also gets a BODY frame.
    """
    assert root_joint == "freeflyer"
    root = ET.parse(urdf_path).getroot()
    links = {l.attrib["name"]: l for l in root.findall("link")}
    joints = {j.attrib["name"]: j for j in root.findall("joint")}
    children = {name: [] for name in links}
    has_parent = set()
    for jname in sorted(joints):                     # std::map order
        j = joints[jname]
        children[j.find("parent").attrib["link"]].append(jname)
        has_parent.add(j.find("child").attrib["link"])
    roots = [l for l in links if l not in has_parent]
    assert len(roots) == 1, "URDF must have exactly one root link"
    root_link = roots[0]

    model = Model()
    model.name = root.attrib.get("name", "")
    inf = DBL_MAX
    jid = model.addJoint(0, "freeflyer", [0, 0, 0], SE3(), "root_joint",
                         [-inf] * 7, [inf] * 7, [inf] * 6, [inf] * 6)
    model.frames.append(Frame("root_joint", jid, SE3(), FrameType.JOINT))

    def add_body(link_name, joint_id, link_placement):
        link = links[link_name]
        inertial = link.find("inertial")
        if inertial is not None:
            mass = float(inertial.find("mass").attrib["value"])
            o = inertial.find("origin")
            xyz = _floats(o.attrib.get("xyz") if o is not None else None)
            model.appendBody(joint_id, mass, link_placement.act(np.array(xyz)))
        model.frames.append(Frame(link_name, joint_id, link_placement.copy(), FrameType.BODY))
        for col in link.findall("collision"):
            geom = col.find("geometry")
            shape = geom[0] if geom is not None and len(geom) else None
            if shape is not None:
                rad = shape.attrib.get("radius")
                model.collision_geoms.append((link_name, shape.tag, float(rad) if rad is not None else None))

    add_body(root_link, jid, SE3())

    def visit(link_name, joint_id, link_placement):
        # link_placement: pose of `link_name` expressed in the frame of joint `joint_id`
        for jname in children[link_name]:
            j = joints[jname]
            o = j.find("origin")
            xyz = _floats(o.attrib.get("xyz") if o is not None else None)
            rpy = _floats(o.attrib.get("rpy") if o is not None else None)
            origin = SE3(rpy_to_matrix(*rpy), xyz)
            placement = link_placement * origin
            child = j.find("child").attrib["link"]
            jtype = j.attrib["type"]
            if jtype == "fixed":
                model.frames.append(Frame(jname, joint_id, placement.copy(), FrameType.FIXED_JOINT))
                add_body(child, joint_id, placement)
                visit(child, joint_id, placement)
                continue
            axis = _floats(j.find("axis").attrib.get("xyz") if j.find("axis") is not None else "1 0 0")
            lim = j.find("limit")
            la = lim.attrib if lim is not None else {}
            if jtype in ("revolute", "prismatic"):
                lo, up = float(la.get("lower", 0.0)), float(la.get("upper", 0.0))
                kind = jtype
            else:
                raise NotImplementedError(
                    f"joint type {jtype!r} ({jname}) is outside the hot path: the reference's "
                    "indexing assumes nq == nv + 1 (see SURVEY Appendix A)")
            vel, eff = float(la.get("velocity", inf)), float(la.get("effort", inf))
            new_id = model.addJoint(joint_id, kind, axis, placement, jname, [lo], [up], [vel], [eff])
            model.frames.append(Frame(jname, new_id, SE3(), FrameType.JOINT))
            add_body(child, new_id, SE3())
            visit(child, new_id, SE3())

    visit(root_link, jid, SE3())
    return model.finalize()


# ---------------------------------------------------------------------------------------
# Algorithms
# ---------------------------------------------------------------------------------------
def neutral(model):
    q = np.zeros(model.nq)
    q[6] = 1.0
    return q


def quat_to_matrix(x, y, z, w):
    """Eigen::Quaternion::toRotationMatrix -- no normalisation (pin FK uses the raw quaternion)."""
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1 - (txx + tyy)],
    ])


def matrix_to_quat(R):
    """Eigen::Quaternion = Matrix3 (used by pin.integrate). Returns (x, y, z, w)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def axis_angle_matrix(axis, s, c):
    """Rotation about a unit axis (Eigen AngleAxis::toRotationMatrix, as JointModelRevoluteUnaligned)."""
    x, y, z = axis
    if abs(x) + abs(y) + abs(z) == 1.0 and max(x, y, z) == 1.0:
        # JointModelRX / RY / RZ: exact elementary rotations
        if x == 1.0:
            return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
        if y == 1.0:
            return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
        return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    t = 1 - c
    return np.array([
        [t * x * x + c, t * x * y - s * z, t * x * z + s * y],
        [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
        [t * x * z - s * y, t * y * z + s * x, t * z * z + c],
    ])


def joint_transform(model, j, q):
    jt = model.jtypes[j]
    iq = model.idx_qs[j]
    if jt == "freeflyer":
        return SE3(quat_to_matrix(q[iq + 3], q[iq + 4], q[iq + 5], q[iq + 6]), q[iq:iq + 3])
    if jt == "revolute":
        return SE3(axis_angle_matrix(model.axes[j], math.sin(q[iq]), math.cos(q[iq])), np.zeros(3))
    if jt == "prismatic":
        return SE3(np.eye(3), model.axes[j] * q[iq])
    raise ValueError(jt)


def forwardKinematics(model, data, q):
    """oMi = oM_parent * placement_i * JointTransform(q_i)   (Robot_Wrapper4.py:400)."""
    q = np.asarray(q, dtype=float)
    for j in range(1, model.njoints):
        liMi = model.jointPlacements[j] * joint_transform(model, j, q)
        data.oMi[j] = liMi if model.parents[j] == 0 else data.oMi[model.parents[j]] * liMi


def computeJointJacobians(model, data, q=None):
    """data.J (6 x nv), every column expressed in the WORLD frame (Robot_Wrapper4.py:403).

    Column of joint j with motion subspace S (joint frame): oMi[j].act(S) as a spatial motion =
    [p x (R a) + R v ; R a] -- i.e. the velocity of the body point coincident with the world origin.
    """
    if q is not None:
        forwardKinematics(model, data, q)
    J = np.zeros((6, model.nv))
    for j in range(1, model.njoints):
        R, p = data.oMi[j].rotation, data.oMi[j].translation
        iv = model.idx_vs[j]
        jt = model.jtypes[j]
        if jt == "freeflyer":
            for k in range(3):
                J[0:3, iv + k] = R[:, k]
                J[0:3, iv + 3 + k] = np.cross(p, R[:, k])
                J[3:6, iv + 3 + k] = R[:, k]
        elif jt == "revolute":
            a = R @ model.axes[j]
            J[0:3, iv] = np.cross(p, a)
            J[3:6, iv] = a
        elif jt == "prismatic":
            J[0:3, iv] = R @ model.axes[j]
    data.J = J
    return J


def updateFramePlacements(model, data):
    for i, f in enumerate(model.frames):
        data.oMf[i] = f.placement.copy() if f.parent == 0 else data.oMi[f.parent] * f.placement


def framesForwardKinematics(model, data, q):
    forwardKinematics(model, data, q)
    updateFramePlacements(model, data)


def _support_columns(model, joint_id):
    cols = []
    for j in model.supports[joint_id]:
        cols += list(range(model.idx_vs[j], model.idx_vs[j] + model.nvs[j]))
    return cols


def _express(Jw, cols, nv, rf, R, p):
    """Re-express WORLD columns `cols` of data.J for a frame placed at (R, p) in the world."""
    out = np.zeros((6, nv))
    lin, ang = Jw[0:3, cols], Jw[3:6, cols]
    if rf == ReferenceFrame.WORLD:
        out[0:3, cols], out[3:6, cols] = lin, ang
    elif rf == ReferenceFrame.LOCAL_WORLD_ALIGNED:
        out[0:3, cols] = lin - np.cross(p, ang, axis=0)
        out[3:6, cols] = ang
    elif rf == ReferenceFrame.LOCAL:
        out[0:3, cols] = R.T @ (lin - np.cross(p, ang, axis=0))
        out[3:6, cols] = R.T @ ang
    else:
        raise ValueError(rf)
    return out


def getFrameJacobian(model, data, frame_id, rf):
    """pin.getFrameJacobian(model, data, frame, rf) -> fresh 6 x nv array (Robot_Wrapper4.py:480,488,709,758).

    Requires computeJointJacobians + updateFramePlacements.  Only the columns supporting the
    frame's parent joint are non-zero.
    """
    f = model.frames[frame_id]
    oMf = data.oMi[f.parent] * f.placement     # pinocchio recomputes the placement from oMi
    cols = _support_columns(model, f.parent)
    return _express(data.J, cols, model.nv, rf, oMf.rotation, oMf.translation)


def getJointJacobian(model, data, joint_id, rf):
    """pin.getJointJacobian (Robot_Wrapper4.py:1233,1254,1486)."""
    oMi = data.oMi[joint_id]
    cols = _support_columns(model, joint_id)
    return _express(data.J, cols, model.nv, rf, oMi.rotation, oMi.translation)


TAYLOR_T2 = math.sqrt(math.sqrt(np.finfo(float).eps))  # pinocchio TaylorSeriesExpansion precision<3> ~ eps^(1/4)


def exp3(w):
    """pinocchio::exp3 (Rodrigues)."""
    w = np.asarray(w, dtype=float)
    t2 = float(w @ w)
    t = math.sqrt(t2)
    if t < TAYLOR_T2:
        alpha_vxvx = 0.5 - t2 / 24.0
        alpha_vx = 1.0 - t2 / 6.0
        ct = 1.0 - t2 / 2.0
    else:
        st, ct = math.sin(t), math.cos(t)
        alpha_vxvx = (1.0 - ct) / t2
        alpha_vx = st / t
    R = alpha_vxvx * np.outer(w, w)
    R[0, 1] -= alpha_vx * w[2]
    R[1, 0] += alpha_vx * w[2]
    R[0, 2] += alpha_vx * w[1]
    R[2, 0] -= alpha_vx * w[1]
    R[1, 2] -= alpha_vx * w[0]
    R[2, 1] += alpha_vx * w[0]
    R[0, 0] += ct
    R[1, 1] += ct
    R[2, 2] += ct
    return R


def exp6(nu):
    """pinocchio::exp6: SE3 exponential of a spatial motion (v, w)."""
    v, w = np.asarray(nu[:3], dtype=float), np.asarray(nu[3:6], dtype=float)
    t2 = float(w @ w)
    t = math.sqrt(t2)
    if t < TAYLOR_T2:
        alpha_wxv = 0.5 - t2 / 24.0
        alpha_v = 1.0 - t2 / 6.0
        alpha_w = 1.0 / 6.0 - t2 / 120.0
    else:
        st, ct = math.sin(t), math.cos(t)
        alpha_wxv = (1.0 - ct) / t2
        alpha_v = st / t
        alpha_w = (1.0 - alpha_v) / t2
    p = alpha_v * v + alpha_wxv * np.cross(w, v) + alpha_w * float(w @ v) * w
    return SE3(exp3(w), p)


def integrate(model, q, v):
    """pin.integrate(model, q, v) = q (+) v  (Robot_Wrapper4.py:441).

    Free-flyer: M1 = M0 * exp6(v) with v = (linear in body frame, angular in body frame); the
    quaternion of M1 is sign-aligned with the input quaternion and first-order renormalised
    (pinocchio SpecialEuclideanOperationTpl<3>::integrate_impl).  1-DoF joints: q + v.
    """
    q = np.asarray(q, dtype=float)
    v = np.asarray(v, dtype=float)
    out = np.empty(model.nq)
    for j in range(1, model.njoints):
        iq, iv = model.idx_qs[j], model.idx_vs[j]
        if model.jtypes[j] == "freeflyer":
            quat = q[iq + 3:iq + 7]
            M0 = SE3(quat_to_matrix(*quat), q[iq:iq + 3])
            M1 = M0 * exp6(v[iv:iv + 6])
            out[iq:iq + 3] = M1.translation
            rq = matrix_to_quat(M1.rotation)
            if float(rq @ quat) < 0:
                rq = -rq
            n2 = float(rq @ rq)
            rq = rq * ((3.0 - n2) / 2.0)           # quaternion::firstOrderNormalize
            out[iq + 3:iq + 7] = rq
        else:
            out[iq] = q[iq] + v[iv]
    return out


def jacobianCenterOfMass(model, data, q):
    """pin.jacobianCenterOfMass(model, data, q) -> 3 x nv; fills data.com[0] (Robot_Wrapper4.py:670)."""
    forwardKinematics(model, data, q)
    computeJointJacobians(model, data)
    M = sum(model.masses)
    com = np.zeros(3)
    Jc = np.zeros((3, model.nv))
    for j in range(1, model.njoints):
        m = model.masses[j]
        if m == 0.0:
            continue
        c = data.oMi[j].act(model.coms[j])
        com += m * c
        cols = _support_columns(model, j)
        lin, ang = data.J[0:3, cols], data.J[3:6, cols]
        Jc[:, cols] += m * (lin - np.cross(c, ang, axis=0))
    data.com[0] = com / M
    data.Jcom = Jc / M
    return data.Jcom
