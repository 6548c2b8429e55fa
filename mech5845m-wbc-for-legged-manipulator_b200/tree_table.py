"""URDF -> flat kinematic-tree table (host side, runs once per model).

This replaces ``pin.buildModelFromUrdf(urdf_path, pin.JointModelFreeFlyer())`` and the name->index
look-ups of ``RobotModel.__init__`` (reference ``wrappers/Robot_Wrapper4.py:19-52``).  The table is
what the CUDA kernels consume: parent indices, joint placements, axes, q/v offsets, frame offsets,
limits -- all flat arrays, uploaded through the C ABI (``include/wbc_b200.h: WbcTreeTable``).

Pinocchio is used if it is importable (it is not in this image); otherwise the URDF is walked with
``xml.etree`` following Pinocchio's conventions:

* joints are numbered depth-first, children of a link visited in ASCII order of the connecting
  *joint name* (urdfdom keeps joints in a name-keyed map) -- this reproduces the reference's own
  hard-coded slices FL, FR, RL, RR, arm (``Robot_Wrapper4.py:341-345``);
* joint 0 is the universe, joint 1 the free-flyer ``root_joint`` (q = x y z qx qy qz qw);
* fixed joints are folded into their parent joint: their offset accumulates into the placement of
  descendant joints and they become FIXED_JOINT frames (parent joint + SE3 offset);
* masses / centres of mass of fixed links are merged into the parent joint's body.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

JT_UNIVERSE, JT_FREEFLYER, JT_REVOLUTE, JT_PRISMATIC = 0, 1, 2, 3
_JT_NAME = {JT_UNIVERSE: "universe", JT_FREEFLYER: "freeflyer", JT_REVOLUTE: "revolute", JT_PRISMATIC: "prismatic"}
_JT_CODE = {v: k for k, v in _JT_NAME.items()}
DBL_MAX = 1.7976931348623157e308

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _vec(text, default="0 0 0"):
    return np.array([float(t) for t in (text if text is not None else default).split()])


class TreeTable:
    """Flat kinematic tree.  All arrays are NumPy; ``to_dict`` / ``from_dict`` give the JSON schema."""

    def __init__(self):
        self.name = ""
        self.joint_names = ["universe"]
        self.jtype = [JT_UNIVERSE]
        self.parent = [0]
        self.axis = [np.zeros(3)]
        self.placement_R = [np.eye(3)]
        self.placement_p = [np.zeros(3)]
        self.idx_q = [-1]
        self.idx_v = [-1]
        self.mass = [0.0]
        self.com = [np.zeros(3)]
        self.frame_names = ["universe"]
        self.frame_type = ["FIXED_JOINT"]
        self.frame_parent = [0]
        self.frame_R = [np.eye(3)]
        self.frame_p = [np.zeros(3)]
        self.lower = []
        self.upper = []
        self.velocity = []
        self.effort = []
        self.collision_geoms = []
        self.nq = 0
        self.nv = 0

    # ------------------------------------------------------------------------------ sizes / lookups
    @property
    def njoints(self):
        return len(self.joint_names)

    @property
    def nframes(self):
        return len(self.frame_names)

    def getJointId(self, name):
        return self.joint_names.index(name) if name in self.joint_names else self.njoints

    def getFrameId(self, name, ftype=None):
        for i, (n, t) in enumerate(zip(self.frame_names, self.frame_type)):
            if n == name and (ftype is None or t == ftype):
                return i
        return self.nframes

    def depth(self):
        d = [0] * self.njoints
        for j in range(1, self.njoints):
            d[j] = d[self.parent[j]] + 1
        return d

    def support_mask(self, joint_id):
        """Bit k set <=> velocity column k moves joint `joint_id` (columns of its ancestors incl. itself)."""
        mask = 0
        j = joint_id
        while j > 0:
            nvj = 6 if self.jtype[j] == JT_FREEFLYER else 1
            for k in range(self.idx_v[j], self.idx_v[j] + nvj):
                mask |= 1 << k
            j = self.parent[j]
        return mask

    # ------------------------------------------------------------------------------ construction
    def _add_joint(self, parent, jtype, axis, R, p, name, lo, up, vel, eff):
        nq, nv = (7, 6) if jtype == JT_FREEFLYER else (1, 1)
        self.joint_names.append(name)
        self.jtype.append(jtype)
        self.parent.append(parent)
        self.axis.append(np.asarray(axis, dtype=float))
        self.placement_R.append(np.asarray(R, dtype=float).reshape(3, 3))
        self.placement_p.append(np.asarray(p, dtype=float).reshape(3))
        self.idx_q.append(self.nq)
        self.idx_v.append(self.nv)
        self.mass.append(0.0)
        self.com.append(np.zeros(3))
        self.nq += nq
        self.nv += nv
        self.lower += list(lo)
        self.upper += list(up)
        self.velocity += list(vel)
        self.effort += list(eff)
        return self.njoints - 1

    def _add_frame(self, name, ftype, parent, R, p):
        self.frame_names.append(name)
        self.frame_type.append(ftype)
        self.frame_parent.append(parent)
        self.frame_R.append(np.asarray(R, dtype=float).reshape(3, 3).copy())
        self.frame_p.append(np.asarray(p, dtype=float).reshape(3).copy())

    def _add_mass(self, jid, mass, com):
        m0 = self.mass[jid]
        m1 = m0 + mass
        if m1 > 0.0:
            self.com[jid] = (m0 * self.com[jid] + mass * np.asarray(com, dtype=float)) / m1
        self.mass[jid] = m1

    @classmethod
    def from_urdf(cls, urdf_path):
        try:                                           # pragma: no cover - pinocchio is absent in this image
            import pinocchio  # noqa: F401
            return cls._from_pinocchio(urdf_path)
        except ImportError:
            pass
        tree = ET.parse(urdf_path).getroot()
        link_xml = {l.attrib["name"]: l for l in tree.findall("link")}
        joint_xml = {j.attrib["name"]: j for j in tree.findall("joint")}
        kids, child_links = {n: [] for n in link_xml}, set()
        for jn in sorted(joint_xml):
            kids[joint_xml[jn].find("parent").attrib["link"]].append(jn)
            child_links.add(joint_xml[jn].find("child").attrib["link"])
        roots = [n for n in link_xml if n not in child_links]
        if len(roots) != 1:
            raise ValueError(f"{urdf_path}: expected one root link, found {roots}")

        t = cls()
        t.name = tree.attrib.get("name", "")
        root_id = t._add_joint(0, JT_FREEFLYER, [0, 0, 0], np.eye(3), np.zeros(3), "root_joint",
                               [-DBL_MAX] * 7, [DBL_MAX] * 7, [DBL_MAX] * 6, [DBL_MAX] * 6)
        t._add_frame("root_joint", "JOINT", root_id, np.eye(3), np.zeros(3))

        def body(link, jid, R, p):
            x = link_xml[link]
            ine = x.find("inertial")
            if ine is not None:
                o = ine.find("origin")
                c = _vec(o.attrib.get("xyz") if o is not None else None)
                t._add_mass(jid, float(ine.find("mass").attrib["value"]), R @ c + p)
            t._add_frame(link, "BODY", jid, R, p)
            for col in x.findall("collision"):
                g = col.find("geometry")
                if g is not None and len(g):
                    rad = g[0].attrib.get("radius")
                    t.collision_geoms.append([link, g[0].tag, float(rad) if rad is not None else None])

        body(roots[0], root_id, np.eye(3), np.zeros(3))
        # explicit stack, entries pushed in reverse so they pop in name order (depth-first, pre-order)
        stack = [(jn, root_id, np.eye(3), np.zeros(3)) for jn in reversed(kids[roots[0]])]
        while stack:
            jn, jid, R, p = stack.pop()
            x = joint_xml[jn]
            o = x.find("origin")
            oR = _rpy(*_vec(o.attrib.get("rpy") if o is not None else None))
            op = _vec(o.attrib.get("xyz") if o is not None else None)
            R1, p1 = R @ oR, R @ op + p
            child = x.find("child").attrib["link"]
            kind = x.attrib["type"]
            if kind == "fixed":
                t._add_frame(jn, "FIXED_JOINT", jid, R1, p1)
                body(child, jid, R1, p1)
                nxt = (jid, R1, p1)
            elif kind in ("revolute", "prismatic"):
                ax = x.find("axis")
                lim = x.find("limit")
                la = lim.attrib if lim is not None else {}
                new = t._add_joint(jid, JT_REVOLUTE if kind == "revolute" else JT_PRISMATIC,
                                   _vec(ax.attrib.get("xyz") if ax is not None else None, "1 0 0"), R1, p1, jn,
                                   [float(la.get("lower", 0.0))], [float(la.get("upper", 0.0))],
                                   [float(la.get("velocity", DBL_MAX))], [float(la.get("effort", DBL_MAX))])
                t._add_frame(jn, "JOINT", new, np.eye(3), np.zeros(3))
                body(child, new, np.eye(3), np.zeros(3))
                nxt = (new, np.eye(3), np.zeros(3))
            else:
                raise NotImplementedError(
                    f"URDF joint type {kind!r} ({jn}): the reference path assumes nq == nv + 1 "
                    "(free-flyer + 1-DoF revolute/prismatic joints only)")
            stack += [(cj,) + nxt for cj in reversed(kids[child])]
        return t

    @classmethod
    def _from_pinocchio(cls, urdf_path):               # pragma: no cover
        import pinocchio as pin
        m = pin.buildModelFromUrdf(urdf_path, pin.JointModelFreeFlyer())
        t = cls()
        t.name = m.name
        for j in range(1, m.njoints):
            jm = m.joints[j]
            sn = jm.shortname()
            if sn == "JointModelFreeFlyer":
                jt, axis = JT_FREEFLYER, np.zeros(3)
            elif sn.startswith("JointModelR"):
                jt = JT_REVOLUTE
                axis = {"X": [1, 0, 0], "Y": [0, 1, 0], "Z": [0, 0, 1]}.get(sn[-1]) or np.array(jm.extract().axis)
            elif sn.startswith("JointModelP"):
                jt = JT_PRISMATIC
                axis = {"X": [1, 0, 0], "Y": [0, 1, 0], "Z": [0, 0, 1]}.get(sn[-1]) or np.array(jm.extract().axis)
            else:
                raise NotImplementedError(sn)
            pl = m.jointPlacements[j]
            nq, nv = jm.nq, jm.nv
            t._add_joint(m.parents[j], jt, axis, pl.rotation, pl.translation, m.names[j],
                         m.lowerPositionLimit[jm.idx_q:jm.idx_q + nq], m.upperPositionLimit[jm.idx_q:jm.idx_q + nq],
                         m.velocityLimit[jm.idx_v:jm.idx_v + nv], m.effortLimit[jm.idx_v:jm.idx_v + nv])
            t.mass[j] = m.inertias[j].mass
            t.com[j] = np.array(m.inertias[j].lever)
        names = {pin.FrameType.JOINT: "JOINT", pin.FrameType.FIXED_JOINT: "FIXED_JOINT",
                 pin.FrameType.BODY: "BODY", pin.FrameType.OP_FRAME: "OP_FRAME"}
        for f in list(m.frames)[1:]:
            t._add_frame(f.name, names.get(f.type, "OP_FRAME"), f.parent, f.placement.rotation, f.placement.translation)
        return t

    # ------------------------------------------------------------------------------ (de)serialisation
    def to_dict(self):
        return {
            "name": self.name, "nq": self.nq, "nv": self.nv, "njoints": self.njoints,
            "joints": [{
                "name": self.joint_names[j], "type": _JT_NAME[self.jtype[j]], "parent": self.parent[j],
                "axis": [float(v) for v in self.axis[j]],
                "R": [float(v) for v in self.placement_R[j].reshape(-1)],
                "p": [float(v) for v in self.placement_p[j]],
                "idx_q": self.idx_q[j], "idx_v": self.idx_v[j],
                "mass": float(self.mass[j]), "com": [float(v) for v in self.com[j]],
            } for j in range(self.njoints)],
            "frames": [{
                "name": self.frame_names[i], "type": self.frame_type[i], "parent": self.frame_parent[i],
                "R": [float(v) for v in self.frame_R[i].reshape(-1)], "p": [float(v) for v in self.frame_p[i]],
            } for i in range(self.nframes)],
            "lower": [float(v) for v in self.lower], "upper": [float(v) for v in self.upper],
            "velocity": [float(v) for v in self.velocity], "effort": [float(v) for v in self.effort],
            "collision_geoms": [list(g) for g in self.collision_geoms],
        }

    @classmethod
    def from_dict(cls, d):
        t = cls()
        t.name = d.get("name", "")
        for j, jd in enumerate(d["joints"]):
            if j == 0:
                continue
            jt = _JT_CODE[jd["type"]]
            nq, nv = (7, 6) if jt == JT_FREEFLYER else (1, 1)
            iq, iv = jd["idx_q"], jd["idx_v"]
            jid = t._add_joint(jd["parent"], jt, jd["axis"], jd["R"], jd["p"], jd["name"],
                               d["lower"][iq:iq + nq], d["upper"][iq:iq + nq],
                               d["velocity"][iv:iv + nv], d["effort"][iv:iv + nv])
            t.mass[jid] = float(jd.get("mass", 0.0))
            t.com[jid] = np.array(jd.get("com", [0.0, 0.0, 0.0]), dtype=float)
        for f in d["frames"][1:]:
            t._add_frame(f["name"], f["type"], f["parent"], f["R"], f["p"])
        t.collision_geoms = [list(g) for g in d.get("collision_geoms", [])]
        return t

    def save(self, path):
        with open(path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=1)

    @classmethod
    def load(cls, path_or_name):
        """Load a table from a JSON path, a URDF path, or the name of a table shipped in ``data/``."""
        p = str(path_or_name)
        if p.endswith(".urdf"):
            shipped = os.path.join(DATA_DIR, os.path.splitext(os.path.basename(p))[0] + ".json")
            if os.path.exists(p):
                return cls.from_urdf(p)
            if os.path.exists(shipped):              # URDF not on this machine: use the pre-extracted table
                p = shipped
            else:
                raise FileNotFoundError(p)
        if not os.path.exists(p):
            p = os.path.join(DATA_DIR, p if p.endswith(".json") else p + ".json")
        with open(p) as fh:
            return cls.from_dict(json.load(fh))
