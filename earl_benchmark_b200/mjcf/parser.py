"""MJCF reader: <include> expansion, <compiler>/<option>, nested <default> classes with `childclass`,
bodies / inertials / joints / geoms / sites / mocap, position actuators, weld + joint equalities.

Host-side, run once per task: the reference loads these files through mujoco-py
(`load_model_from_path`, reference `earl_benchmark/envs/sawyer_door.py:67-70`,
`kitchen_assets/adept_envs/adept_envs/simulation/sim_robot.py:67-69`); here they are turned into a
structure-of-arrays model for the CUDA engine (mjcf/compile.py).  Rendering-only elements (<texture>,
<material>, <visual>, <light>, <camera>) are ignored, so checkouts with missing texture blobs still load
(SURVEY.md Appendix A #15).
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

GEOM_TYPES = ("plane", "hfield", "sphere", "capsule", "ellipsoid", "cylinder", "box", "mesh")
IGNORED = {"texture", "material", "visual", "light", "camera", "size", "statistic", "custom", "keyframe", "sensor",
           "contact", "tendon"}

# MuJoCo element defaults (only attributes this engine consumes)
BUILTIN = {
    "geom": dict(type="sphere", contype="1", conaffinity="1", condim="3", group="0", priority="0", friction="1 0.005 0.0001",
                 solmix="1", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", margin="0", gap="0", density="1000"),
    "joint": dict(type="hinge", pos="0 0 0", axis="0 0 1", limited="false", range="0 0", damping="0", armature="0",
                  stiffness="0", springref="0", ref="0", frictionloss="0", margin="0", solreflimit="0.02 1",
                  solimplimit="0.9 0.95 0.001 0.5 2", solreffriction="0.02 1", solimpfriction="0.9 0.95 0.001 0.5 2"),
    "site": dict(pos="0 0 0", size="0.005"),
    "position": dict(kp="1", ctrllimited="false", ctrlrange="0 0", forcelimited="false", forcerange="0 0", gear="1"),
    "motor": dict(ctrllimited="false", ctrlrange="0 0", forcelimited="false", forcerange="0 0", gear="1"),
    "velocity": dict(kv="1", ctrllimited="false", ctrlrange="0 0", forcelimited="false", forcerange="0 0", gear="1"),
    "general": dict(ctrllimited="false", ctrlrange="0 0", forcelimited="false", forcerange="0 0", gear="1"),
    "mesh": dict(scale="1 1 1"),
    "equality": dict(solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", active="true"),
}


def _floats(s, n=None):
    v = np.array([float(x) for x in str(s).split()], dtype=np.float64)
    if n is not None and len(v) < n:
        v = np.concatenate([v, np.zeros(n - len(v))])
    return v


class DefaultClass:
    def __init__(self, name, parent=None):
        self.name, self.parent = name, parent
        self.attrs = {}  # element tag -> dict

    def resolved(self, tag):
        out = dict(self.parent.resolved(tag)) if self.parent is not None else dict(BUILTIN.get(tag, {}))
        out.update(self.attrs.get(tag, {}))
        return out


class Spec:
    """Parsed model: flat lists of dict-like records with string attributes already default-resolved."""

    def __init__(self):
        self.compiler = dict(angle="degree", eulerseq="xyz", inertiafromgeom="auto", inertiagrouprange="0 5", meshdir=None,
                             coordinate="local", autolimits="false")
        self.option = dict(timestep="0.002", gravity="0 0 -9.81", iterations="100", tolerance="1e-8", solver="Newton",
                           cone="pyramidal", jacobian="auto", impratio="1", integrator="Euler", noslip_iterations="0")
        self.classes = {"main": DefaultClass("main")}
        self.meshes = {}     # name -> dict(file=abs path, scale)
        self.bodies = []     # dicts: name, parent (index), pos, quat, mocap, inertial|None, joints[], geoms[], sites[]
        self.actuators = []
        self.equalities = []
        self.model_dir = None


def _expand_includes(elem, base_dir):
    """Replace every <include file=...> by the children of the included file's root (paths relative to the
    MAIN model file's directory, as MuJoCo resolves them)."""
    out = []
    for child in list(elem):
        if child.tag == "include":
            path = os.path.normpath(os.path.join(base_dir, child.attrib["file"]))
            root = ET.parse(path).getroot()
            _expand_includes(root, base_dir)
            out.extend(list(root))
        else:
            _expand_includes(child, base_dir)
            out.append(child)
    for c in list(elem):
        elem.remove(c)
    for c in out:
        elem.append(c)


class Parser:
    def __init__(self, path):
        self.path = os.path.abspath(path)
        self.spec = Spec()
        self.spec.model_dir = os.path.dirname(self.path)

    # -------------------------------------------------------------- orientation helpers
    def _angle(self, v):
        return np.deg2rad(v) if self.spec.compiler["angle"] == "degree" else v

    def quat_of(self, a):
        """quat (w,x,y,z) of an element from quat | euler | axisangle | xyaxes | zaxis."""
        if "quat" in a:
            q = _floats(a["quat"])
            return q / np.linalg.norm(q)
        if "euler" in a:
            e = self._angle(_floats(a["euler"]))
            q = np.array([1.0, 0, 0, 0])
            for ch, ang in zip(self.spec.compiler["eulerseq"], e):
                ax = {"x": 0, "y": 1, "z": 2}[ch.lower()]
                r = np.zeros(4)
                r[0], r[1 + ax] = np.cos(ang / 2), np.sin(ang / 2)
                # lowercase = intrinsic (rotating frame): post-multiply; uppercase = extrinsic: pre-multiply
                q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
            return q
        if "axisangle" in a:
            v = _floats(a["axisangle"])
            ang = self._angle(v[3])
            ax = v[:3] / np.linalg.norm(v[:3])
            return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        if "xyaxes" in a:
            v = _floats(a["xyaxes"])
            x = v[:3] / np.linalg.norm(v[:3])
            y = v[3:] - x * np.dot(x, v[3:])
            y /= np.linalg.norm(y)
            return mat2quat(np.stack([x, y, np.cross(x, y)], axis=1))
        if "zaxis" in a:
            z = _floats(a["zaxis"])
            z /= np.linalg.norm(z)
            return quat_z2vec(z)
        return np.array([1.0, 0, 0, 0])

    # -------------------------------------------------------------- sections
    def parse(self):
        root = ET.parse(self.path).getroot()
        _expand_includes(root, self.spec.model_dir)
        # pass 1: compiler / option / defaults / assets (order-independent in MuJoCo: last one wins)
        for sec in root:
            if sec.tag == "compiler":
                self.spec.compiler.update(sec.attrib)
            elif sec.tag == "option":
                self.spec.option.update(sec.attrib)
        for sec in root:
            if sec.tag == "default":
                self._defaults(sec, self.spec.classes["main"])
        for sec in root:
            if sec.tag == "asset":
                for m in sec:
                    if m.tag == "mesh":
                        a = self.spec.classes["main"].resolved("mesh")
                        a.update(m.attrib)
                        mdir = self.spec.compiler.get("meshdir")
                        base = os.path.join(self.spec.model_dir, mdir) if mdir else self.spec.model_dir
                        name = a.get("name") or os.path.splitext(os.path.basename(a["file"]))[0]
                        self.spec.meshes[name] = dict(file=os.path.normpath(os.path.join(base, a["file"])),
                                                      scale=_floats(a["scale"]))
        # pass 2: world tree, actuators, equalities
        world = dict(name="world", parent=-1, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), mocap=False, inertial=None,
                     joints=[], geoms=[], sites=[])
        self.spec.bodies.append(world)
        for sec in root:
            if sec.tag == "worldbody":
                self._body_children(sec, 0, None)
        for sec in root:
            if sec.tag == "actuator":
                for a in sec:
                    rec = self._resolve(a, None)
                    rec["tag"] = a.tag
                    self.spec.actuators.append(rec)
            elif sec.tag == "equality":
                for e in sec:
                    rec = dict(self.spec.classes[e.attrib.get("class", "main")].resolved("equality"))
                    rec.update(e.attrib)
                    rec["tag"] = e.tag
                    self.spec.equalities.append(rec)
        return self.spec

    def _defaults(self, elem, cls):
        for c in elem:
            if c.tag == "default":
                name = c.attrib.get("class")
                if name is None:            # nested <default> without class inside top-level: same class
                    self._defaults(c, cls)
                    continue
                sub = self.spec.classes.get(name) or DefaultClass(name, cls)
                self.spec.classes[name] = sub
                self._defaults(c, sub)
            elif c.tag not in IGNORED:
                cls.attrs.setdefault(c.tag, {}).update(c.attrib)

    def _resolve(self, elem, childclass):
        cname = elem.attrib.get("class", childclass or "main")
        rec = self.spec.classes[cname].resolved(elem.tag)
        rec.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return rec

    def _body_children(self, elem, body_idx, childclass):
        body = self.spec.bodies[body_idx]
        for c in elem:
            if c.tag == "body":
                cc = c.attrib.get("childclass", childclass)
                rec = dict(name=c.attrib.get("name", f"body{len(self.spec.bodies)}"), parent=body_idx,
                           pos=_floats(c.attrib.get("pos", "0 0 0")), quat=self.quat_of(c.attrib),
                           mocap=c.attrib.get("mocap", "false") == "true", inertial=None, joints=[], geoms=[], sites=[])
                self.spec.bodies.append(rec)
                self._body_children(c, len(self.spec.bodies) - 1, cc)
            elif c.tag == "inertial":
                a = c.attrib
                I = None
                if "fullinertia" in a:
                    f = _floats(a["fullinertia"])
                    I = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                body["inertial"] = dict(pos=_floats(a.get("pos", "0 0 0")), quat=self.quat_of(a), mass=float(a["mass"]),
                                        diag=_floats(a["diaginertia"]) if "diaginertia" in a else None, full=I)
            elif c.tag in ("joint", "freejoint"):
                rec = self._resolve(c, childclass) if c.tag == "joint" else dict(BUILTIN["joint"], type="free", **c.attrib)
                rec.setdefault("name", f"joint{sum(len(b['joints']) for b in self.spec.bodies)}")
                body["joints"].append(rec)
            elif c.tag == "geom":
                rec = self._resolve(c, childclass)
                rec["quat_resolved"] = self.quat_of(rec)
                body["geoms"].append(rec)
            elif c.tag == "site":
                rec = self._resolve(c, childclass)
                rec["quat_resolved"] = self.quat_of(rec)
                body["sites"].append(rec)


# -------------------------------------------------------------------------------------------- small math

def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def quat2mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def mat2quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def quat_z2vec(v):
    """Shortest rotation taking +z to unit vector v."""
    z = np.array([0.0, 0, 1])
    ax = np.cross(z, v)
    s, c = np.linalg.norm(ax), float(np.dot(z, v))
    if s < 1e-12:
        return np.array([1.0, 0, 0, 0]) if c > 0 else np.array([0.0, 1, 0, 0])
    ang = np.arctan2(s, c)
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax / s])


def load(path):
    return Parser(path).parse()
