"""MJCF Spec -> structure-of-arrays model for the articulated-body engine.

Mirrors what MuJoCo's compiler does for the features the EARL Sawyer / Franka scenes use (SURVEY.md
Appendix B/C): body inertias (explicit <inertial>, else from geoms inside `inertiagrouprange`), legacy mesh
centring, `body_invweight0` / `dof_invweight0` at qpos0 (the scale of every soft-constraint regulariser), and
then -- unlike MuJoCo -- fuses every joint-less body into its nearest moving ancestor, so the device tree has
only bodies that move (Sawyer door scene: 36 bodies -> 11).  Fusion is physics-preserving; constraint scales
that MuJoCo derives from the ORIGINAL bodies are computed before fusing and carried along.
"""
import numpy as np

from . import mesh as meshlib
from .parser import GEOM_TYPES, _floats, quat2mat, quat_mul, mat2quat

MJMINVAL = 1e-15
MASSLESS_IPOS_FROM_POS = True
JOINT_TYPES = {"free": 0, "ball": 1, "slide": 2, "hinge": 3}


def _bool(s):
    return str(s).lower() == "true"


class Transform:
    """Rigid transform carried as (pos, unit quaternion).  Composition multiplies the quaternions WITHOUT
    canonicalising their sign: MuJoCo builds body orientations the same way (xquat = xquat[parent] * body_quat),
    and the sign is observable -- the weld residual is the vector part of a quaternion product, so q and -q pull a
    body with a large orientation error the opposite way round."""

    def __init__(self, pos=None, quat=None):
        self.pos = np.zeros(3) if pos is None else np.asarray(pos, float)
        self.quat = np.array([1.0, 0, 0, 0]) if quat is None else np.asarray(quat, float)
        self.R = quat2mat(self.quat)

    def __matmul__(self, o):  # self o other:  x -> self.R (o.R x + o.pos) + self.pos
        return Transform(self.R @ o.pos + self.pos, quat_mul(self.quat, o.quat))

    def apply(self, p):
        return self.R @ np.asarray(p, float) + self.pos


def geom_mass_inertia(g, meshes):
    """(mass, com, inertia about com) of one geom in ITS OWN frame."""
    t = g["type"]
    size = _floats(g.get("size", "0"), 3)
    dens = float(g["density"])
    if t == "box":
        a, b, c = size
        vol = 8 * a * b * c
        I = np.diag([b * b + c * c, a * a + c * c, a * a + b * b]) / 3.0
    elif t == "sphere":
        r = size[0]
        vol = 4.0 / 3.0 * np.pi * r ** 3
        I = np.eye(3) * 0.4 * r * r
    elif t == "cylinder":
        r, h = size[0], size[1]
        vol = np.pi * r * r * 2 * h
        I = np.diag([(3 * r * r + 4 * h * h) / 12.0] * 2 + [r * r / 2.0])
    elif t == "capsule":
        r, h = size[0], size[1]
        vc, vs = np.pi * r * r * 2 * h, 4.0 / 3.0 * np.pi * r ** 3
        vol = vc + vs
        ixx = vc * (3 * r * r + 4 * h * h) / 12.0 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
        izz = vc * r * r / 2.0 + vs * 0.4 * r * r
        I = np.diag([ixx, ixx, izz]) / vol
    elif t == "mesh":
        m = meshes[g["mesh"]]
        mass, I = meshlib.legacy_inertia(m["tris"], m["center"], dens)
        if "mass" in g:
            I = I * float(g["mass"]) / mass
            mass = float(g["mass"])
        return mass, np.zeros(3), I
    else:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    mass = float(g["mass"]) if "mass" in g else dens * vol
    return mass, np.zeros(3), I * mass


class RawModel:
    """Un-fused model (MuJoCo's own body structure), with numpy kinematics for compile-time quantities."""

    def __init__(self, spec, body_pos_overrides=None):
        self.spec = spec
        self.opt = spec.option
        lo, hi = [int(x) for x in spec.compiler["inertiagrouprange"].split()]
        infer = spec.compiler["inertiafromgeom"]
        self.meshes = {}
        for name, m in spec.meshes.items():
            self.meshes[name] = dict(m)
        nb = len(spec.bodies)
        self.nbody = nb
        self.names = [b["name"] for b in spec.bodies]
        self.parent = np.array([b["parent"] for b in spec.bodies])
        self.pos = np.array([b["pos"] for b in spec.bodies], float)
        for k, v in (body_pos_overrides or {}).items():
            self.pos[self.names.index(k)] = np.asarray(v, float)
        self.quat = np.array([b["quat"] for b in spec.bodies], float)
        self.mocap = np.array([b["mocap"] for b in spec.bodies])
        self.mass = np.zeros(nb)
        self.ipos = np.zeros((nb, 3))
        self.inertia = np.zeros((nb, 3, 3))   # about the CoM, body-frame axes
        self.joints, self.geoms, self.sites = [], [], []
        nq = nv = 0
        for bi, b in enumerate(spec.bodies):
            for g in b["geoms"]:
                g = dict(g)
                g["body"] = bi
                if g["type"] == "mesh":
                    self._load_mesh(g["mesh"])
                g["lpos"], g["lquat"] = self._geom_frame(g)
                self.geoms.append(g)
            for s in b["sites"]:
                s = dict(s)
                s["body"] = bi
                s["lpos"], s["lquat"] = _floats(s["pos"], 3), s["quat_resolved"]
                self.sites.append(s)
            for j in b["joints"]:
                j = dict(j)
                j["body"] = bi
                j["jtype"] = JOINT_TYPES[j["type"]]
                j["qposadr"], j["dofadr"] = nq, nv
                dq, dv = {0: (7, 6), 1: (4, 3), 2: (1, 1), 3: (1, 1)}[j["jtype"]]
                nq, nv = nq + dq, nv + dv
                j["jpos"] = _floats(j["pos"], 3)
                ax = _floats(j["axis"], 3)
                j["jaxis"] = ax / max(np.linalg.norm(ax), 1e-300)
                self.joints.append(j)
            # inertial
            if b["inertial"] is not None and infer != "true":
                it = b["inertial"]
                self.mass[bi], self.ipos[bi] = it["mass"], it["pos"]
                if it["full"] is not None:
                    self.inertia[bi] = it["full"]
                else:
                    R = quat2mat(it["quat"])
                    self.inertia[bi] = R @ np.diag(it["diag"]) @ R.T
            elif infer != "false" and bi > 0:
                parts = []
                for g in self.geoms:
                    if g["body"] == bi and lo <= int(g["group"]) <= hi:
                        m, c, I = geom_mass_inertia(g, self.meshes)
                        if m > 0:
                            R = quat2mat(g["lquat"])
                            parts.append((m, g["lpos"] + R @ c, R @ I @ R.T))
                if parts:
                    M = sum(p[0] for p in parts)
                    com = sum(p[0] * p[1] for p in parts) / M
                    I = np.zeros((3, 3))
                    for m, c, Ig in parts:
                        d = c - com
                        I += Ig + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                    self.mass[bi], self.ipos[bi], self.inertia[bi] = M, com, I
                elif MASSLESS_IPOS_FROM_POS:
                    # MuJoCo's compiler (mjCBody::Compile): a body whose inertial frame is still undefined after the geom
                    # pass gets "ipos undefined: copy body frame into inertial" -- with local coordinates that puts the
                    # PARENT-relative offset `pos` into the BODY-relative slot `ipos`.  Massless, so no dynamics change,
                    # but mj_jacBodyCom (hence body_invweight0, hence the weld / contact regularisers) is taken there.
                    self.ipos[bi] = self.pos[bi]
        self.nq, self.nv = nq, nv
        self.qpos0 = np.zeros(nq)
        for j in self.joints:
            if j["jtype"] == 0:
                b = j["body"]
                self.qpos0[j["qposadr"]:j["qposadr"] + 3] = self.pos[b]
                self.qpos0[j["qposadr"] + 3:j["qposadr"] + 7] = self.quat[b]
            elif j["jtype"] in (2, 3):
                self.qpos0[j["qposadr"]] = float(j["ref"])
        self.body_joints = [[j for j in self.joints if j["body"] == b] for b in range(nb)]
        # weld id: bodies without joints up the whole chain are welded to the world
        self.moving = np.zeros(nb, bool)
        for b in range(1, nb):
            self.moving[b] = bool(self.body_joints[b]) or self.moving[self.parent[b]]

    def _load_mesh(self, name):
        m = self.meshes[name]
        if "tris" not in m:
            m["tris"] = meshlib.load_stl(m["file"], m["scale"])
            m["center"] = meshlib.legacy_center(m["tris"])

    def _geom_frame(self, g):
        pos, quat = _floats(g.get("pos", "0 0 0"), 3), g["quat_resolved"]
        if "fromto" in g:
            ft = _floats(g["fromto"])
            pos = 0.5 * (ft[:3] + ft[3:])
            d = ft[3:] - ft[:3]
            from .parser import quat_z2vec
            quat = quat_z2vec(d / np.linalg.norm(d))
            g["size"] = f"{_floats(g['size'])[0]} {0.5 * np.linalg.norm(d)}"
        if g["type"] == "mesh":
            # the geom frame sits at the mesh's (legacy) centre; axes kept as authored (MuJoCo also rotates them to the
            # principal axes, which changes no position and no physics)
            pos = pos + quat2mat(quat) @ self.meshes[g["mesh"]]["center"]
        return pos, quat

    # ---------------------------------------------------------------- kinematics on the raw tree
    def fk(self, qpos, mocap_pos=None, mocap_quat=None):
        """World poses of all bodies: xpos [nb,3], xmat [nb,3,3]; also per-dof (anchor, axis, type) in the world."""
        nb = self.nbody
        xpos, xmat = np.zeros((nb, 3)), np.zeros((nb, 3, 3))
        xmat[0] = np.eye(3)
        dofs = []
        for b in range(1, nb):
            p = self.parent[b]
            if self.mocap[b] and mocap_pos is not None:
                pos, R = np.asarray(mocap_pos, float), quat2mat(np.asarray(mocap_quat, float))
            else:
                pos, R = xpos[p] + xmat[p] @ self.pos[b], xmat[p] @ quat2mat(self.quat[b])
            for j in self.body_joints[b]:
                qa = j["qposadr"]
                if j["jtype"] == 0:
                    pos, R = qpos[qa:qa + 3].copy(), quat2mat(qpos[qa + 3:qa + 7])
                    for k in range(3):
                        dofs.append((b, "t", pos.copy(), np.eye(3)[k]))
                    for k in range(3):
                        dofs.append((b, "r", pos.copy(), R[:, k].copy()))
                elif j["jtype"] == 3:
                    anchor, axis = pos + R @ j["jpos"], R @ j["jaxis"]
                    ang = qpos[qa] - float(j["ref"])
                    Rj = quat2mat(np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * j["jaxis"]]))
                    R = R @ Rj
                    pos = anchor - R @ j["jpos"]
                    dofs.append((b, "r", anchor, axis))
                elif j["jtype"] == 2:
                    axis = R @ j["jaxis"]
                    pos = pos + axis * (qpos[qa] - float(j["ref"]))
                    dofs.append((b, "t", pos + R @ j["jpos"], axis))
                else:
                    raise NotImplementedError("ball joints")
            xpos[b], xmat[b] = pos, R
        return xpos, xmat, dofs

    def _ancestors(self, b):
        out = set()
        while b > 0:
            out.add(b)
            b = self.parent[b]
        return out

    def jac(self, b, point, dofs):
        """Translational and rotational Jacobians (3 x nv each) of body b at world `point`."""
        Jp, Jr = np.zeros((3, self.nv)), np.zeros((3, self.nv))
        anc = self._ancestors(b)
        for d, (db, kind, anchor, axis) in enumerate(dofs):
            if db in anc:
                if kind == "t":
                    Jp[:, d] = axis
                else:
                    Jr[:, d] = axis
                    Jp[:, d] = np.cross(axis, point - anchor)
        return Jp, Jr

    def mass_matrix(self, qpos):
        xpos, xmat, dofs = self.fk(qpos)
        M = np.zeros((self.nv, self.nv))
        for b in range(1, self.nbody):
            if self.mass[b] > 0 and self.moving[b]:
                com = xpos[b] + xmat[b] @ self.ipos[b]
                Jp, Jr = self.jac(b, com, dofs)
                Iw = xmat[b] @ self.inertia[b] @ xmat[b].T
                M += self.mass[b] * Jp.T @ Jp + Jr.T @ Iw @ Jr
        for j in self.joints:
            n = {0: 6, 1: 3, 2: 1, 3: 1}[j["jtype"]]
            for k in range(n):
                M[j["dofadr"] + k, j["dofadr"] + k] += float(j["armature"])
        return M, (xpos, xmat, dofs)

    def invweight0(self):
        """body_invweight0 [nb,2] and dof_invweight0 [nv] at qpos0 (MuJoCo engine_setconst.c: set0)."""
        M, (xpos, xmat, dofs) = self.mass_matrix(self.qpos0)
        Minv = np.linalg.inv(M)
        biw = np.zeros((self.nbody, 2))
        for b in range(1, self.nbody):
            if self.moving[b]:
                com = xpos[b] + xmat[b] @ self.ipos[b]
                Jp, Jr = self.jac(b, com, dofs)
                biw[b, 0] = max(np.trace(Jp @ Minv @ Jp.T) / 3.0, MJMINVAL)
                biw[b, 1] = max(np.trace(Jr @ Minv @ Jr.T) / 3.0, MJMINVAL)
        diw = np.zeros(self.nv)
        for j in self.joints:
            a = j["dofadr"]
            if j["jtype"] in (2, 3):
                diw[a] = Minv[a, a]
            elif j["jtype"] == 0:
                diw[a:a + 3] = np.trace(Minv[a:a + 3, a:a + 3]) / 3.0
                diw[a + 3:a + 6] = np.trace(Minv[a + 3:a + 6, a + 3:a + 6]) / 3.0
        return biw, diw


# ------------------------------------------------------------------------------------------------ fused model

class Model:
    """Fused structure-of-arrays model.  Every field is a numpy array (see `FIELDS`), so it serialises to a flat
    blob every consumer (the CUDA library, test checkers) reads identically (save/load)."""

    FIELDS = (
        # scalars
        ("nbody", "i4"), ("nq", "i4"), ("nv", "i4"), ("ngeom", "i4"), ("nsite", "i4"), ("nu", "i4"), ("nweld", "i4"),
        ("nhullvert", "i4"), ("iterations", "i4"), ("cone_elliptic", "i4"),
        ("timestep", "f8"), ("tolerance", "f8"), ("impratio", "f8"), ("gravity", "f8"),
        # bodies (index 0 = world)
        ("body_parent", "i4"), ("body_pos", "f8"), ("body_quat", "f8"), ("body_mass", "f8"), ("body_ipos", "f8"),
        ("body_inertia", "f8"),           # [nb,6]: xx yy zz xy xz yz about the CoM, body axes
        ("body_jnt", "i4"),               # joint index of the body (one joint per moving body), -1 for world
        # joints
        ("jnt_type", "i4"), ("jnt_body", "i4"), ("jnt_qposadr", "i4"), ("jnt_dofadr", "i4"), ("jnt_pos", "f8"), ("jnt_axis", "f8"),
        ("jnt_limited", "i4"), ("jnt_range", "f8"), ("jnt_margin", "f8"), ("jnt_solref", "f8"), ("jnt_solimp", "f8"),
        ("jnt_stiffness", "f8"), ("jnt_springref", "f8"),
        # dofs
        ("dof_body", "i4"), ("dof_damping", "f8"), ("dof_armature", "f8"), ("dof_frictionloss", "f8"), ("dof_invweight0", "f8"),
        ("qpos0", "f8"),
        # geoms
        ("geom_body", "i4"), ("geom_type", "i4"), ("geom_size", "f8"), ("geom_pos", "f8"), ("geom_quat", "f8"),
        ("geom_contype", "i4"), ("geom_conaffinity", "i4"), ("geom_condim", "i4"), ("geom_priority", "i4"),
        ("geom_friction", "f8"), ("geom_margin", "f8"), ("geom_gap", "f8"), ("geom_solref", "f8"), ("geom_solimp", "f8"),
        ("geom_solmix", "f8"), ("geom_invweight0", "f8"), ("geom_rbound", "f8"), ("geom_hulladr", "i4"), ("geom_hullnum", "i4"),
        ("geom_srcbody", "i4"),           # index of the ORIGINAL body (parent-child collision filter uses original bodies)
        ("geom_srcparent", "i4"),
        ("hull_vert", "f8"),
        # sites
        ("site_body", "i4"), ("site_pos", "f8"), ("site_quat", "f8"),
        # actuators (position servos on a joint)
        ("act_dof", "i4"), ("act_qposadr", "i4"), ("act_kp", "f8"), ("act_ctrlrange", "f8"), ("act_ctrllimited", "i4"),
        ("act_forcerange", "f8"), ("act_forcelimited", "i4"),
        # welds (body1 = mocap frame, body2 = frame on a fused body)
        ("weld_body", "i4"), ("weld_pos", "f8"), ("weld_quat", "f8"), ("weld_relpose", "f8"), ("weld_solref", "f8"),
        ("weld_solimp", "f8"), ("weld_invweight", "f8"),
        ("mocap_pos0", "f8"), ("mocap_quat0", "f8"),
    )
    # blob version 2 (written only when a model has joint equalities; the kitchen): appended after FIELDS
    EXT_FIELDS = (
        ("neq", "i4"),
        ("eq_qposadr", "i4"), ("eq_dofadr", "i4"),   # [neq,2]: joint1, joint2 (hinge / slide)
        ("eq_polycoef", "f8"),                       # [neq,5]: q1 - q1_0 = poly(q2 - q2_0)
        ("eq_solref", "f8"), ("eq_solimp", "f8"), ("eq_invweight", "f8"),
        ("dof_solref_friction", "f8"), ("dof_solimp_friction", "f8"),   # [nv,2], [nv,5]: friction-loss rows
    )

    def __init__(self):
        self.names = {"body": [], "joint": [], "geom": [], "site": []}

    def _ext_defaults(self):
        nv = int(self.nv)
        d = dict(neq=np.int32(0), eq_qposadr=np.zeros((0, 2), np.int32), eq_dofadr=np.zeros((0, 2), np.int32),
                 eq_polycoef=np.zeros((0, 5)), eq_solref=np.zeros((0, 2)), eq_solimp=np.zeros((0, 5)), eq_invweight=np.zeros(0),
                 dof_solref_friction=np.tile(np.array([0.02, 1.0]), (nv, 1)),
                 dof_solimp_friction=np.tile(np.array([0.9, 0.95, 0.001, 0.5, 2.0]), (nv, 1)))
        for k, v in d.items():
            if not hasattr(self, k):
                setattr(self, k, v)

    def save(self, path):
        self._ext_defaults()
        np.savez_compressed(path, **{k: getattr(self, k) for k, _ in self.FIELDS + self.EXT_FIELDS},
                            **{f"name_{k}": np.array(v) for k, v in self.names.items()})

    @classmethod
    def load(cls, path):
        m = cls()
        with np.load(path, allow_pickle=False) as z:
            for k, _ in cls.FIELDS:
                setattr(m, k, z[k])
            for k, _ in cls.EXT_FIELDS:
                if k in z.files:
                    setattr(m, k, z[k])
            for k in m.names:
                m.names[k] = [str(x) for x in z[f"name_{k}"]]
        return m

    def to_blob(self):
        """Flat binary: magic 'EMDL', version, then for each field: i4 ndim, i4 dims..., data (i4 or f8), 8-byte aligned.
        Version 1 = FIELDS (door, peg); version 2 = FIELDS + EXT_FIELDS (models with joint equalities or friction loss)."""
        self._ext_defaults()
        version = 2 if int(self.neq) > 0 or np.any(np.asarray(self.dof_frictionloss) > 0) else 1
        out = [np.array([0x4C444D45, version], dtype="<i4").tobytes()]
        for k, dt in self.FIELDS + (self.EXT_FIELDS if version == 2 else ()):
            a = np.ascontiguousarray(np.asarray(getattr(self, k)), dtype="<" + dt)
            hdr = np.array([a.ndim] + list(a.shape), dtype="<i4").tobytes()
            if len(hdr) % 8:
                hdr += b"\0" * 4
            data = a.tobytes()
            if len(data) % 8:
                data += b"\0" * (8 - len(data) % 8)
            out += [hdr, data]
        return b"".join(out)

    def site_id(self, name):
        return self.names["site"].index(name)

    def geom_id(self, name):
        return self.names["geom"].index(name)


# Scale of the weld's three translational regulariser rows relative to MuJoCo's documented diagApprox
# (body_invweight0[body1] + body_invweight0[body2], translational part).  Round 1 had to CALIBRATE this to 3.35 against the
# shipped demonstrations; the cause was the inertial frame of the massless `hand` body (MASSLESS_IPOS_FROM_POS above): with
# MuJoCo's "ipos <- pos" rule the documented value itself is 2.816 x larger (6.106 instead of 2.168 1/kg for the Sawyer
# hand at qpos0, rotational part unchanged), which is the fitted range (2.9 from the rest pose, 3.35 from the door demos)
# and reproduces the golden hand rest pose of sawyer_door.py:13 to 0.65 mm.  Nothing is calibrated any more.
WELD_TRAN_SCALE = 1.0
WELD_ROT_USES_TRAN_WEIGHT = True


def compile_model(spec, body_pos_overrides=None, keep_geoms=(), keep_sites=None, frame_sites=(), weld_tran_scale=WELD_TRAN_SCALE,
                  weld_relpose="identity"):
    """Spec -> fused Model.

    keep_geoms: names of non-colliding geoms to keep (their frames are observed, e.g. 'handle').
    frame_sites: names of ORIGINAL bodies whose frames are needed (observations, welds): each becomes a site
                 named 'body:<name>' on the fused body.
    weld_relpose: "identity" = metaworld's reset_mocap_welds() (eq_data <- [0 0 0 1 0 0 0]); "qpos0" = MuJoCo's compiler
                 default when the MJCF gives no relpose: pose of body2 in the frame of body1 at qpos0 (kitchen).
    """
    raw = RawModel(spec, body_pos_overrides)
    biw, diw = raw.invweight0()
    nb = raw.nbody
    # anchor (nearest moving-joint ancestor-or-self, else world) and transform to it
    anchor = np.zeros(nb, int)
    T = [Transform() for _ in range(nb)]
    for b in range(1, nb):
        local = Transform(raw.pos[b], raw.quat[b])
        if raw.body_joints[b]:
            anchor[b] = b
        elif raw.mocap[b]:
            anchor[b] = -1
        else:
            p = raw.parent[b]
            anchor[b] = anchor[p]
            T[b] = T[p] @ local
    moving = [b for b in range(1, nb) if anchor[b] == b]
    fid = {0: 0}
    for k, b in enumerate(moving):
        fid[b] = k + 1
    m = Model()
    nf = len(moving) + 1
    m.nbody = np.int32(nf)
    m.names["body"] = ["world"] + [raw.names[b] for b in moving]
    m.body_parent = np.zeros(nf, np.int32)
    m.body_pos, m.body_quat = np.zeros((nf, 3)), np.tile([1.0, 0, 0, 0], (nf, 1))
    m.body_mass, m.body_ipos, m.body_inertia = np.zeros(nf), np.zeros((nf, 3)), np.zeros((nf, 6))
    m.body_jnt = -np.ones(nf, np.int32)
    m.body_parent[0] = -1
    for b in moving:
        p = raw.parent[b]
        pa = anchor[p]
        m.body_parent[fid[b]] = fid[pa]
        loc = T[p] @ Transform(raw.pos[b], raw.quat[b])
        m.body_pos[fid[b]], m.body_quat[fid[b]] = loc.pos, loc.quat
    # composite inertias
    acc = {f: [] for f in range(nf)}
    for b in range(1, nb):
        if raw.mass[b] > 0 and anchor[b] > 0:
            acc[fid[anchor[b]]].append((raw.mass[b], T[b].apply(raw.ipos[b]), T[b].R @ raw.inertia[b] @ T[b].R.T))
    for f, parts in acc.items():
        if not parts:
            continue
        M = sum(p[0] for p in parts)
        com = sum(p[0] * p[1] for p in parts) / M
        I = np.zeros((3, 3))
        for ms, c, Ig in parts:
            d = c - com
            I += Ig + ms * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        m.body_mass[f], m.body_ipos[f] = M, com
        m.body_inertia[f] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
    # joints / dofs
    nj = len(raw.joints)
    m.nq, m.nv = np.int32(raw.nq), np.int32(raw.nv)
    m.names["joint"] = [j["name"] for j in raw.joints]
    m.jnt_type = np.array([j["jtype"] for j in raw.joints], np.int32)
    m.jnt_body = np.array([fid[j["body"]] for j in raw.joints], np.int32)
    m.jnt_qposadr = np.array([j["qposadr"] for j in raw.joints], np.int32)
    m.jnt_dofadr = np.array([j["dofadr"] for j in raw.joints], np.int32)
    m.jnt_pos = np.array([j["jpos"] for j in raw.joints]).reshape(nj, 3)
    m.jnt_axis = np.array([j["jaxis"] for j in raw.joints]).reshape(nj, 3)
    m.jnt_limited = np.array([_bool(j["limited"]) for j in raw.joints], np.int32)
    m.jnt_range = np.array([_floats(j["range"], 2) for j in raw.joints]).reshape(nj, 2)
    m.jnt_margin = np.array([float(j["margin"]) for j in raw.joints])
    m.jnt_solref = np.array([_floats(j["solreflimit"], 2) for j in raw.joints]).reshape(nj, 2)
    m.jnt_solimp = np.array([_solimp(j["solimplimit"]) for j in raw.joints]).reshape(nj, 5)
    m.jnt_stiffness = np.array([float(j["stiffness"]) for j in raw.joints])
    m.jnt_springref = np.array([float(j["springref"]) for j in raw.joints])
    for k, j in enumerate(raw.joints):
        if m.body_jnt[fid[j["body"]]] >= 0:
            raise NotImplementedError(f"body {raw.names[j['body']]} has more than one joint")
        m.body_jnt[fid[j["body"]]] = k
    m.dof_body = np.zeros(raw.nv, np.int32)
    m.dof_damping, m.dof_armature, m.dof_frictionloss = np.zeros(raw.nv), np.zeros(raw.nv), np.zeros(raw.nv)
    for j in raw.joints:
        n = {0: 6, 1: 3, 2: 1, 3: 1}[j["jtype"]]
        s = slice(j["dofadr"], j["dofadr"] + n)
        m.dof_body[s] = fid[j["body"]]
        m.dof_damping[s], m.dof_armature[s] = float(j["damping"]), float(j["armature"])
        m.dof_frictionloss[s] = float(j["frictionloss"])
    m.dof_invweight0 = diw
    m.qpos0 = raw.qpos0
    # geoms
    keep = []
    for gi, g in enumerate(raw.geoms):
        ct, ca = int(g["contype"]), int(g["conaffinity"])
        if (ct or ca) or g.get("name") in keep_geoms:
            keep.append(gi)
    ng = len(keep)
    m.ngeom = np.int32(ng)
    m.names["geom"] = [raw.geoms[gi].get("name", f"geom{gi}") for gi in keep]
    hull, hadr, hnum = [], [], []
    arr = lambda f, dt=float: np.array([f(raw.geoms[gi]) for gi in keep], dtype=dt)  # noqa: E731
    m.geom_body = arr(lambda g: fid[anchor[g["body"]]] if anchor[g["body"]] >= 0 else 0, np.int32)
    m.geom_srcbody = arr(lambda g: g["body"], np.int32)
    m.geom_srcparent = arr(lambda g: raw.parent[g["body"]], np.int32)
    m.geom_type = arr(lambda g: GEOM_TYPES.index(g["type"]), np.int32)
    m.geom_size = arr(lambda g: _floats(g.get("size", "0"), 3)[:3]).reshape(ng, 3)
    m.geom_pos = arr(lambda g: T[g["body"]].apply(g["lpos"])).reshape(ng, 3)
    m.geom_quat = arr(lambda g: quat_mul(T[g["body"]].quat, g["lquat"])).reshape(ng, 4)
    m.geom_contype = arr(lambda g: int(g["contype"]), np.int32)
    m.geom_conaffinity = arr(lambda g: int(g["conaffinity"]), np.int32)
    m.geom_condim = arr(lambda g: int(g["condim"]), np.int32)
    m.geom_priority = arr(lambda g: int(g["priority"]), np.int32)
    m.geom_friction = arr(lambda g: _floats(g["friction"], 3)).reshape(ng, 3)
    m.geom_margin = arr(lambda g: float(g["margin"]))
    m.geom_gap = arr(lambda g: float(g["gap"]))
    m.geom_solref = arr(lambda g: _floats(g["solref"], 2)).reshape(ng, 2)
    m.geom_solimp = arr(lambda g: _solimp(g["solimp"])).reshape(ng, 5)
    m.geom_solmix = arr(lambda g: float(g["solmix"]))
    m.geom_invweight0 = arr(lambda g: biw[g["body"]]).reshape(ng, 2)
    rb = []
    for gi in keep:
        g = raw.geoms[gi]
        sz = _floats(g.get("size", "0"), 3)
        if g["type"] == "mesh":
            hv = meshlib.convex_hull_vertices(raw.meshes[g["mesh"]]["tris"]) - raw.meshes[g["mesh"]]["center"]
            hadr.append(sum(len(h) for h in hull))
            hnum.append(len(hv))
            hull.append(hv)
            rb.append(np.linalg.norm(hv, axis=1).max())
        else:
            hadr.append(0)
            hnum.append(0)
            rb.append({"sphere": sz[0], "capsule": sz[0] + sz[1], "cylinder": np.hypot(sz[0], sz[1]),
                       "box": np.linalg.norm(sz), "plane": 0.0}.get(g["type"], 0.0))
    m.geom_rbound = np.array(rb)
    m.geom_hulladr, m.geom_hullnum = np.array(hadr, np.int32), np.array(hnum, np.int32)
    m.hull_vert = np.concatenate(hull) if hull else np.zeros((0, 3))
    m.nhullvert = np.int32(len(m.hull_vert))
    # sites (+ frames of requested original bodies)
    sites = [s for s in raw.sites if keep_sites is None or s.get("name") in keep_sites]
    recs = [(s.get("name", "site"), anchor[s["body"]], T[s["body"]] @ Transform(s["lpos"], s["lquat"])) for s in sites]
    for bname in frame_sites:
        b = raw.names.index(bname)
        recs.append((f"body:{bname}", anchor[b], T[b] if anchor[b] != b else Transform()))
    m.nsite = np.int32(len(recs))
    m.names["site"] = [r[0] for r in recs]
    m.site_body = np.array([fid[r[1]] if r[1] >= 0 else 0 for r in recs], np.int32)
    m.site_pos = np.array([r[2].pos for r in recs]).reshape(len(recs), 3)
    m.site_quat = np.array([r[2].quat for r in recs]).reshape(len(recs), 4)
    # actuators
    acts = spec.actuators
    m.nu = np.int32(len(acts))
    jn = m.names["joint"]
    m.act_dof = np.array([raw.joints[jn.index(a["joint"])]["dofadr"] for a in acts], np.int32)
    m.act_qposadr = np.array([raw.joints[jn.index(a["joint"])]["qposadr"] for a in acts], np.int32)
    for a in acts:
        if a["tag"] != "position":
            raise NotImplementedError(f"actuator <{a['tag']}>")
    m.act_kp = np.array([float(a["kp"]) for a in acts])
    m.act_ctrlrange = np.array([_floats(a["ctrlrange"], 2) for a in acts]).reshape(len(acts), 2)
    m.act_ctrllimited = np.array([_bool(a["ctrllimited"]) for a in acts], np.int32)
    m.act_forcerange = np.array([_floats(a["forcerange"], 2) for a in acts]).reshape(len(acts), 2)
    m.act_forcelimited = np.array([_bool(a["forcelimited"]) for a in acts], np.int32)
    # welds: body1 must be the mocap body (static), body2 a frame on a fused body
    welds = [e for e in spec.equalities if e["tag"] == "weld" and _bool(e.get("active", "true"))]
    m.nweld = np.int32(len(welds))
    wb, wp, wq, wr, wsr, wsi, wiw = [], [], [], [], [], [], []
    for e in welds:
        b1, b2 = raw.names.index(e["body1"]), raw.names.index(e["body2"])
        if not raw.mocap[b1]:
            raise NotImplementedError("weld body1 must be a mocap body")
        wb.append(fid[anchor[b2]])
        t2 = T[b2] if anchor[b2] != b2 else Transform()
        wp.append(t2.pos)
        wq.append(t2.quat)
        if "relpose" in e and np.any(_floats(e["relpose"], 7)[3:] != 0):
            wr.append(_floats(e["relpose"], 7))
        elif weld_relpose == "identity":
            # metaworld's reset_mocap_welds() overwrites eq_data with the identity relative pose
            wr.append(_floats("0 0 0 1 0 0 0", 7))
        else:
            xpos0, xmat0 = raw.fk(raw.qpos0)[:2]
            R1 = xmat0[b1]
            rq = mat2quat(R1.T @ xmat0[b2])
            wr.append(np.concatenate([R1.T @ (xpos0[b2] - xpos0[b1]), rq]))
        wsr.append(_floats(e["solref"], 2))
        wsi.append(_solimp(e["solimp"]))
        tr = (biw[b1] + biw[b2])
        # mj_diagApprox of MuJoCo 2.1.0 gives ALL SIX weld rows the bodies' TRANSLATIONAL inverse weight (the rotational
        # entry of body_invweight0 only enters weld rows in later releases).  Established against the reference's own MuJoCo
        # run: with it the fp64 checker reproduces the golden hand rest poses of sawyer_door.py:13 / sawyer_peg.py:18 to
        # < 5 um and the free-space hand trajectory of all 40 shipped door / peg episodes to < 5 um rms; with the rotational
        # weight (284.9 instead of 6.106 1/kg m^2 for the Sawyer hand) the same quantities are off by 0.65 mm / 10 mm and
        # 7 mm rms (tools/freespace_fit.py; the pin tests of the fp64 checker).
        wiw.append(np.array([tr[0] * weld_tran_scale, tr[0] * weld_tran_scale if WELD_ROT_USES_TRAN_WEIGHT else tr[1]]))
    nw = len(welds)
    m.weld_body = np.array(wb, np.int32)
    m.weld_pos, m.weld_quat = np.array(wp).reshape(nw, 3), np.array(wq).reshape(nw, 4)
    m.weld_relpose = np.array(wr).reshape(nw, 7)
    m.weld_solref, m.weld_solimp = np.array(wsr).reshape(nw, 2), np.array(wsi).reshape(nw, 5)
    m.weld_invweight = np.array(wiw).reshape(nw, 2)
    # joint equalities (mjEQ_JOINT): q1 - q1_0 = poly(q2 - q2_0), hinge / slide joints
    eqs = [e for e in spec.equalities if e["tag"] == "joint" and _bool(e.get("active", "true"))]
    m.neq = np.int32(len(eqs))
    eq_q, eq_d, eq_c, eq_sr, eq_si, eq_iw = [], [], [], [], [], []
    for e in eqs:
        j1, j2 = raw.joints[jn.index(e["joint1"])], raw.joints[jn.index(e["joint2"])]
        if j1["jtype"] < 2 or j2["jtype"] < 2:
            raise NotImplementedError("joint equality on a free / ball joint")
        eq_q.append((j1["qposadr"], j2["qposadr"]))
        eq_d.append((j1["dofadr"], j2["dofadr"]))
        eq_c.append(_floats(e.get("polycoef", "0 1 0 0 0"), 5))
        eq_sr.append(_floats(e["solref"], 2))
        eq_si.append(_solimp(e["solimp"]))
        eq_iw.append(diw[j1["dofadr"]] + diw[j2["dofadr"]])     # mj_diagApprox, mjEQ_JOINT
    ne = len(eqs)
    m.eq_qposadr, m.eq_dofadr = np.array(eq_q, np.int32).reshape(ne, 2), np.array(eq_d, np.int32).reshape(ne, 2)
    m.eq_polycoef, m.eq_solref, m.eq_solimp = np.array(eq_c).reshape(ne, 5), np.array(eq_sr).reshape(ne, 2), np.array(eq_si).reshape(ne, 5)
    m.eq_invweight = np.array(eq_iw).reshape(ne)
    m.dof_solref_friction = np.array([_floats(raw.joints[k]["solreffriction"], 2) for k in range(len(raw.joints)) for _ in range({0: 6, 1: 3, 2: 1, 3: 1}[raw.joints[k]["jtype"]])]).reshape(raw.nv, 2)
    m.dof_solimp_friction = np.array([_solimp(raw.joints[k]["solimpfriction"]) for k in range(len(raw.joints)) for _ in range({0: 6, 1: 3, 2: 1, 3: 1}[raw.joints[k]["jtype"]])]).reshape(raw.nv, 5)
    mc = [b for b in range(nb) if raw.mocap[b]]
    m.mocap_pos0 = raw.pos[mc[0]].copy() if mc else np.zeros(3)
    m.mocap_quat0 = raw.quat[mc[0]].copy() if mc else np.array([1.0, 0, 0, 0])
    # options
    o = spec.option
    m.timestep, m.tolerance, m.impratio = np.float64(o["timestep"]), np.float64(o["tolerance"]), np.float64(o["impratio"])
    m.iterations = np.int32(o["iterations"])
    m.cone_elliptic = np.int32(o["cone"] == "elliptic")
    m.gravity = _floats(o["gravity"], 3)
    m.raw = raw  # not serialised: kept for compile-time cross-checks
    return m


def _solimp(s):
    v = _floats(s)
    full = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    full[:len(v)] = v
    return full


def fk_fused(m, qpos, mocap_pos=None, mocap_quat=None):
    """numpy forward kinematics on the fused model (compile-time cross-checks and tests only)."""
    nb = int(m.nbody)
    xpos, xmat = np.zeros((nb, 3)), np.zeros((nb, 3, 3))
    xmat[0] = np.eye(3)
    for b in range(1, nb):
        p = m.body_parent[b]
        pos, R = xpos[p] + xmat[p] @ m.body_pos[b], xmat[p] @ quat2mat(m.body_quat[b])
        j = m.body_jnt[b]
        qa = m.jnt_qposadr[j]
        if m.jnt_type[j] == 0:
            pos, R = np.array(qpos[qa:qa + 3], float), quat2mat(np.asarray(qpos[qa + 3:qa + 7], float))
        elif m.jnt_type[j] == 3:
            anchor = pos + R @ m.jnt_pos[j]
            ang = qpos[qa] - m.qpos0[qa]
            R = R @ quat2mat(np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * m.jnt_axis[j]]))
            pos = anchor - R @ m.jnt_pos[j]
        elif m.jnt_type[j] == 2:
            pos = pos + (R @ m.jnt_axis[j]) * (qpos[qa] - m.qpos0[qa])
        xpos[b], xmat[b] = pos, R
    geom_xpos = np.array([xpos[m.geom_body[g]] + xmat[m.geom_body[g]] @ m.geom_pos[g] for g in range(int(m.ngeom))]).reshape(-1, 3)
    site_xpos = np.array([xpos[m.site_body[s]] + xmat[m.site_body[s]] @ m.site_pos[s] for s in range(int(m.nsite))]).reshape(-1, 3)
    return xpos, xmat, geom_xpos, site_xpos
