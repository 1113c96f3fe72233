"""ctypes binding of tests/host_emulation/libemul.so: the engine kernel source (csrc/mj_*.cuh) compiled for the host
with one lane.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "earl_benchmark_b200", "csrc")
_LIB = None


class TaskSpec(C.Structure):
    _fields_ = [("frame_skip", C.c_int32), ("hand_site", C.c_int32), ("ree_site", C.c_int32), ("lee_site", C.c_int32),
                ("obj_geom", C.c_int32), ("obj_site", C.c_int32), ("max_newton", C.c_int32), ("obj_qpos_count", C.c_int32),
                ("mocap_low", C.c_float * 3), ("mocap_high", C.c_float * 3), ("action_scale", C.c_float),
                ("success_radius", C.c_float), ("obj_init_pos", C.c_float * 3), ("hand_init_pos", C.c_float * 3),
                ("grasp_site", C.c_int32), ("lpad_site", C.c_int32), ("rpad_site", C.c_int32), ("corner_site", C.c_int32 * 4)]


_LIBS = {}


def lib(capset="small"):
    """capset "small": the door / peg capacities (default build); "kitchen": -DMJ_CAPSET_KITCHEN (23 dofs, 118 geoms)."""
    global _LIB
    if capset != "small":
        if capset not in _LIBS:
            _LIBS[capset] = _load(os.path.join(_HERE, f"libemul_{capset}.so"), ["-DMJ_CAPSET_" + capset.upper()])
        return _LIBS[capset]
    if _LIB is None:
        _LIB = _load(os.path.join(_HERE, "libemul.so"), [])
    return _LIB


def _load(so, defines):
    if True:
        srcs = [os.path.join(_HERE, "emul.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.startswith("mj_")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-ffp-contract=off"] + defines +
                                  ["-o", so, os.path.join(_HERE, "emul.cpp")])
        L = C.CDLL(so)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_char_p, C.c_longlong, C.POINTER(TaskSpec), C.c_char_p, C.c_int]
        L.emu_destroy.argtypes = [C.c_void_p]
        DP = np.ctypeslib.ndpointer(np.float64, flags="C")
        FP = np.ctypeslib.ndpointer(np.float32, flags="C")
        IP = np.ctypeslib.ndpointer(np.int32, flags="C")
        L.emu_set_state.argtypes = [C.c_void_p, DP, DP, DP, DP, DP, DP]
        L.emu_get_state.argtypes = [C.c_void_p, DP, DP, DP, DP]
        L.emu_substeps.argtypes = [C.c_void_p, C.c_int]
        L.emu_env_step.argtypes = [C.c_void_p, FP, FP]
        L.emu_forward_parts.argtypes = [C.c_void_p, DP, DP, DP]
        L.emu_info.argtypes = [C.c_void_p, C.c_int]
        L.emu_contacts.argtypes = [C.c_void_p, DP, DP, DP, IP]
        return L


class Emu:
    def __init__(self, model, task, capset="small"):
        blob = model.to_blob()
        err = C.create_string_buffer(256)
        self.L = lib(capset)
        self.h = self.L.emu_create(blob, len(blob), C.byref(task), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.nq, self.nv, self.nb = int(model.nq), int(model.nv), int(model.nbody)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_destroy(self.h)

    def set_state(self, qpos, qvel, warm, mocap_pos, mocap_quat=(1, 0, 1, 0), ctrl=(0, 0)):
        f = lambda a: np.ascontiguousarray(a, np.float64)  # noqa: E731
        self.L.emu_set_state(self.h, f(qpos), f(qvel), f(warm), f(mocap_pos), f(mocap_quat), f(ctrl))

    def get_state(self):
        q, v, w, mp = np.zeros(self.nq), np.zeros(self.nv), np.zeros(self.nv), np.zeros(3)
        self.L.emu_get_state(self.h, q, v, w, mp)
        return q, v, w, mp

    def substeps(self, n=1):
        self.L.emu_substeps(self.h, int(n))

    def env_step(self, action):
        obs = np.zeros(7, np.float32)
        self.L.emu_env_step(self.h, np.ascontiguousarray(action, np.float32), obs)
        return obs

    def forward_parts(self):
        M, b, x = np.zeros((self.nv, self.nv)), np.zeros(self.nv), np.zeros((self.nb, 3))
        self.L.emu_forward_parts(self.h, M, b, x)
        return M, b, x

    def info(self, what):
        return self.L.emu_info(self.h, {"nefc": 0, "ncon": 1, "iter": 2, "bad": 3, "npair": 4, "nhit": 5}[what])

    def contacts(self):
        n = self.info("ncon")
        d, p, f, g = np.zeros(32), np.zeros((32, 3)), np.zeros((32, 9)), np.zeros((32, 2), np.int32)
        self.L.emu_contacts(self.h, d, p, f, g)
        return d[:n], p[:n], f[:n], g[:n]


def door_task(model, max_newton=0):
    t = TaskSpec()
    t.frame_skip, t.max_newton, t.obj_qpos_count = 5, max_newton, 1
    t.grasp_site = t.lpad_site = t.rpad_site = -1
    t.corner_site[:] = [-1] * 4
    t.hand_site, t.ree_site, t.lee_site = model.site_id("body:hand"), model.site_id("rightEndEffector"), model.site_id("leftEndEffector")
    t.obj_geom, t.obj_site = model.geom_id("handle"), -1
    t.mocap_low[:] = [-0.5, 0.40, 0.05]
    t.mocap_high[:] = [0.5, 1.0, 0.5]
    t.action_scale, t.success_radius = 0.01, 0.02
    return t


def peg_task(model, max_newton=0):
    t = TaskSpec()
    t.frame_skip, t.max_newton, t.obj_qpos_count = 5, max_newton, 3
    t.grasp_site = t.lpad_site = t.rpad_site = -1
    t.corner_site[:] = [-1] * 4
    t.hand_site, t.ree_site, t.lee_site = model.site_id("body:hand"), model.site_id("rightEndEffector"), model.site_id("leftEndEffector")
    t.obj_geom, t.obj_site = -1, model.site_id("pegHead")
    t.mocap_low[:] = [-0.5, 0.40, 0.05]
    t.mocap_high[:] = [0.5, 1.0, 0.5]
    t.action_scale, t.success_radius = 0.01, 0.05
    return t


def kitchen_task(model, max_newton=0):
    """Engine-level task spec for the kitchen model (the kitchen task layer is not in the kernel source yet: only
    `substeps` is meaningful, the Sawyer observation fields point at a valid site)."""
    t = TaskSpec()
    t.frame_skip, t.max_newton, t.obj_qpos_count = 40, max_newton, 0
    t.grasp_site = t.lpad_site = t.rpad_site = -1
    t.corner_site[:] = [-1] * 4
    t.hand_site = t.ree_site = t.lee_site = model.site_id("end_effector")
    t.obj_geom, t.obj_site = -1, model.site_id("slide_site")
    t.mocap_low[:] = [-0.7, -0.1, 1.8]
    t.mocap_high[:] = [0.4, 0.5, 2.6]
    t.action_scale, t.success_radius = 0.01, 0.3
    return t
