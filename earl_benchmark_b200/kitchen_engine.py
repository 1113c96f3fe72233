"""Engine-level access to the kitchen capacity set (include/earl_mj_kitchen_b200.h): batched `n x mj_step` of the compiled
Franka-kitchen model on caller-held torch tensors.  This is NOT the kitchen task (EARLEnvs('kitchen') is not built): it is
the device engine that task will run on, exposed so it can be checked and measured.
Reference: 40 x sim.step() in kitchen_assets/adept_envs/adept_envs/mujoco_env.py:148-153."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .mjcf.compile import Model

MODEL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models", "kitchen.npz")


class KitchenEngine:
    def __init__(self, device="cuda:0", model_path=MODEL_PATH):
        self.model = Model.load(model_path)
        self.device = torch.device(device)
        blob = self.model.to_blob()
        self._h = C.c_void_p()
        _lib.check(_lib.lib().earl_mjk_engine_create(blob, len(blob), self.device.index or 0, C.byref(self._h)))
        self.nv = int(_lib.lib().earl_mjk_engine_nv(self._h))
        self._quat = np.ascontiguousarray(self.model.mocap_quat0, np.float32)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().earl_mjk_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def substeps(self, qpos, qvel, warm, mocap_pos, ctrl, nsub=1):
        """In place on float32 [N, nv] tensors qpos / qvel / warm; mocap_pos float64 [N,3]; ctrl float32 [N,2].
        Returns info int32 [N,4] = rows, contacts (last substep), Newton iterations (sum), flags."""
        n = qpos.shape[0]
        for t, dt, sh in ((qpos, torch.float32, (n, self.nv)), (qvel, torch.float32, (n, self.nv)), (warm, torch.float32, (n, self.nv)),
                          (mocap_pos, torch.float64, (n, 3)), (ctrl, torch.float32, (n, 2))):
            if t.dtype != dt or tuple(t.shape) != sh or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"expected a contiguous {dt} tensor of shape {sh} on {self.device}")
        info = torch.zeros((n, 4), dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().earl_mjk_engine_substeps(self._h, n, int(nsub), qpos.data_ptr(), qvel.data_ptr(), warm.data_ptr(),
                                                      mocap_pos.data_ptr(), self._quat.ctypes.data, ctrl.data_ptr(),
                                                      info.data_ptr(), C.c_void_p(stream)))
        return info
