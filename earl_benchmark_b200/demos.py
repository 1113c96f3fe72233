"""Shipped demonstrations, re-encoded as .npz package data.

Same arrays and keys as the reference's `demonstrations/<env>/<forward|reverse>/demo_data.pkl`
(reference `earl_benchmark/__init__.py:238-247`): observations, actions, rewards, terminals,
next_observations, infos.  Converted by oracle/gen_golden.py:convert_demos.
"""
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demonstrations")
KEYS = ("observations", "actions", "rewards", "terminals", "next_observations", "infos")


def available(env_name):
    return os.path.exists(os.path.join(_DIR, env_name, "forward.npz"))


def load(env_name, direction):
    path = os.path.join(_DIR, env_name, f"{direction}.npz")
    with np.load(path) as z:
        return {k: z[k] for k in KEYS}


def load_to_device(env_name, direction, device):
    """Same dict with the numeric arrays as tensors on `device` (replay-buffer seeding next to a device-resident env:
    what EARL algorithms do with `get_demonstrations()`).  `infos` stays a host object."""
    import torch
    d = load(env_name, direction)
    return {k: (torch.from_numpy(np.ascontiguousarray(v)).to(device) if k != "infos" else v) for k, v in d.items()}


def episodes(demo):
    """[(start, end)) index ranges of the episodes of a demonstration dict (`terminals` is True exactly on the success
    step that ends each episode)."""
    ends = list(np.nonzero(np.asarray(demo["terminals"]).ravel())[0] + 1)
    return list(zip([0] + ends[:-1], ends))


# ---- state reconstruction from observations (the demonstrations carry 14-d observations, not qpos / qvel) ----------
# sawyer_door: handle(theta) = hinge + Rz(theta) p0 with hinge = obj_init_pos + door_link offset and p0 = the legacy
# mesh centre of door_handle.stl in the door_link frame
_DOOR_HINGE_XY = np.array([-0.085, 0.85])
_DOOR_P0_XY = np.array([0.375721629, -0.107139896])


def door_angle_from_obs(obs):
    """Door hinge angle(s) from observation(s) [..., 14] (inverse of the handle forward kinematics)."""
    h = np.asarray(obs)[..., 4:6]
    return (np.arctan2(h[..., 1] - _DOOR_HINGE_XY[1], h[..., 0] - _DOOR_HINGE_XY[0])
            - np.arctan2(_DOOR_P0_XY[1], _DOOR_P0_XY[0]))


def peg_position_from_obs(obs):
    """Peg body position(s) from observation(s) [..., 14] for a peg lying flat (identity orientation):
    pegHead = peg position + R (-0.1, 0, 0)  (sawyer_peg_insertion_side.xml: site 'pegHead')."""
    return np.asarray(obs)[..., 4:7] + np.array([0.1, 0.0, 0.0])
