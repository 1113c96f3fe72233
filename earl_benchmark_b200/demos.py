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
