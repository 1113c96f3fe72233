#!/usr/bin/env python
"""Compile the reference's MJCF assets into structure-of-arrays model files (earl_benchmark_b200/models/*.npz).

Run in the build container (needs /root/reference); the GPU box only sees the committed .npz files.
    python tools/compile_models.py
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from earl_benchmark_b200.mjcf import compile as C, parser  # noqa: E402

REF = os.environ.get("EARL_REFERENCE", "/root/reference")
MW = os.path.join(REF, "earl_benchmark/envs/metaworld_assets/sawyer_xyz")
OUT = os.path.join(REPO, "earl_benchmark_b200", "models")


def sawyer_door():
    spec = parser.load(os.path.join(MW, "sawyer_door_pull.xml"))
    # reset_model() moves the door body to obj_init_pos, an fp32 array (reference envs/sawyer_door.py:36,119-120)
    door_pos = np.array([0.1, 0.95, 0.1], dtype=np.float32).astype(np.float64)
    return C.compile_model(spec, body_pos_overrides={"door": door_pos}, keep_geoms=("handle",), frame_sites=("hand",),
                           keep_sites=("rightEndEffector", "leftEndEffector"))


def sawyer_peg():
    spec = parser.load(os.path.join(MW, "sawyer_peg_insertion_side.xml"))
    # reset_model() moves the block to goal_states[0][4:] - (0.03, 0, 0.13) (reference envs/sawyer_peg.py:195-197)
    box_pos = np.array([-0.3 + 0.03, 0.6, 0.0 + 0.13]) - np.array([0.03, 0.0, 0.13])
    return C.compile_model(spec, body_pos_overrides={"box": box_pos}, frame_sites=("hand", "leftpad", "rightpad"),
                           keep_sites=("rightEndEffector", "leftEndEffector", "pegHead", "pegGrasp",
                                       "bottom_right_corner_collision_box_1", "top_left_corner_collision_box_1",
                                       "bottom_right_corner_collision_box_2", "top_left_corner_collision_box_2"))


def kitchen():
    """Franka kitchen (ENV/kitchen.py -> adept_envs KitchenV0.MODEl).  The weld follows the same derived rule as the
    Sawyer welds (all six rows on the translational body_invweight0, scale 1.0: DESIGN 8.4; round 1's fitted 3.35 is gone)
    and keeps the compiler's default relative pose (kitchen never calls reset_mocap_welds)."""
    spec = parser.load(os.path.join(REF, "earl_benchmark/envs/kitchen_assets/adept_envs/adept_envs/franka/assets",
                                    "franka_kitchen_jntpos_act_ab.xml"))
    return C.compile_model(spec, weld_tran_scale=1.0, weld_relpose="qpos0",
                           keep_sites=("microhandle_site", "hinge_site2", "slide_site", "knob1_site", "knob2_site", "knob3_site",
                                       "knob4_site", "light_site", "end_effector"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("sawyer_door", sawyer_door), ("sawyer_peg", sawyer_peg), ("kitchen", kitchen)):
        m = fn()
        m.save(os.path.join(OUT, name + ".npz"))
        print(name, ": bodies", int(m.nbody), "nq", int(m.nq), "nv", int(m.nv), "geoms", int(m.ngeom), "sites", int(m.nsite),
              "blob", len(m.to_blob()), "B")
