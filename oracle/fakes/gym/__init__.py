"""Minimal stand-in for gym 0.23.1 -- TEST INFRASTRUCTURE ONLY.

Lets the *unmodified* reference package (`/root/reference/earl_benchmark`) import and
run its kinematic tabletop task, both wrappers and the `EARLEnvs` loader in a container
that has neither gym nor MuJoCo.  Valid for tabletop because that task never calls
`sim.step()`: `sim.forward()` does not alter `qpos`, and nothing else is read
(reference `earl_benchmark/envs/tabletop_manipulation.py:55-60,128-174`).

Only `oracle/gen_golden.py` and the optional live-reference tests put this directory on
`sys.path`; nothing in the product package imports it.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def seed(self, seed=None):
        return [seed]


class Wrapper(Env):
    """gym.Wrapper: holds `env`, forwards unknown attributes (gym/core.py, 0.23.1)."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(f"attempted to get missing private attribute '{name}'")
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)
