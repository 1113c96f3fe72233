"""gym.envs.registration stand-in (test infrastructure only): the reference's kitchen package registers env ids at
import time (kitchen_assets/adept_envs/adept_envs/franka/__init__.py:17-37) and never looks them up on this path."""


class _Registry:
    def __init__(self):
        self.env_specs = {}


registry = _Registry()


def register(id, **kwargs):  # noqa: A002
    registry.env_specs[id] = kwargs
