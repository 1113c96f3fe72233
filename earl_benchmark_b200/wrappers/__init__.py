from . import lifelong_wrapper, persistent_state_wrapper  # noqa: F401
