"""Batched mirror of `earl_benchmark/wrappers/persistent_state_wrapper.py:8-48`.

The counters and the reset-free horizon live ON THE DEVICE, fused into the step kernel
(csrc/tabletop_kernels.cuh); this class only configures the env and exposes the reference surface.
"""


class PersistentStateWrapper:

    def __init__(self, env, episode_horizon):
        self.env = env
        self._episode_horizon = int(episode_horizon)
        env._configure(episode_horizon=self._episode_horizon)

    def reset(self, **kwargs):
        # num_interventions += 1 and steps_since_reset = 0 happen in the reset kernel (reference :17-20)
        return self.env.reset(**kwargs)

    def step(self, action, **kwargs):
        # total_step_count / steps_since_reset / horizon `done` happen in the step kernel (reference :22-31)
        return self.env.step(action, **kwargs)

    def is_successful(self, obs=None):
        if hasattr(self.env, "is_successful"):
            return self.env.is_successful(obs)
        return False

    @property
    def num_interventions(self):
        """int64 tensor [N] (reference returns the single env's int)."""
        return self.env._counters()[1]

    @property
    def total_steps(self):
        """Python int: every env of the batch has taken the same number of steps."""
        return self.env._counters()[0]

    @property
    def steps_since_reset(self):
        return self.env._counters()[2]

    def __getattr__(self, name):
        if name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)
