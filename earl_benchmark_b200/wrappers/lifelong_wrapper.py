"""Batched mirror of `earl_benchmark/wrappers/lifelong_wrapper.py:8-52`.

lifelong_return, steps_since_goal_change and the periodic goal swap run inside the step kernel
(EARL_FLAG_LIFELONG); on goal-change steps the returned observation carries the NEW goal and the
reward is the pre-swap one, as in the reference (:31-42).
"""


class LifelongWrapper:

    def __init__(self, env, goal_change_frequency):
        self.env = env
        self._goal_change_frequency = int(goal_change_frequency)
        base = env
        while hasattr(base, "env"):
            base = base.env
        self._base = base
        base._configure(lifelong=True, goal_change_frequency=self._goal_change_frequency)

    def reset(self, **kwargs):
        # own num_interventions += 1 and steps_since_goal_change = 0 happen in the reset kernel (:25-28)
        return self.env.reset(**kwargs)

    def step(self, action, **kwargs):
        return self.env.step(action, **kwargs)

    @property
    def lifelong_return(self):
        """float64 tensor [N]."""
        return self._base._counters(want_ll=True)[3]

    @property
    def num_interventions(self):
        return self._base._counters()[1]

    def __getattr__(self, name):
        if name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)
