"""gym 0.23.1 `gym.utils.seeding.np_random` restated (test infrastructure only; gym is pinned in the reference's
setup.py:8 / env.yml:10 and absent from this image).  gym/utils/seeding.py, 0.23.1:

    seed_seq = np.random.SeedSequence(seed); np_seed = seed_seq.entropy
    rng = RandomNumberGenerator(np.random.PCG64(seed_seq))      # a np.random.Generator subclass

so `env.np_random.uniform(...)` is numpy's Generator.uniform on a PCG64 stream seeded through SeedSequence."""
import numpy as np


class RandomNumberGenerator(np.random.Generator):
    pass


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, int) and 0 <= seed):
        raise ValueError(f"Seed must be a non-negative integer or omitted, not {seed}")
    seed_seq = np.random.SeedSequence(seed)
    return RandomNumberGenerator(np.random.PCG64(seed_seq)), seed_seq.entropy
