"""gym.utils.seeding stand-in (TEST INFRASTRUCTURE ONLY). The reference never uses the RNG it returns."""
import numpy as np


def np_random(seed=None):
    if seed is None:
        seed = 0
    return np.random.RandomState(int(seed) % (2 ** 32)), seed
