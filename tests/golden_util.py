"""Helpers shared by the golden-vector tests."""
import os
import re

import numpy as np

STATE_KEYS = ("rob", "rhist", "rflag", "ball", "step")


def parse_name(path):
    """GAME_RoboRugbySimpleDuel-v2_chase_s2.npz -> (preset, env_id, kind)."""
    m = re.match(r"(GAME|TRAIN)_(RoboRugby[A-Za-z]*-v\d|DuelAllCoords)_([a-z]+)", os.path.basename(path))
    return m.group(1), m.group(2), m.group(3)


# Ad-hoc class compositions (not registered ids) used by some golden files: base id + observer override.
OBS_ALLCOORDS = 3
CUSTOM = {"DuelAllCoords": ("RoboRugbySimpleDuel-v2", OBS_ALLCOORDS)}


def resolve_env(env_id):
    """(registered id to take the default config from, observer override or None)."""
    return CUSTOM.get(env_id, (env_id, None))


def state_at(d, i, t):
    return {k: d[k][i, t] for k in STATE_KEYS}


def actions_at(d, i, t):
    a = d["act"][i, t]
    return a[~np.isnan(a)]


def states_equal(a, b):
    return all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in STATE_KEYS)


def state_diff(a, b):
    out = {}
    for k in STATE_KEYS:
        x, y = np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)
        if not np.array_equal(x, y):
            out[k] = float(np.max(np.abs(x - y)))
    return out
