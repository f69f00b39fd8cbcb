"""Helpers shared by the golden-vector tests."""
import os
import re

import numpy as np

STATE_KEYS = ("rob", "rhist", "rflag", "ball", "step")


def parse_name(path):
    """GAME_RoboRugbySimpleDuel-v2_chase_s2.npz -> (preset, env_id, kind)."""
    m = re.match(r"(GAME|TRAIN)_(RoboRugby[A-Za-z]*-v\d|DuelAllCoordsPrior|DuelAllCoords|DuelAllMixins|DuelCutChain|DuelLidar6v1|DuelGoals)_([a-z]+)", os.path.basename(path))
    return m.group(1), m.group(2), m.group(3)


# Ad-hoc class compositions (not registered ids) used by some golden files: base id + observer override.
OBS_ALLCOORDS = 3
OBS_ALLCOORDS_PRIOR = 4
OBS_LIDAR6_V1 = 5
CUSTOM = {"DuelAllCoords": ("RoboRugbySimpleDuel-v2", OBS_ALLCOORDS),
          "DuelAllCoordsPrior": ("RoboRugbySimpleDuel-v2", OBS_ALLCOORDS_PRIOR),
          "DuelLidar6v1": ("RoboRugbySimpleDuel-v2", OBS_LIDAR6_V1)}
# ... and reward-mixin compositions, in class-definition order (oracle/ref_harness.py builds exactly these classes)
CUSTOM_MIXINS = {
    "DuelAllMixins": ("RoboRugbySimpleDuel-v2", ["KeepMovingGuys", "DontDriveInGoals", "BaseDestruction", "PushNegBallsFromGoal",
                                                 "PushPosBallsToGoal", "ChasePosBall", "NaughtyBots"]),
    # NaughtyBots.on_step_end does not call super(): KeepMovingGuys, listed after it, never runs its on_step_end
    "DuelCutChain": ("RoboRugbySimpleDuel-v2", ["DontDriveInGoals", "ChasePosBall", "NaughtyBots", "KeepMovingGuys"]),
    "DuelGoals": ("RoboRugbySimpleDuel-v2", ["BaseDestruction", "PushPosBallsToGoal", "ChasePosBall"]),
}


def resolve_env(env_id):
    """(registered id to take the default config from, observer override or None)."""
    if env_id in CUSTOM_MIXINS:
        return CUSTOM_MIXINS[env_id][0], None
    return CUSTOM.get(env_id, (env_id, None))


def resolve_rewards(env_id):
    """(reward_mask, reward_order) override of an ad-hoc mixin composition, or None for registered ids."""
    if env_id not in CUSTOM_MIXINS:
        return None
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from roborugby_b200.constants import reward_config_from_mixins
    return reward_config_from_mixins(CUSTOM_MIXINS[env_id][1])


def apply_overrides(cfg, env_id):
    """Observer / reward-mixin overrides of an ad-hoc composition onto a config struct (oracle or C ABI)."""
    _, observer = resolve_env(env_id)
    if observer is not None:
        cfg.observer = observer
    rw = resolve_rewards(env_id)
    if rw is not None:
        cfg.reward_mask, cfg.reward_order = rw
    return cfg


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
