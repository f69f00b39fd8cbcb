"""The two constant sets of the reference (robo_rugby/gym_env/RR_Constants.py:4-34).

The reference selects them with one import-time flag, GAME_MODE; here they are two named presets
that can coexist in one process.  Kernels are instantiated for these entity counts.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Preset:
    name: str
    index: int             # RR_PRESET_* in include/rr_b200.h
    arena_width: int       # ARENA_WIDTH   :6
    arena_height: int      # ARENA_HEIGHT  :7
    num_robots_happy: int  # :32
    num_robots_grumpy: int  # :33
    num_ball_pos: int      # :30
    num_ball_neg: int      # :31
    game_length_steps: int  # :25
    framerate: int = 30    # :23

    @property
    def num_robots_total(self):
        return self.num_robots_happy + self.num_robots_grumpy

    @property
    def num_balls_total(self):
        return self.num_ball_pos + self.num_ball_neg


GAME = Preset("GAME", 0, 800, 800, 2, 2, 4, 4, 4500)
TRAIN = Preset("TRAIN", 1, 600, 600, 1, 0, 1, 0, 300)
PRESETS = {"GAME": GAME, "TRAIN": TRAIN}

TEAM_HAPPY, TEAM_GRUMPY = 1, -1  # :52-53
MOVES_PER_FRAME = 12             # :12-13

ENV_IDS = ("RoboRugby-v0", "RoboRugbySimple-v0", "RoboRugbySimpleDuel-v2", "RoboRugbySimpleDuel-v3")


def get_preset(p):
    if isinstance(p, Preset):
        return p
    return PRESETS[str(p).upper()]
