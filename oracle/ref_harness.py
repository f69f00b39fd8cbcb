"""Harness that imports the UNMODIFIED reference (read-only at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY — used by oracle/gen_golden.py to produce tests/golden/*.npz and by
oracle/time_reference.py.  It never runs on the GPU box (/root/reference does not exist there)
and nothing under roborugby_b200/ imports it.

What it does:
  * puts oracle/shims (stub pygame + gym) and the reference root on sys.path;
  * selects the preset: the reference binds its constants at import time from one flag
    (RR_Constants.py:4, GAME_MODE).  For the TRAIN preset the module source is read from the
    read-only tree, the single flag line is flipped IN MEMORY and the module is installed in
    sys.modules before the package imports it; no reference file is copied or written;
  * works around the construction crash of the v0/v2 observers (RR_Observers.py:30-37,133-141)
    by pre-seeding the class-level observation_space;
  * extract()/inject() read and write the complete physics state of an env, including the
    redundant, independently drifting FloatRect fields (MyUtils.py:141-148) and the one robot
    history slot that is ever read (RR_Robot.py:43-58,110-137).

State layout (all float64 unless noted), R robots, B balls:
  rob[R, 7]   cx, cy, left, right, top, bottom, rot                 (FloatRect fields)
  rhist[R, 3] x, y, rot of history slot (count-1)  (valid iff rflag[:,2])
  rflag[R, 3] int32: thrust_l, thrust_r, hist_valid
  ball[B, 8]  cx, cy, left, right, top, bottom, vx, vy
  step        int32 lngStepCount
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("RR_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

ENV_IDS = {
    "RoboRugby-v0": ("robo_rugby.gym_env", "GameEnv"),
    "RoboRugbySimple-v0": ("robo_rugby.gym_env.RR_Environments", "SimpleChasePos"),
    "RoboRugbySimpleDuel-v2": ("robo_rugby.gym_env.RR_Environments", "SimpleDuel2"),
    "RoboRugbySimpleDuel-v3": ("robo_rugby.gym_env.RR_Environments", "SimpleDuel3"),
}


# In-memory source patches that switch on the reference's goal scoring "as intended" (SURVEY.md §8f rank 3).  At the
# reference's HEAD Goal.track_balls / update_score are only reached from GameEnv.__old_step (RR_EnvBase.py:458-520), whose
# commit block is commented out, so scores, goal destruction and ball removal are dead code.  The patches below make
# exactly that code live and repair what stops it from running; nothing else of the reference changes:
#   P1 RR_Goal.py:80      `sprBall.is_positive()` -> `sprBall.is_positive` (it is a property: the call raises TypeError)
#   P2 RR_EnvBase.py:294  before on_step_end(): the two track_balls calls (:463-464) and the commit block (:497-511,
#                         uncommented): balls that stayed TIME_BALL_IN_GOAL_STEPS in a goal are scored and killed; the
#                         score delta (`dblHappyScore = dblScoreDelta; dblGrumpyScore = -dblScoreDelta`, :510-511) is
#                         added to the scorekeepers' step rewards before their own on_step_end terms
#   P3 kill() also stops the ball (velocity and force 0): _roll_balls iterates lstBalls, dead balls included (:341-343),
#                         and a dead ball never gets on_frame_begin again, so without this it would coast on a stale force
#   P4 RR_EnvBase.py:204-207 reset() re-adds dead balls in lstBalls order (Group is insertion-ordered: a revived ball
#                         would otherwise move to the end of every pair enumeration)
_GOAL_SCORING_PATCHES = {
    "robo_rugby.gym_env.RR_Goal": [
        (b"                if sprBall.is_positive():", b"                if sprBall.is_positive:"),
    ],
    "robo_rugby.gym_env.RR_EnvBase": [
        (b"            self.on_frame_end()\n\n        self.on_step_end()\n",
         b"            self.on_frame_end()\n\n"
         b"        self.sprHappyGoal.track_balls(self.grpBalls.sprites())\n"
         b"        self.sprGrumpyGoal.track_balls(self.grpBalls.sprites())\n"
         b"        dblScoreDelta = 0\n"
         b"        for sprBall in self.sprHappyGoal.update_score():\n"
         b"            dblScoreDelta += const.POINTS_BALL_SCORED if sprBall.tplColor == const.COLOR_BALL_POS else -const.POINTS_BALL_SCORED\n"
         b"            sprBall.kill()\n"
         b"            sprBall.dbl_velocity_x = sprBall.dbl_velocity_y = 0\n"
         b"            sprBall.dbl_force_x = sprBall.dbl_force_y = 0\n"
         b"        for sprBall in self.sprGrumpyGoal.update_score():\n"
         b"            dblScoreDelta += -const.POINTS_BALL_SCORED if sprBall.tplColor == const.COLOR_BALL_POS else const.POINTS_BALL_SCORED\n"
         b"            sprBall.kill()\n"
         b"            sprBall.dbl_velocity_x = sprBall.dbl_velocity_y = 0\n"
         b"            sprBall.dbl_force_x = sprBall.dbl_force_y = 0\n"
         b"        self.dblGoalScoreDelta = dblScoreDelta\n"
         b"        if hasattr(self, 'reward_happy'):\n"
         b"            self.reward_happy += dblScoreDelta\n"
         b"            self.reward_grumpy -= dblScoreDelta\n\n"
         b"        self.on_step_end()\n"),
        (b"        for spr_ball in self.lstBalls:\n            if not spr_ball.alive():\n"
         b"                self.grpBalls.add(spr_ball)\n                self.grpAllSprites.add(spr_ball)\n",
         b"        if not all(b.alive() for b in self.lstBalls):\n"
         b"            for spr_ball in self.lstBalls:\n                spr_ball.kill()\n"
         b"            for spr_ball in self.lstBalls:\n"
         b"                self.grpBalls.add(spr_ball)\n                self.grpAllSprites.add(spr_ball)\n"),
    ],
}


def load_reference(preset, goal_scoring=False):
    """Import the reference with preset 'GAME' (as shipped) or 'TRAIN' (GAME_MODE=False); goal_scoring=True also
    applies _GOAL_SCORING_PATCHES.  All patches are applied to the module SOURCE IN MEMORY as it is read from the
    read-only tree; no reference file is copied or written."""
    assert preset in ("GAME", "TRAIN")
    if "robo_rugby" in sys.modules:
        raise RuntimeError("reference already imported in this process; constants bind at import")
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for p in (REF_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    patches = {}
    if preset == "TRAIN":
        patches["robo_rugby.gym_env.RR_Constants"] = [(b"GAME_MODE = True", b"GAME_MODE = False")]
    if goal_scoring:
        patches.update(_GOAL_SCORING_PATCHES)
    if patches:
        paths = {name: os.path.join(REF_ROOT, *name.split(".")) + ".py" for name in patches}

        class _Loader(importlib.machinery.SourceFileLoader):
            def get_data(self, p):
                data = super().get_data(p)
                for name, path in paths.items():
                    if p == path:
                        data = data.replace(b"\r\n", b"\n")
                        for old, new in patches[name]:
                            assert data.count(old) == 1, (name, old)
                            data = data.replace(old, new)
                return data

            def get_code(self, fullname):  # never use or write a .pyc for a patched module
                return self.source_to_code(self.get_data(paths[fullname]), paths[fullname])

        class _Finder(importlib.abc.MetaPathFinder):
            def find_spec(self, fullname, p=None, target=None):
                if fullname in paths:
                    return importlib.util.spec_from_file_location(fullname, paths[fullname], loader=_Loader(fullname, paths[fullname]))
                return None

        sys.meta_path.insert(0, _Finder())
    import robo_rugby  # noqa: F401  (registers the env ids with the stub gym)
    const = importlib.import_module("robo_rugby.gym_env.RR_Constants")
    assert const.GAME_MODE == (preset == "GAME")
    return const


# must stay identical to tests/golden_util.py CUSTOM_MIXINS (the tests translate the same lists into a config)
MIXIN_COMPOSITIONS = {
    "DuelAllMixins": ["KeepMovingGuys", "DontDriveInGoals", "BaseDestruction", "PushNegBallsFromGoal",
                      "PushPosBallsToGoal", "ChasePosBall", "NaughtyBots"],
    "DuelCutChain": ["DontDriveInGoals", "ChasePosBall", "NaughtyBots", "KeepMovingGuys"],
    # no NaughtyBots: its on_step_end does not call super(), which also cuts GameEnv.on_step_end and with it the goals'
    # own on_step_end (RR_Goal.py:58-64), without which no ball ever scores
    "DuelGoals": ["BaseDestruction", "PushPosBallsToGoal", "ChasePosBall"],
}


def make_env(env_id, through_gym=False):
    """Construct a reference env.  through_gym=True returns the TimeLimit-wrapped env."""
    import gym
    import numpy as np
    from robo_rugby.gym_env.RR_EnvBase import GameEnv
    const = importlib.import_module("robo_rugby.gym_env.RR_Constants")
    if (env_id in ("RoboRugbySimple-v0", "RoboRugbySimpleDuel-v2") or env_id in MIXIN_COMPOSITIONS) \
            and GameEnv.observation_space is None:
        hi = max(const.ARENA_WIDTH, const.ARENA_HEIGHT, 360)
        GameEnv.observation_space = gym.spaces.Box(-hi, hi, dtype=np.float32, shape=(5,))
    if env_id == "DuelAllCoordsPrior":
        # SimpleDuel2's reward mixins with the AllCoords_WithPrior observer (RR_Observers.py:86-110)
        import robo_rugby.gym_env.RR_ScoreKeepers as sk
        import robo_rugby.gym_env.RR_Observers as obs
        import robo_rugby.gym_env.RR_EnvBase as base

        class DuelAllCoordsPrior(sk.PushPosBallsToGoal, sk.ChasePosBall, sk.NaughtyBots, obs.AllCoords_WithPrior,
                                 base.GameEnv_Simple):
            pass

        env = DuelAllCoordsPrior()
        env.spec = gym.spec("RoboRugbySimpleDuel-v2")
        return env
    if env_id == "DuelAllCoords":
        # Not a registered id: SimpleDuel2's reward mixins composed with the AllCoords observer
        # (RR_Observers.py:47-83), the way main.py:42-49 composes ad-hoc classes.  Exercises observer O4.
        import robo_rugby.gym_env.RR_ScoreKeepers as sk
        import robo_rugby.gym_env.RR_Observers as obs
        import robo_rugby.gym_env.RR_EnvBase as base

        class DuelAllCoords(sk.PushPosBallsToGoal, sk.ChasePosBall, sk.NaughtyBots, obs.AllCoords, base.GameEnv_Simple):
            pass

        env = DuelAllCoords()
        env.spec = gym.spec("RoboRugbySimpleDuel-v2")
        return env
    if env_id == "DuelLidar6v1":
        # SimpleDuel2's reward mixins with the first 6-way lidar observer (RR_Observers.py:168-284), the one main.py:42-49
        # composes "so Stephen can play"
        import robo_rugby.gym_env.RR_ScoreKeepers as sk
        import robo_rugby.gym_env.RR_Observers as obs
        import robo_rugby.gym_env.RR_EnvBase as base

        class DuelLidar6v1(sk.PushPosBallsToGoal, sk.ChasePosBall, sk.NaughtyBots, obs.SingleBall_6wayLidar, base.GameEnv_Simple):
            pass

        env = DuelLidar6v1()
        env.spec = gym.spec("RoboRugbySimpleDuel-v2")
        return env
    if env_id in MIXIN_COMPOSITIONS:
        # ad-hoc reward-mixin compositions (class-definition order) on SimpleDuel2's observer and action space
        import robo_rugby.gym_env.RR_ScoreKeepers as sk
        import robo_rugby.gym_env.RR_Observers as obs
        import robo_rugby.gym_env.RR_EnvBase as base
        bases = tuple(getattr(sk, n) for n in MIXIN_COMPOSITIONS[env_id]) + (obs.PosBall_BasicLidar, base.GameEnv_Simple)
        env = type(env_id, bases, {})()
        env.spec = gym.spec("RoboRugbySimpleDuel-v2")
        return env
    if through_gym:
        return gym.make(env_id)
    mod, cls = ENV_IDS[env_id]
    env = getattr(importlib.import_module(mod), cls)()
    env.spec = gym.spec(env_id)
    return env


def stephen_hive(env, robot_indices):
    """The reference's "Stephen" players (DQN_pytorch_player.py) for the given robots WITHOUT their pickled network
    (absent from the tree): the objects are created without __init__ (which would load the pickle and insist on the
    v1 lidar observer) and registered in the class-level hive, so that the reference's own greedy nearest-ball
    assignment, Stephen.__ponder (:39-61), can be called.  Returns (Stephen class, [player per robot index])."""
    import DQN_pytorch_player as dp
    S = dp.Stephen
    S._Stephen__hive = set()
    S._Stephen__env = env.unwrapped
    players = []
    for i in robot_indices:
        p = S.__new__(S)
        p.env, p.robot = env.unwrapped, env.unwrapped.lstRobots[i]
        S._Stephen__hive.add(p)
        players.append(p)
    return S, players


def stephen_assignments(env, S, players):
    """Run Stephen.__ponder and return the assigned ball index per player (-1 = none)."""
    S._Stephen__ponder()
    asg = S._Stephen__assignments
    balls = env.unwrapped.lstBalls
    return [balls.index(asg[p]) if p in asg else -1 for p in players]


def reset_scratch():
    """Put the module-global scratch rect (RR_TrashyPhysics.py:29-35) back to its import-time state.

    Its centre is moved incrementally (cx += new - cx), so its low bits depend on every earlier
    call in the process; golden records start from the pristine value so they are reproducible.
    """
    tp = importlib.import_module("robo_rugby.gym_env.RR_TrashyPhysics")
    r = tp._rectBallInner
    h = tp._dblHalfRadius_Rad2
    r._dblCenterX = (-h + h) / 2
    r._dblCenterY = (-h + h) / 2
    r._dblLeft, r._dblRight, r._dblTop, r._dblBottom = -h, h, -h, h
    r._dblRotation = 0
    r._dctCornersRelCenter = r._dctInitialCornersRelCenter.copy()


def extract(env):
    import numpy as np
    from robo_rugby.gym_env.RR_Robot import Robot
    env = env.unwrapped
    R, B = len(env.lstRobots), len(env.lstBalls)
    rob = np.zeros((R, 7)); rhist = np.zeros((R, 3)); rflag = np.zeros((R, 3), np.int32)
    for i, r in enumerate(env.lstRobots):
        f = r.rectDbl
        rob[i] = (f._dblCenterX, f._dblCenterY, f._dblLeft, f._dblRight, f._dblTop, f._dblBottom, f._dblRotation)
        rflag[i, 0], rflag[i, 1] = r.lngLThrust, r.lngRThrust
        slot = r._lstStates[(r.lngMoveCount - 1) % Robot._slngMoveHistorySize]
        if slot is not None and slot[3] == r.lngMoveCount - 1:
            rhist[i] = slot[:3]
            rflag[i, 2] = 1
    ball = np.zeros((B, 8))
    for i, b in enumerate(env.lstBalls):
        f = b.rectDbl
        assert f._dblRotation == 0
        ball[i] = (f._dblCenterX, f._dblCenterY, f._dblLeft, f._dblRight, f._dblTop, f._dblBottom,
                   b.dbl_velocity_x, b.dbl_velocity_y)
    return dict(rob=rob, rhist=rhist, rflag=rflag, ball=ball, step=np.int32(env.lngStepCount))


def inject(env, st):
    """Overwrite the env's physics state with `st` (same layout as extract()).

    Fresh FloatRects are built so no alias to an older rect survives; the rotation setter is used
    for the rotation-dependent corner table (a pure function of the stored angle,
    MyUtils.py:277-322), then the centre and L/R/T/B doubles are written verbatim.
    """
    from MyUtils import FloatRect
    from robo_rugby.gym_env.RR_Robot import Robot
    const = importlib.import_module("robo_rugby.gym_env.RR_Constants")
    env = env.unwrapped
    for i, r in enumerate(env.lstRobots):
        cx, cy, L, Rr, T, Bm, rot = (float(v) for v in st["rob"][i])
        f = FloatRect(0, const.ROBOT_LENGTH, 0, const.ROBOT_WIDTH)
        f.rotation = rot
        assert f._dblRotation == rot or (rot == 0 and f._dblRotation == 0), (f._dblRotation, rot)
        f._dblCenterX, f._dblCenterY = cx, cy
        f._dblLeft, f._dblRight, f._dblTop, f._dblBottom = L, Rr, T, Bm
        r.rectDbl = f
        r.lngLThrust, r.lngRThrust = int(st["rflag"][i, 0]), int(st["rflag"][i, 1])
        r.lngFrameMass = const.MASS_ROBOT
        r._lstStates = [None] * Robot._slngMoveHistorySize
        if int(st["rflag"][i, 2]):
            # any count >= 1 is equivalent; use 1 so slot 0 holds the older pose
            r.lngMoveCount = 1
            hx, hy, hr = (float(v) for v in st["rhist"][i])
            r._lstStates[0] = (hx, hy, hr, 0)
        else:
            r.lngMoveCount = 0
        r.rectDblPriorStep = r.rectDbl.copy()
    for i, b in enumerate(env.lstBalls):
        cx, cy, L, Rr, T, Bm, vx, vy = (float(v) for v in st["ball"][i])
        f = FloatRect(0, 2 * const.BALL_RADIUS, 0, 2 * const.BALL_RADIUS)
        f._dblCenterX, f._dblCenterY = cx, cy
        f._dblLeft, f._dblRight, f._dblTop, f._dblBottom = L, Rr, T, Bm
        b.rectDbl = f
        b.dbl_velocity_x, b.dbl_velocity_y = vx, vy
        b.dbl_force_x = b.dbl_force_y = 0
        b.lngFrameMass = const.MASS_BALL
        b.bln_moved_cur_frame = False
        b.rectDblPriorStep = b.rectDbl.copy()
        b.rectDblPriorFrame = b.rectDbl.copy()
    env.lngStepCount = int(st["step"])
    for g in (env.sprHappyGoal, env.sprGrumpyGoal):
        g.on_reset()
    if hasattr(env, "set_naughty_bots"):
        env.set_naughty_bots.clear()
    if hasattr(env, "reward_happy"):
        env.reward_happy = env.reward_grumpy = 0.0
