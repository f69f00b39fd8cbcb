#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only, /root/reference).

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (the reference tree does not exist on
the GPU box); its outputs are committed so that every later test is self-contained.

    python oracle/gen_golden.py            # regenerate everything (a few minutes on 8 cores)

Each file holds trajectories `[n, T+1, ...]` of complete physics states (layout documented in
oracle/ref_harness.py), the actions fed to the reference's own step() and everything step()
returned.  The module-global scratch rect of the reference is restored to its import-time value
before every step (ref_harness.reset_scratch) so that records are reproducible in isolation.

Kinds:
  random   uniformly random discrete actions from a reference reset()
  chase    robots steer at the nearest ball (15 % random) -> contact-rich play
  inject   single-step records from hand-built, contact-free but about-to-collide states:
           ball in front of a robot, robot pairs, ball pairs, balls and robots at every wall,
           axis-aligned headings, squeezes
  reset    (state before, randint draws, state after, first observation) of reference reset()
  timelimit end-of-episode done flags, raw env and gym TimeLimit wrapper
"""
import math
import multiprocessing as mp
import os
import random
import signal
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

STATE_KEYS = ("rob", "rhist", "rflag", "ball", "step")


class _Timeout(Exception):
    pass


def _alarm(sig, frm):
    raise _Timeout()


def _obs(env, team):
    o = env.unwrapped.get_game_state(int_team=team)
    return None if o is None else np.asarray(o, np.float64)


def _step_record(H, env, actions, obs_dim):
    """Run one reference step; returns (after_state, obs_h, obs_g, rew[2], done, naughty, exc)."""
    import robo_rugby.gym_env.RR_Constants as const
    H.reset_scratch()
    exc = ""
    try:
        _, r_h, done, info = env.step(actions)
        r_g = info.dblGrumpyScore
    except _Timeout:
        raise
    except Exception as e:  # the reference signals unresolved physics with bare Exceptions
        exc = str(e)[:60]
        r_h = r_g = float("nan")
        done = False
    u = env.unwrapped
    oh = _obs(env, const.TEAM_HAPPY)
    og = _obs(env, const.TEAM_GRUMPY)
    nan = np.full(obs_dim, np.nan)
    naughty = len(u.set_naughty_bots) if hasattr(u, "set_naughty_bots") else 0
    return (H.extract(env), nan if oh is None else oh, nan if og is None else og,
            np.array([r_h, r_g], np.float64), bool(done), naughty, exc)


def _chase_action(u, i, rng):
    from MyUtils import angle_degrees, distance
    r = u.lstRobots[i]
    if rng.random() < 0.15:
        return rng.randrange(8)
    c = r.rectDbl.center
    tgt = min(u.lstBalls, key=lambda b: distance(c, b.rectDbl.center))
    want = angle_degrees(c, tgt.rectDbl.center)
    diff = (want - r.rectDbl.rotation + 540) % 360 - 180
    if abs(diff) < 8:
        return 0
    if abs(diff) < 30:
        return 4 if diff > 0 else 5
    return 2 if diff > 0 else 3


def _empty(n, T, R, B, A, D):
    return dict(
        rob=np.zeros((n, T + 1, R, 7)), rhist=np.zeros((n, T + 1, R, 3)),
        rflag=np.zeros((n, T + 1, R, 3), np.int32), ball=np.zeros((n, T + 1, B, 8)),
        step=np.zeros((n, T + 1), np.int32), act=np.full((n, T, A), np.nan),
        obs_h=np.full((n, T, D), np.nan), obs_g=np.full((n, T, D), np.nan),
        rew=np.zeros((n, T, 2)), done=np.zeros((n, T), np.int8), naughty=np.zeros((n, T), np.int8),
        exc=np.zeros((n, T), np.int8), restart=np.zeros((n, T + 1), np.int8))


def _put_state(out, i, t, st):
    for k in STATE_KEYS:
        out[k][i, t] = st[k]


def _obs_dim(env_id, R, B):
    return {"RoboRugby-v0": 0, "RoboRugbySimple-v0": 5, "RoboRugbySimpleDuel-v2": 5,
            "RoboRugbySimpleDuel-v3": 11, "DuelAllCoords": 3 * R + 2 * B, "DuelAllMixins": 5, "DuelCutChain": 5,
            "DuelAllCoordsPrior": 6 * R + 4 * B, "DuelLidar6v1": 11, "DuelGoals": 5}[env_id]


# ------------------------------------------------------------------ state builders for `inject`

def _place(f, cx, cy, rot=None):
    f.center = (float(cx), float(cy))
    if rot is not None:
        f.rotation = float(rot)


def _clean(u):
    """True when the state has no contact of any kind (required before injecting)."""
    import robo_rugby.gym_env.RR_TrashyPhysics as tp
    for i, a in enumerate(u.lstRobots):
        f = a.rectDbl
        if f.left < 0.5 or f.right > u.rect_walls.right - 0.5 or f.top <= 0.5 or f.bottom >= u.rect_walls.bottom - 0.5:
            return False
        for b in u.lstRobots[i + 1:]:
            if tp.robots_collided(a, b):
                return False
    for i, a in enumerate(u.lstBalls):
        if tp.collided_wall(a):
            return False
        f = a.rectDbl
        if f.left < 0 or f.right > u.rect_walls.right or f.top <= 0 or f.bottom >= u.rect_walls.bottom:
            return False
        for b in u.lstBalls[i + 1:]:
            if tp.balls_collided(a, b):
                return False
        for r in u.lstRobots:
            if tp.ball_robot_collided(a, r):
                return False
    return True


def _build_injected(u, rng, const):
    """Mutate a freshly reset env into an about-to-collide state.  Returns suggested actions."""
    W, Hh = const.ARENA_WIDTH, const.ARENA_HEIGHT
    R, B = len(u.lstRobots), len(u.lstBalls)
    acts = [rng.randrange(8) for _ in range(R)]
    case = rng.randrange(10)
    axis = rng.random() < 0.3
    rot_choices = [0, 45, 90, 135, 180, 225, 270, 315, 360]

    def rot():
        return rng.choice(rot_choices) if axis else rng.uniform(0, 360)

    def ball_vel(b, scale=1.5, p=0.6):
        if rng.random() < p:
            b.dbl_velocity_x = rng.uniform(-scale, scale)
            b.dbl_velocity_y = rng.uniform(-scale, scale)

    j = rng.randrange(R)
    rb = u.lstRobots[j]
    if case in (0, 1, 2):  # ball just outside a robot, robot usually drives/turns into it
        _place(rb.rectDbl, rng.uniform(120, W - 120), rng.uniform(120, Hh - 120), rot())
        b = u.lstBalls[rng.randrange(B)]
        ang = math.radians(rng.uniform(0, 360))
        if case == 2:  # aim at a corner region
            ang = math.radians(rb.rectDbl.rotation + rng.choice([63.4, 116.6, 243.4, 296.6]) + rng.uniform(-8, 8))
        d = rng.uniform(16.5, 36)
        _place(b.rectDbl, rb.rectDbl.centerx + d * math.cos(ang), rb.rectDbl.centery - d * math.sin(ang))
        ball_vel(b)
        acts[j] = rng.choice([0, 0, 1, 2, 3, 4, 5, 6, 7])
    elif case == 3 and R > 1:  # robot pair about to touch
        k = (j + 1 + rng.randrange(R - 1)) % R
        _place(rb.rectDbl, rng.uniform(150, W - 150), rng.uniform(150, Hh - 150), rot())
        ang = math.radians(rng.uniform(0, 360)); d = rng.uniform(22, 58)
        _place(u.lstRobots[k].rectDbl, rb.rectDbl.centerx + d * math.cos(ang), rb.rectDbl.centery - d * math.sin(ang), rot())
    elif case == 4 and B > 1:  # ball pair closing
        a = u.lstBalls[rng.randrange(B)]
        others = [x for x in u.lstBalls if x is not a]
        b = rng.choice(others)
        _place(a.rectDbl, rng.uniform(60, W - 60), rng.uniform(60, Hh - 60))
        ang = rng.uniform(0, 2 * math.pi); d = rng.uniform(14.05, 22)
        _place(b.rectDbl, a.rectDbl.centerx + d * math.cos(ang), a.rectDbl.centery + d * math.sin(ang))
        s = rng.uniform(0.05, 1.2)
        a.dbl_velocity_x, a.dbl_velocity_y = s * math.cos(ang), s * math.sin(ang)
        if rng.random() < 0.7:
            s2 = rng.uniform(0.0, 1.2)
            b.dbl_velocity_x, b.dbl_velocity_y = -s2 * math.cos(ang), -s2 * math.sin(ang)
        if rng.random() < 0.3 and len(others) > 1:  # third ball for chains
            c = rng.choice([x for x in others if x is not b])
            _place(c.rectDbl, b.rectDbl.centerx + 14.5 * math.cos(ang + 0.4), b.rectDbl.centery + 14.5 * math.sin(ang + 0.4))
    elif case in (5, 6):  # ball at a wall (or a corner) moving outwards
        b = u.lstBalls[rng.randrange(B)]
        wall = rng.randrange(6)
        x, y = rng.uniform(30, W - 30), rng.uniform(30, Hh - 30)
        vx, vy = rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5)
        g = rng.uniform(7.02, 10.5)
        if wall in (0, 4): x, vx = g, -abs(vx) - 0.05
        if wall in (1, 5): x, vx = W - g, abs(vx) + 0.05
        if wall in (2, 4): y, vy = g, -abs(vy) - 0.05
        if wall in (3, 5): y, vy = Hh - g, abs(vy) + 0.05
        _place(b.rectDbl, x, y)
        b.dbl_velocity_x, b.dbl_velocity_y = vx, vy
        if case == 6:  # robot pushing the ball into the wall: squeeze
            dx = 1 if x < W / 2 else -1
            _place(rb.rectDbl, x + dx * rng.uniform(19, 30), y + rng.uniform(-6, 6), rng.choice([0, 180, 0.5, 179.5]) if axis else rng.uniform(-20, 20) % 360)
            acts[j] = 1 if (dx > 0) == (rb.rectDbl.rotation < 90 or rb.rectDbl.rotation > 270) else 0
    elif case == 7:  # robot at a wall heading into it
        wall = rng.randrange(4)
        r0 = rot()
        x, y = rng.uniform(60, W - 60), rng.uniform(60, Hh - 60)
        _place(rb.rectDbl, x, y, r0)
        f = rb.rectDbl
        g = rng.uniform(0.6, 9)
        if wall == 0: f.left = g
        if wall == 1: f.right = W - g
        if wall == 2: f.top = g
        if wall == 3: f.bottom = Hh - g
    elif case == 8 and R > 1:  # ball squeezed between two robots
        k = (j + 1 + rng.randrange(R - 1)) % R
        r0 = rot()
        _place(rb.rectDbl, rng.uniform(200, W - 200), rng.uniform(200, Hh - 200), r0)
        th = math.radians(rb.rectDbl.rotation)
        b = u.lstBalls[rng.randrange(B)]
        d1 = rng.uniform(17.2, 20)
        _place(b.rectDbl, rb.rectDbl.centerx + d1 * math.cos(th), rb.rectDbl.centery - d1 * math.sin(th))
        d2 = d1 + rng.uniform(17.2, 20)
        _place(u.lstRobots[k].rectDbl, rb.rectDbl.centerx + d2 * math.cos(th), rb.rectDbl.centery - d2 * math.sin(th), r0 + 180)
        acts[j] = 0; acts[k] = 0
    else:  # moving balls + thrust already set on robots that will get no action
        for b in u.lstBalls:
            ball_vel(b, 2.0, 0.8)
    for r in u.lstRobots:  # pre-existing thrust (kept when no action is supplied, RR_EnvBase.py:272-273)
        if rng.random() < 0.5:
            r.lngLThrust, r.lngRThrust = rng.choice([-1, 0, 1]), rng.choice([-1, 0, 1])
    return acts


# ------------------------------------------------------------------ tasks (one process each)

def task_rollout(args):
    preset, env_id, kind, seed, n, T = args
    import ref_harness as H
    const = H.load_reference(preset)
    env = H.make_env(env_id)
    u = env.unwrapped
    R, B = len(u.lstRobots), len(u.lstBalls)
    discrete = env_id != "RoboRugby-v0"
    n_act = 1 if env_id == "RoboRugbySimple-v0" else R
    if kind == "partial":   # only the happy robots are driven: the others keep zero thrust and never move (KeepMovingGuys)
        n_act = max(1, R // 2)
    A = R if discrete else 2 * R
    D = _obs_dim(env_id, R, B)
    out = _empty(n, T, R, B, A, D)
    rng = random.Random(seed)
    t0 = time.time()
    signal.signal(signal.SIGALRM, _alarm)
    i = attempt = 0
    while i < n:
        # The reference can spin forever in _undo_naughty_movement (RR_EnvBase.py:415-421, GAME_MODE
        # only prints a warning), e.g. when reset() leaves two balls exactly 14 px apart.  Such
        # trajectories are abandoned and redrawn.
        random.seed(seed * 1000 + attempt)
        attempt += 1
        signal.alarm(15 + 2 * T)
        try:
            _rollout_one(H, env, u, out, i, T, R, D, n_act, discrete, kind, rng)
            i += 1
        except _Timeout:
            print(f"timeout in {args}, attempt {attempt}", flush=True)
        finally:
            signal.alarm(0)
    out["_secs"] = time.time() - t0
    return args, out


def _rollout_one(H, env, u, out, i, T, R, D, n_act, discrete, kind, rng):
    if True:
        env.reset()
        st = H.extract(env)
        H.inject(env, st)
        _put_state(out, i, 0, st)
        hold = [0] * R
        for t in range(T):
            if discrete:
                if kind == "chase":
                    acts = [_chase_action(u, k, rng) for k in range(n_act)]
                elif kind == "sticky":
                    acts = [hold[k] if rng.random() < 0.9 else rng.randrange(8) for k in range(n_act)]
                    hold[:n_act] = acts
                else:
                    acts = [rng.randrange(8) for _ in range(n_act)]
                out["act"][i, t, :n_act] = acts
                call = list(acts)
            else:
                vals = [rng.uniform(-1.45, 1.45) for _ in range(2 * R)]
                if rng.random() < 0.2:
                    vals[rng.randrange(2 * R)] = rng.choice([0.5, -0.5, 1.5, -1.5, 2.5])
                out["act"][i, t, :] = vals
                call = [tuple(vals[2 * k:2 * k + 2]) for k in range(R)]
            st, oh, og, rew, done, ng, exc = _step_record(H, env, call, D)
            _put_state(out, i, t + 1, st)
            out["obs_h"][i, t], out["obs_g"][i, t] = oh, og
            out["rew"][i, t], out["done"][i, t], out["naughty"][i, t] = rew, done, ng
            out["exc"][i, t] = 1 if exc else 0
            if exc:
                # The reference raised mid-step (e.g. "UNABLE TO RESOLVE ALL COLLISIONS FOR FRAME",
                # RR_EnvBase.py:421, when a robot squeezes a ball against a wall).  The env is left
                # half-updated; restart the trajectory from a fresh reset and flag the restart.
                print(f"exception in rollout {kind} traj {i} step {t}: {exc}", flush=True)
                env.reset()
                st = H.extract(env)
                H.inject(env, st)
                _put_state(out, i, t + 1, st)
                out["restart"][i, t + 1] = 1


def task_inject(args):
    preset, env_id, seed, n = args
    import ref_harness as H
    const = H.load_reference(preset)
    env = H.make_env(env_id)
    u = env.unwrapped
    R, B = len(u.lstRobots), len(u.lstBalls)
    n_act_max = 1 if env_id == "RoboRugbySimple-v0" else R
    D = _obs_dim(env_id, R, B)
    out = _empty(n, 1, R, B, R, D)
    rng = random.Random(seed)
    signal.signal(signal.SIGALRM, _alarm)
    i = 0
    stats = dict(tried=0, dirty=0, timeout=0, exc=0)
    while i < n:
        stats["tried"] += 1
        random.seed(rng.randrange(1 << 30))
        env.reset()
        acts = _build_injected(u, rng, const)
        if not _clean(u):
            stats["dirty"] += 1
            continue
        st = H.extract(env)
        st["step"] = np.int32(rng.randrange(0, const.GAME_LENGTH_STEPS - 1))
        # robots that "have moved before" carry a history slot: previous pose = a small step back
        for k, r in enumerate(u.lstRobots):
            if rng.random() < 0.8:
                st["rflag"][k, 2] = 1
                st["rhist"][k] = (st["rob"][k, 0] + rng.uniform(-1, 1), st["rob"][k, 1] + rng.uniform(-1, 1),
                                  (st["rob"][k, 6] + rng.choice([0, 0, 0.6, -0.6, 1.2, -1.2]) + 720) % 360)
        H.inject(env, st)
        n_act = rng.choice([n_act_max, n_act_max, n_act_max, max(1, n_act_max - 1), 1])
        signal.alarm(20)
        try:
            st2, oh, og, rew, done, ng, exc = _step_record(H, env, acts[:n_act], D)
        except _Timeout:
            stats["timeout"] += 1
            continue
        finally:
            signal.alarm(0)
        if exc:
            stats["exc"] += 1
        _put_state(out, i, 0, st)
        out["act"][i, 0, :n_act] = acts[:n_act]
        _put_state(out, i, 1, st2)
        out["obs_h"][i, 0], out["obs_g"][i, 0] = oh, og
        out["rew"][i, 0], out["done"][i, 0], out["naughty"][i, 0] = rew, done, ng
        out["exc"][i, 0] = 1 if exc else 0
        i += 1
    out["_stats"] = stats
    return args, out


def task_reset(args):
    preset, env_id, seed, n = args
    import ref_harness as H
    const = H.load_reference(preset)
    env = H.make_env(env_id)
    u = env.unwrapped
    R, B = len(u.lstRobots), len(u.lstBalls)
    D = _obs_dim(env_id, R, B)
    out = _empty(n, 1, R, B, R, D)
    draws = np.full((n, 512), -1, np.int32)
    ndraws = np.zeros(n, np.int32)
    orig = random.randint
    log = []

    def logged(a, b):
        v = orig(a, b)
        log.append(v)
        return v

    rng = random.Random(seed)
    signal.signal(signal.SIGALRM, _alarm)
    for i in range(n):
        random.seed(seed * 77 + i)
        # put the env somewhere "used": a few steps after a previous reset
        for _ in range(rng.randrange(0, 4)):
            signal.alarm(10)
            try:
                env.step([rng.randrange(8)])
            except _Timeout:
                break
            finally:
                signal.alarm(0)
        _put_state(out, i, 0, H.extract(env))
        log.clear()
        random.randint = logged
        try:
            env.reset()
        finally:
            random.randint = orig
        assert len(log) <= 512
        draws[i, :len(log)] = log
        ndraws[i] = len(log)
        _put_state(out, i, 1, H.extract(env))
        oh, og = _obs(env, const.TEAM_HAPPY), _obs(env, const.TEAM_GRUMPY)
        if oh is not None: out["obs_h"][i, 0] = oh
        if og is not None: out["obs_g"][i, 0] = og
    out["draws"], out["ndraws"] = draws, ndraws
    return args, out


def task_timelimit(args):
    preset, env_id = args
    import ref_harness as H
    const = H.load_reference(preset)
    T = const.GAME_LENGTH_STEPS
    env = H.make_env(env_id, through_gym=True)
    random.seed(5)
    env.reset()
    st = H.extract(env); st["step"] = np.int32(T - 3)
    H.inject(env, st)
    env._elapsed_steps = T - 3
    wrapped, trunc = [], []
    for _ in range(3):
        _, _, d, info = env.step([0])
        wrapped.append(bool(d)); trunc.append(bool(info.get("TimeLimit.truncated", False)))
    raw = []
    u = env.unwrapped
    H.inject(u, st)
    raised = False
    for _ in range(6):
        try:
            _, _, d, _ = u.step([0])
            raw.append(bool(d))
        except Exception:
            raised = True
            break
    return args, dict(T=np.int32(T), start_step=np.int32(T - 3), wrapped_done=np.array(wrapped), truncated=np.array(trunc),
                      raw_done=np.array(raw), raised_after=np.int32(len(raw)), raised=np.bool_(raised))


def _save(name, out):
    os.makedirs(OUT, exist_ok=True)
    meta = {k: out.pop(k) for k in list(out) if k.startswith("_")}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB {meta}", flush=True)


def task_reset_fixed(args):
    """reset(False) records: (state before, starting layout, state after, first observations).

    mode 'own': env built with CONFIG_RANDOM, so the layout is its construction-time placement
    (RR_EnvBase.py:112-113); mode 'standard' (GAME only): env built with CONFIG_STANDARD (:35-52,114-116),
    whose record 0 is the freshly constructed state itself (state before = sprites at the origin)."""
    preset, env_id, mode, seed, n = args
    import importlib
    import ref_harness as H
    const = H.load_reference(preset)
    from robo_rugby.gym_env.RR_EnvBase import GameEnv
    random.seed(seed)
    if mode == "standard":
        mod, cls = H.ENV_IDS[env_id]
        H.make_env(env_id)  # seeds the class-level spaces (construction work-around)
        env = getattr(importlib.import_module(mod), cls)(GameEnv.CONFIG_STANDARD)
    else:
        env = H.make_env(env_id)
    u = env.unwrapped
    R, B = len(u.lstRobots), len(u.lstBalls)
    D = _obs_dim(env_id, R, B)
    out = _empty(n, 1, R, B, R, D)
    start = np.zeros((n, 3 * R + 2 * B))
    rng = random.Random(seed)
    signal.signal(signal.SIGALRM, _alarm)
    for i in range(n):
        lay = u._lst_starting_positions
        start[i, :3 * R] = np.asarray(lay[0], np.float64).reshape(-1)
        start[i, 3 * R:] = np.asarray(lay[1], np.float64).reshape(-1)
        if not (mode == "standard" and i == 0):
            for _ in range(rng.randrange(1, 5)):
                signal.alarm(10)
                try:
                    env.step([rng.randrange(8) for _ in range(R)])
                except Exception:
                    break
                finally:
                    signal.alarm(0)
            _put_state(out, i, 0, H.extract(env))
            env.reset(False)
        else:
            out["restart"][i, 0] = 1  # marks "state before = freshly constructed sprites"
        _put_state(out, i, 1, H.extract(env))
        oh, og = _obs(env, const.TEAM_HAPPY), _obs(env, const.TEAM_GRUMPY)
        if oh is not None: out["obs_h"][i, 0] = oh
        if og is not None: out["obs_g"][i, 0] = og
    out["start"] = start
    return args, out


def task_entity(args):
    """Per-robot / per-ball observations and the "Stephen" ball assignment along chase rollouts:
    ent[i, t, r, b] = get_game_state(obj_robot=lstRobots[r], obj_ball=lstBalls[b]) after step t (RR_Observers.py:133-136,
    :187-203, :304-320), asg[i, t] = Stephen.__ponder's assignment for the players of `hive` (DQN_pytorch_player.py:39-61)."""
    preset, env_id, seed, n, T, hive = args
    import ref_harness as H
    const = H.load_reference(preset)
    env = H.make_env(env_id)
    u = env.unwrapped
    R, B = len(u.lstRobots), len(u.lstBalls)
    D = _obs_dim(env_id, R, B)
    out = _empty(n, T, R, B, R, D)
    out["ent"] = np.full((n, T, R, B, D), np.nan)
    out["asg"] = np.full((n, T, len(hive)), -2, np.int32)
    out["hive"] = np.asarray(hive, np.int32)
    S, players = H.stephen_hive(env, hive)
    rng = random.Random(seed)
    signal.signal(signal.SIGALRM, _alarm)
    i = attempt = 0
    while i < n:
        random.seed(seed * 1000 + attempt)
        attempt += 1
        signal.alarm(30 + 3 * T)
        try:
            env.reset()
            st = H.extract(env)
            H.inject(env, st)
            _put_state(out, i, 0, st)
            ok = True
            for t in range(T):
                acts = [_chase_action(u, k, rng) for k in range(R)]
                out["act"][i, t, :R] = acts
                st, oh, og, rew, done, ng, exc = _step_record(H, env, list(acts), D)
                if exc:
                    ok = False
                    break
                _put_state(out, i, t + 1, st)
                out["obs_h"][i, t], out["obs_g"][i, t] = oh, og
                out["rew"][i, t], out["done"][i, t], out["naughty"][i, t] = rew, done, ng
                for r in range(R):
                    for b in range(B):
                        out["ent"][i, t, r, b] = np.asarray(u.get_game_state(obj_robot=u.lstRobots[r], obj_ball=u.lstBalls[b]), np.float64)
                out["asg"][i, t] = H.stephen_assignments(env, S, players)
            if ok:
                i += 1
        except _Timeout:
            print(f"timeout in {args}, attempt {attempt}", flush=True)
        finally:
            signal.alarm(0)
    return args, out


def task_goals(args):
    """Goal scoring "as intended" (ref_harness._GOAL_SCORING_PATCHES): trajectories of T steps from states with balls
    parked inside the goal triangles; records the usual step outputs (the reward includes the +-500 score delta) plus,
    after every step: alive[B], score[2] (happy goal, grumpy goal get_score()), destroyed[2], dwell[2][B] (steps the ball
    has been in that goal so far, 0 = not tracked)."""
    preset, env_id, seed, n, T = args
    import ref_harness as H
    const = H.load_reference(preset, goal_scoring=True)
    env = H.make_env(env_id)
    u = env.unwrapped
    R, B = len(u.lstRobots), len(u.lstBalls)
    D = _obs_dim(env_id, R, B)
    W, Hh = const.ARENA_WIDTH, const.ARENA_HEIGHT
    continuous = env_id == "RoboRugby-v0"
    thrust = {0: (1, 1), 1: (-1, -1), 2: (-1, 1), 3: (1, -1), 4: (0, 1), 5: (1, 0), 6: (-1, 0), 7: (0, -1)}
    out = _empty(n, T, R, B, 2 * R if continuous else R, D)
    out["alive"] = np.zeros((n, T + 1, B), np.int32)
    out["score"] = np.zeros((n, T + 1, 2), np.int32)
    out["destroyed"] = np.zeros((n, T + 1, 2), np.int32)
    out["dwell"] = np.zeros((n, T + 1, 2, B), np.int32)
    out["delta"] = np.zeros((n, T), np.int32)
    rng = random.Random(seed)
    signal.signal(signal.SIGALRM, _alarm)

    def goal_record(i, t):
        out["alive"][i, t] = [int(b.alive()) for b in u.lstBalls]
        for gi, g in enumerate((u.sprHappyGoal, u.sprGrumpyGoal)):
            out["score"][i, t, gi] = g.get_score()
            out["destroyed"][i, t, gi] = int(g.is_destroyed())
            for bi, b in enumerate(u.lstBalls):
                out["dwell"][i, t, gi, bi] = (g.lngFrameCount - g.dctBallsPrior[b] + 1) if b in g.dctBallsPrior else 0

    i = attempt = 0
    while i < n:
        random.seed(seed * 1000 + attempt)
        attempt += 1
        signal.alarm(60 + 3 * T)
        try:
            env.reset()
            # park some balls (at rest) inside the goal triangles: happy = bottom right, grumpy = top left
            kind = i % 4
            spots_h = [(W - 40 - 35 * k, Hh - 40 - 30 * (k % 2)) for k in range(5)]
            spots_g = [(40 + 35 * k, 40 + 30 * (k % 2)) for k in range(5)]
            neg0 = const.NUM_BALL_POS
            if kind == 0:    # three negative balls + one positive in the happy goal: it is destroyed
                plan = [(neg0 + k, spots_h[k]) for k in range(min(3, const.NUM_BALL_NEG))] + [(0, spots_h[3])]
            elif kind == 1:  # one of each in the grumpy goal, one positive in the happy goal
                plan = [(0, spots_g[0]), (neg0, spots_g[1]), (1 % const.NUM_BALL_POS, spots_h[0])] if const.NUM_BALL_NEG else [(0, spots_g[0])]
            elif kind == 2:  # three negatives in the grumpy goal
                plan = [(neg0 + k, spots_g[k]) for k in range(min(3, const.NUM_BALL_NEG))] if const.NUM_BALL_NEG else [(0, spots_h[0])]
            else:            # every ball in some goal: the game ends when none is left
                plan = [(b, (spots_h if b % 2 else spots_g)[b // 2]) for b in range(B)]
            for b, (x, y) in plan:
                u.lstBalls[b].rectDbl.center = (float(x), float(y))
            for ri, rb in enumerate(u.lstRobots):   # the robots start in the middle of the arena, clear of every parked ball
                rb.rectDbl.center = (W / 2 - 90.0 + 60.0 * ri, Hh / 2 - 60.0 + 45.0 * ri)
            for bi, bl in enumerate(u.lstBalls):    # and the balls that are not parked keep away from them
                if bi not in [b for b, _ in plan]:
                    bl.rectDbl.center = (80.0 + 70.0 * bi, Hh - 90.0 - 40.0 * (bi % 3)) if bi % 2 else (W - 80.0 - 70.0 * bi, 90.0 + 40.0 * (bi % 3))
            st = H.extract(env)
            H.inject(env, st)
            _put_state(out, i, 0, st)
            goal_record(i, 0)
            ok = True
            # kinds 0, 2: robot 0 spins on the spot (GameEnv_Simple.step needs at least one command: np.concatenate of an
            # empty list raises), the others stand still; kinds 1, 3: chasing robots, which push balls out of the goals
            driven = 1 if kind in (0, 2) else (R if kind == 1 else 1)
            for t in range(T):
                if u.game_is_done():
                    out["exc"][i, t:] = 2   # marks "episode over": no further steps recorded
                    for tt in range(t, T):
                        _put_state(out, i, tt + 1, H.extract(env)); goal_record(i, tt + 1)
                    break
                acts = [2] if kind in (0, 2) else [_chase_action(u, k, rng) for k in range(driven)]
                if continuous:   # GameEnv.step takes (left, right) thrust pairs
                    call = [tuple(float(v) for v in thrust[a]) for a in acts]
                    out["act"][i, t, :2 * driven] = [v for pr in call for v in pr]
                else:
                    call = list(acts)
                    out["act"][i, t, :driven] = acts
                stt, oh, og, rew, done, ng, exc = _step_record(H, env, call, D)
                if exc:
                    ok = False
                    break
                _put_state(out, i, t + 1, stt)
                out["obs_h"][i, t], out["obs_g"][i, t] = oh, og
                out["rew"][i, t], out["done"][i, t], out["naughty"][i, t] = rew, done, ng
                out["delta"][i, t] = u.dblGoalScoreDelta
                goal_record(i, t + 1)
            if ok:
                i += 1
        except _Timeout:
            print(f"timeout in {args}, attempt {attempt}", flush=True)
        finally:
            signal.alarm(0)
    return args, out


def main_goals():
    """goal scoring as intended, from the in-memory patched reference."""
    # (SimpleDuel-v2 has NaughtyBots: its cut on_step_end chain keeps the goals from ever scoring, a negative case)
    jobs = [("GAME", "DuelGoals", 101, 8, 165), ("GAME", "RoboRugby-v0", 102, 4, 165),
            ("GAME", "RoboRugbySimpleDuel-v2", 103, 2, 160), ("TRAIN", "DuelGoals", 104, 4, 165)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(processes=4, maxtasksperchild=1) as pool:
        res = [pool.apply_async(task_goals, (a,)) for a in jobs]
        for r in res:
            args, out = r.get()
            _save(f"{args[0]}_{args[1]}_goals_s{args[2]}", out)


def main_entity():
    """robot-/ball-specific observations of the three lidar observers and the Stephen assignment."""
    jobs = [("GAME", "RoboRugbySimpleDuel-v2", 91, 2, 40, (0, 1)), ("GAME", "RoboRugbySimpleDuel-v3", 92, 3, 48, (0, 1, 2, 3)),
            ("GAME", "DuelLidar6v1", 93, 3, 48, (0, 1))]
    ctx = mp.get_context("spawn")
    with ctx.Pool(processes=3, maxtasksperchild=1) as pool:
        res = [pool.apply_async(task_entity, (a,)) for a in jobs]
        for r in res:
            args, out = r.get()
            _save(f"{args[0]}_{args[1]}_entity_s{args[2]}", out)


def main_extra():
    """Later additions, generated without touching the files of main(): observer O4 (AllCoords)."""
    jobs = [(task_rollout, ("GAME", "DuelAllCoords", "chase", 51, 2, 48)),
            (task_rollout, ("TRAIN", "DuelAllCoords", "chase", 52, 3, 96))]
    jobs += [(task_reset_fixed, ("GAME", "RoboRugbySimpleDuel-v2", "own", 61, 12)),
             (task_reset_fixed, ("GAME", "RoboRugbySimpleDuel-v2", "standard", 62, 6)),
             (task_reset_fixed, ("TRAIN", "RoboRugbySimpleDuel-v2", "own", 63, 16))]
    ctx = mp.get_context("spawn")
    with ctx.Pool(processes=3, maxtasksperchild=1) as pool:
        for fn, a in jobs:
            args, out = pool.apply_async(fn, (a,)).get()
            if fn is task_reset_fixed:
                _save(f"{args[0]}_{args[1]}_resetfixed{args[2]}_s{args[3]}", out)
            else:
                _save(f"{args[0]}_{args[1]}_{args[2]}_s{args[3]}", out)


def main_mixins():
    """§8f rank 2: the remaining reward mixins (DontDriveInGoals, KeepMovingGuys, BaseDestruction, PushNegBallsFromGoal)
    in two ad-hoc compositions (ref_harness.MIXIN_COMPOSITIONS), random / chase / partially driven rollouts."""
    jobs = [("GAME", "DuelAllMixins", "chase", 71, 3, 64), ("GAME", "DuelAllMixins", "partial", 72, 3, 48),
            ("GAME", "DuelCutChain", "partial", 73, 3, 48), ("TRAIN", "DuelAllMixins", "random", 74, 6, 96),
            ("TRAIN", "DuelCutChain", "chase", 75, 4, 96)]
    if len(sys.argv) > 2 and sys.argv[2] == "prior":   # the AllCoords_WithPrior observer, added afterwards
        jobs = [("GAME", "DuelAllCoordsPrior", "chase", 81, 2, 48), ("TRAIN", "DuelAllCoordsPrior", "random", 82, 3, 96)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(processes=5, maxtasksperchild=1) as pool:
        res = [pool.apply_async(task_rollout, (a,)) for a in jobs]
        for r in res:
            args, out = r.get()
            _save(f"{args[0]}_{args[1]}_{args[2]}_s{args[3]}", out)


def main():
    v0, v2, v3, full = "RoboRugbySimple-v0", "RoboRugbySimpleDuel-v2", "RoboRugbySimpleDuel-v3", "RoboRugby-v0"
    jobs = []
    # (preset, env, kind, seed, n_traj, T)
    for a in [("GAME", v2, "random", 1, 3, 48), ("GAME", v2, "chase", 2, 6, 96), ("GAME", v2, "chase", 3, 6, 96),
              ("GAME", v2, "sticky", 4, 3, 96), ("GAME", v3, "chase", 5, 2, 64), ("GAME", v0, "chase", 6, 2, 64),
              ("GAME", full, "random", 7, 2, 64),
              ("TRAIN", v2, "random", 11, 6, 128), ("TRAIN", v2, "chase", 12, 12, 128), ("TRAIN", v2, "sticky", 13, 6, 128),
              ("TRAIN", v3, "chase", 14, 4, 128), ("TRAIN", v0, "chase", 15, 4, 128), ("TRAIN", full, "random", 16, 2, 128)]:
        jobs.append((task_rollout, a))
    for a in [("GAME", v2, 21, 320), ("GAME", v2, 22, 320), ("GAME", v2, 23, 320), ("GAME", v3, 24, 128), ("GAME", v0, 25, 96),
              ("TRAIN", v2, 31, 768), ("TRAIN", v3, 32, 256)]:
        jobs.append((task_inject, a))
    for a in [("GAME", v2, 41, 48), ("TRAIN", v2, 42, 64)]:
        jobs.append((task_reset, a))
    for a in [("GAME", v2), ("TRAIN", v2)]:
        jobs.append((task_timelimit, a))
    ctx = mp.get_context("spawn")
    t0 = time.time()
    with ctx.Pool(processes=min(8, os.cpu_count()), maxtasksperchild=1) as pool:
        res = [(fn.__name__, pool.apply_async(fn, (a,))) for fn, a in jobs]
        for fname, r in res:
            args, out = r.get()
            tag = "_".join(str(x) for x in args[:4] if not isinstance(x, int) or fname != "task_rollout" or True)
            kind = fname.replace("task_", "")
            if kind == "rollout":
                name = f"{args[0]}_{args[1]}_{args[2]}_s{args[3]}"
            elif kind == "timelimit":
                name = f"{args[0]}_{args[1]}_timelimit"
            else:
                name = f"{args[0]}_{args[1]}_{kind}_s{args[2]}"
            _save(name, out)
    print(f"total {time.time() - t0:.0f}s")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        main_extra()
    elif len(sys.argv) > 1 and sys.argv[1] == "mixins":
        main_mixins()
    elif len(sys.argv) > 1 and sys.argv[1] == "entity":
        main_entity()
    elif len(sys.argv) > 1 and sys.argv[1] == "goals":
        main_goals()
    else:
        main()
