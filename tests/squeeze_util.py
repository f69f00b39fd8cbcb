"""Hand-built "ball pinned between a robot and a wall" states (RR_EnvBase.py:345-454: ten failed resolve passes, then
the undo loop, frame after frame): the configuration that the kernels' squeeze memo (rr_sim.cuh squeeze_contacts)
replays instead of recomputing.  Used by the CPU and GPU tiers."""
import numpy as np

V2 = "RoboRugbySimpleDuel-v2"


def scenario(rng, preset):
    """(ball x, ball y, robot x, robot y, robot heading): robot 0 driving ball 0 into one of the four walls."""
    W = 800 if preset == "GAME" else 600
    y0 = rng.uniform(150, W - 150)
    side = int(rng.integers(0, 4))
    th, off, gap = rng.uniform(-12, 12), rng.uniform(-12, 12), rng.uniform(0.0, 3.0)
    if side == 0:
        sc = (7.3 + gap * 0.2, y0, 7.3 + 17 + gap, y0 + off, 180 + th)
    elif side == 1:
        sc = (W - 7.3 - gap * 0.2, y0, W - 7.3 - 17 - gap, y0 + off, 0 + th)
    elif side == 2:
        sc = (y0, 7.3 + gap * 0.2, y0 + off, 7.3 + 17 + gap, 90 + th)
    else:
        sc = (y0, W - 7.3 - gap * 0.2, y0 + off, W - 7.3 - 17 - gap, 270 + th)
    return sc[:4] + (sc[4] % 360,)


def oracle_env(oracle, preset, sc, time_limit=False, spectators=None):
    """An oracle env reset (through the reference's own fixed-layout reset) to the scenario; the other robots and balls
    sit in the middle of the arena, far from the squeeze, except `spectators` = (dx, dy) offsets from the pinned ball
    at which ball 1 (and robot 1, if a second offset is given) are placed: bystanders at a wall next to the squeeze."""
    bx, by, rx, ry, rot = sc
    o = oracle.OracleEnv(preset, V2, time_limit=time_limit)
    W = 800 if preset == "GAME" else 600
    rob3 = [(rx, ry, rot)] + [(W / 2 + 60 * (i - 1), W / 2 + 90 * (i - 2), 45.0 * i) for i in range(1, o.R)]
    ball2 = [(bx, by)] + [(W / 2 - 150 + 40 * i, W / 2 + 200 - 30 * i) for i in range(1, o.B)]
    if spectators and o.B > 1:
        clip = lambda v: float(min(max(v, 8.0), W - 8.0))
        ball2[1] = (clip(bx + spectators[0][0]), clip(by + spectators[0][1]))
        if len(spectators) > 1 and o.R > 1:
            rob3[1] = (float(min(max(bx + spectators[1][0], 60.0), W - 60.0)), float(min(max(by + spectators[1][1], 60.0), W - 60.0)), 30.0)
    o.set_starting_positions(np.array(rob3, float), np.array(ball2, float))
    o.reset_draws([], randomize=False)
    return o


def pincer_env(oracle, rng, time_limit=False):
    """GAME preset: ball 0 in the open between robots 0 and 1, which drive at each other (a squeeze without a wall)."""
    o = oracle.OracleEnv("GAME", V2, time_limit=time_limit)
    W = 800
    x0, y0 = rng.uniform(200, 600), rng.uniform(200, 600)
    ang = rng.uniform(0, 360)
    ca, sa = np.cos(np.radians(ang)), -np.sin(np.radians(ang))      # heading `ang` drives along (cos, -sin)
    g0, g1 = rng.uniform(17.5, 20), rng.uniform(17.5, 20)
    j0, j1 = rng.uniform(-8, 8), rng.uniform(-8, 8)                 # sideways offsets, degrees of misalignment
    rob3 = [(x0 - g0 * ca - j0 * sa, y0 - g0 * sa + j0 * ca, (ang + rng.uniform(-6, 6)) % 360),
            (x0 + g1 * ca - j1 * sa, y0 + g1 * sa + j1 * ca, (ang + 180 + rng.uniform(-6, 6)) % 360)]
    rob3 += [(100.0 + 600 * (x0 < 400), 100.0, 45.0), (100.0 + 600 * (x0 < 400), 700.0, 135.0)]
    ball2 = [(x0, y0)] + [(60.0 + 90 * i, 40.0 + 720 * (y0 < 400)) for i in range(1, o.B)]
    o.set_starting_positions(np.array(rob3, float), np.array(ball2, float))
    o.reset_draws([], randomize=False)
    return o


def stuck_pair_env(oracle, rng, time_limit=False):
    """GAME preset: robots 0 and 1 a few pixels apart, facing each other (or at an angle), no ball near: driving at
    each other they collide and are both undone, frame after frame (RR_EnvBase.py:303-333)."""
    o = oracle.OracleEnv("GAME", V2, time_limit=time_limit)
    x0, y0 = rng.uniform(150, 650), rng.uniform(150, 650)
    ang = rng.uniform(0, 360)
    ca, sa = np.cos(np.radians(ang)), -np.sin(np.radians(ang))
    g = rng.uniform(10.5, 14)                                       # half the centre distance (robots are 20 long)
    j = rng.uniform(-15, 15)
    rob3 = [(x0 - g * ca, y0 - g * sa, (ang + rng.uniform(-25, 25)) % 360),
            (x0 + g * ca - j * sa, y0 + g * sa + j * ca, (ang + 180 + rng.uniform(-25, 25)) % 360),
            (60.0 + 680 * (x0 < 400), 70.0, 45.0), (60.0 + 680 * (x0 < 400), 730.0, 135.0)]
    ball2 = [(50.0 + 95 * i, 45.0 + 710 * (y0 < 400)) for i in range(o.B)]
    o.set_starting_positions(np.array(rob3, float), np.array(ball2, float))
    o.reset_draws([], randomize=False)
    return o


def actions(rng, n_robots, n_steps=6):
    """Robot 0 keeps pushing forward (with one random action in between); the others move at random."""
    lead = [0, 0, 0, int(rng.integers(0, 8)), 0, 0][:n_steps]
    return np.array([[a0] + [int(x) for x in rng.integers(0, 8, n_robots - 1)] for a0 in lead], np.uint8)
