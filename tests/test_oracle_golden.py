"""Pins the CPU oracle (oracle/rr_oracle.c) to the reference's own outputs, bit for bit.

The golden files were produced by oracle/gen_golden.py from the unmodified reference.  Every
record is replayed through the oracle twice: (a) as a trajectory from the first state only and
(b) step by step from each recorded state.  State, observations, rewards, done and the naughty
count must be IDENTICAL (np.array_equal on float64; NaN == NaN for absent observations).
"""
import numpy as np
import pytest

from conftest import golden_files
from golden_util import actions_at, apply_overrides, parse_name, resolve_env, state_at, state_diff, states_equal

ROLLOUTS = [f for f in golden_files("*.npz") if parse_name(f)[2] in ("random", "chase", "sticky", "inject", "partial", "entity")]


def _eq(a, b):
    return np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64), equal_nan=True)


@pytest.mark.parametrize("path", ROLLOUTS, ids=[p.split("/")[-1] for p in ROLLOUTS])
def test_oracle_matches_reference_bitwise(oracle, path):
    preset, env_id, kind = parse_name(path)
    d = np.load(path)
    n, T = d["act"].shape[:2]
    base_id, observer = resolve_env(env_id)
    cfg = apply_overrides(oracle.default_config(preset, base_id), env_id)
    env = oracle.OracleEnv(cfg=cfg)
    oracle.scratch_mode(0)
    bad = []
    for mode in ("trajectory", "per_step"):
        for i in range(n):
            env.set_state(state_at(d, i, 0))
            for t in range(T):
                if mode == "per_step" or d["restart"][i, t]:
                    env.set_state(state_at(d, i, t))
                oracle.scratch_reset()
                out = env.step(actions_at(d, i, t))
                got = env.get_state()
                want = state_at(d, i, t + 1)
                exc = bool(d["exc"][i, t])
                ok = states_equal(got, want) if not exc else True
                ok &= bool(out["err"] != 0) == exc
                if not exc:
                    ok &= _eq(out["obs_h"], d["obs_h"][i, t]) and _eq(out["obs_g"], d["obs_g"][i, t])
                    ok &= _eq(out["rew"], d["rew"][i, t])
                    ok &= out["done"] == int(d["done"][i, t]) and out["naughty"] == int(d["naughty"][i, t])
                if not ok:
                    bad.append((mode, i, t, state_diff(got, want), out["err"], exc))
                    break
    assert not bad, f"{len(bad)} mismatching records, first: {bad[:3]}"


RESETS = golden_files("*_reset_*.npz")


@pytest.mark.parametrize("path", RESETS, ids=[p.split("/")[-1] for p in RESETS])
def test_oracle_reset_matches_reference(oracle, path):
    preset, env_id, _ = parse_name(path)
    d = np.load(path)
    env = oracle.OracleEnv(preset, env_id)
    for i in range(d["draws"].shape[0]):
        env.set_state(state_at(d, i, 0))
        nd = int(d["ndraws"][i])
        used = env.reset_draws(d["draws"][i, :nd])
        assert used == nd
        got, want = env.get_state(), state_at(d, i, 1)
        assert states_equal(got, want), (i, state_diff(got, want))
        assert _eq(env.observe(1), d["obs_h"][i, 0])
        assert _eq(env.observe(-1), d["obs_g"][i, 0])


TL = golden_files("*_timelimit.npz")


@pytest.mark.parametrize("path", TL, ids=[p.split("/")[-1] for p in TL])
def test_oracle_done_semantics(oracle, path):
    preset, env_id, _ = parse_name(path)
    d = np.load(path)
    T, s0 = int(d["T"]), int(d["start_step"])
    for time_limit, want in ((True, d["wrapped_done"]), (False, d["raw_done"])):
        env = oracle.OracleEnv(preset, env_id, time_limit=time_limit)
        env.reset_philox(1, 0, 0)
        st = env.get_state(); st["step"] = np.int32(s0)
        env.set_state(st)
        got = [env.step([0] )["done"] for _ in range(len(want))]
        assert got == [int(x) for x in want]
    # raw env: the step after done raises (RR_EnvBase.py:261-262)
    assert bool(d["raised"]) and env.step([0])["err"] & 1
    assert T == env.cfg.game_length_steps


FIXED = golden_files("*_resetfixed*.npz")


@pytest.mark.parametrize("path", FIXED, ids=[p.split("/")[-1] for p in FIXED])
def test_oracle_fixed_layout_reset_matches_reference(oracle, path):
    """reset(bln_randomize_pos=False) / GameEnv(CONFIG_STANDARD) (RR_EnvBase.py:131-153, :202-216)."""
    preset, env_id, _ = parse_name(path)
    d = np.load(path)
    R, B = d["rob"].shape[2], d["ball"].shape[2]
    for i in range(d["start"].shape[0]):
        env = oracle.OracleEnv(preset, env_id)  # freshly constructed sprites at the origin
        if not d["restart"][i, 0]:
            env.set_state(state_at(d, i, 0))
        env.set_starting_positions(d["start"][i, :3 * R].reshape(R, 3), d["start"][i, 3 * R:].reshape(B, 2))
        assert env.reset_draws([], randomize=False) == 0
        got, want = env.get_state(), state_at(d, i, 1)
        assert states_equal(got, want), (i, state_diff(got, want))
        assert _eq(env.observe(1), d["obs_h"][i, 0]) and _eq(env.observe(-1), d["obs_g"][i, 0])


ENTITY = golden_files("*_entity_*.npz")


@pytest.mark.parametrize("path", ENTITY, ids=[p.split("/")[-1] for p in ENTITY])
def test_oracle_entity_observations_and_stephen_assignment(oracle, path):
    """get_game_state(obj_robot=..., obj_ball=...) of every robot and ball (RR_Observers.py:133-136, :187-203, :304-320)
    and the "Stephen" players' greedy nearest-ball assignment (DQN_pytorch_player.py:39-61), both recorded from the
    reference after every step of a chase rollout: bit for bit."""
    preset, env_id, _ = parse_name(path)
    d = np.load(path)
    n, T, R, B = d["ent"].shape[:4]
    base_id, observer = resolve_env(env_id)
    cfg = apply_overrides(oracle.default_config(preset, base_id), env_id)
    env = oracle.OracleEnv(cfg=cfg)
    hive = d["hive"]
    assigned = 0
    for i in range(n):
        for t in range(T):
            env.set_state(state_at(d, i, t + 1))
            for r in range(R):
                for b in range(B):
                    assert _eq(env.observe_entity(r, b), d["ent"][i, t, r, b]), (i, t, r, b)
                assert _eq(env.observe_entity(r, -1), d["ent"][i, t, r, 0]), "default ball = lstPosBalls[0]"
            got = env.assign_balls(hive)
            assert np.array_equal(got, d["asg"][i, t]), (i, t, got, d["asg"][i, t])
            assigned += int((got >= 0).sum())
    assert assigned > 0
    if env.cfg.observer in (3, 4):
        with pytest.raises(NotImplementedError):
            env.observe_entity(0, 0)


GOALS = golden_files("*_goals_*.npz")


@pytest.mark.parametrize("path", GOALS, ids=[p.split("/")[-1] for p in GOALS])
def test_oracle_goal_scoring_matches_patched_reference(oracle, path):
    """Goal scoring as intended (cfg.goal_scoring): the reference with its dead scoring code made live by the in-memory
    patches of oracle/ref_harness.py (_GOAL_SCORING_PATCHES: RR_Goal.py:54-91, RR_EnvBase.py:461-511).  Whole
    trajectories from the first state: state, observations, rewards (they carry the +-500 score delta), done, and after
    every step the balls still alive, both goals' scores and destroyed flags and the dwell counters, bit for bit."""
    preset, env_id, _ = parse_name(path)
    d = np.load(path)
    n, T = d["act"].shape[:2]
    base_id, observer = resolve_env(env_id)
    cfg = apply_overrides(oracle.default_config(preset, base_id), env_id)
    cfg.goal_scoring = 1
    env = oracle.OracleEnv(cfg=cfg)
    oracle.scratch_mode(0)
    events = killed = ended = 0
    for i in range(n):
        env.set_state(state_at(d, i, 0))
        for t in range(T):
            if d["exc"][i, t] == 2:      # the reference's game_is_done() was already True: step() raises (:261-262)
                assert env.step(actions_at(d, i, t))["err"] & 1, (i, t)
                ended += 1
                break
            oracle.scratch_reset()
            out = env.step(actions_at(d, i, t))
            assert out["err"] == 0, (i, t, out["err"])
            got, want = env.get_state(), state_at(d, i, t + 1)
            assert states_equal(got, want), (i, t, state_diff(got, want))
            assert _eq(out["obs_h"], d["obs_h"][i, t]) and _eq(out["obs_g"], d["obs_g"][i, t]), (i, t)
            assert _eq(out["rew"], d["rew"][i, t]), (i, t, out["rew"], d["rew"][i, t])
            assert out["done"] == int(d["done"][i, t]) and out["naughty"] == int(d["naughty"][i, t]), (i, t)
            g = env.goal_state()
            assert np.array_equal(g["alive"], d["alive"][i, t + 1]), (i, t)
            assert np.array_equal(g["score"], d["score"][i, t + 1]) and np.array_equal(g["destroyed"], d["destroyed"][i, t + 1])
            assert np.array_equal(g["dwell"], d["dwell"][i, t + 1]), (i, t, g["dwell"], d["dwell"][i, t + 1])
            assert g["delta"] == int(d["delta"][i, t])
            events += g["delta"] != 0
        killed += int((d["alive"][i, -1] == 0).sum())
    print(f"{path.split('/')[-1]}: {events} scoring steps, {killed} balls consumed, {ended} episodes ended by the goals")
    if "SimpleDuel-v2" in path:   # NaughtyBots cuts the on_step_end chain before the goals' own hook: nothing ever scores
        assert events == 0 and killed == 0 and d["dwell"].max() == 0
    else:
        assert events > 0 and killed > 0
