"""CPU tier: the squeeze memo of the kernel source (rr_sim.cuh squeeze_contacts), run on the host by tests/emul.

A replayed frame must be indistinguishable from a recomputed one: the same trajectories are run with the memo on and
with RR_FLAG_NO_SQUEEZE_MEMO, and against the oracle, on states that spend most of their frames pinned (the oracle's
failed-frame counter proves that the scenarios exercise the path, the emulator's replay counter that the memo fires).
Also pins the relaxed reset placement (strict_reset = 0) of the kernels against the oracle's restatement of it."""
import numpy as np
import pytest

from squeeze_util import V2, actions, oracle_env, pincer_env, scenario, stuck_pair_env


def _cfg(preset, flags):
    from roborugby_b200 import _lib
    cfg = _lib.default_config(_lib.PRESET_GAME if preset == "GAME" else _lib.PRESET_TRAIN, V2)
    cfg.time_limit = 0
    cfg.auto_reset = 0
    cfg.flags = flags
    return cfg


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_squeeze_memo_replay_is_exact(oracle, seed):
    from emul import emul
    from emul.emul import EmulEnv
    from roborugby_b200 import _lib
    rng = np.random.default_rng(seed)
    failed = replays = whole = 0
    oracle.scratch_mode(1)
    emul.use_libm_sincos(True)
    try:
        for it in range(60):
            preset = "GAME" if it % 4 else "TRAIN"
            spect = None
            if seed == 2:   # bystanders next to the squeeze: a ball 16-45 px away, sometimes a robot 35-80 px away
                ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(16, 45)
                spect = [(d * np.cos(ang), d * np.sin(ang))]
                if it % 3 == 0:
                    ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(35, 80)
                    spect.append((d * np.cos(ang), d * np.sin(ang)))
            if seed == 3:   # a ball pinned between two robots driving at each other
                preset = "GAME"
                o = pincer_env(oracle, rng)
            else:
                o = oracle_env(oracle, preset, scenario(rng, preset), spectators=spect)
            st0 = o.get_state()
            on = EmulEnv(_cfg(preset, 0), o.R, o.B, 5)
            off = EmulEnv(_cfg(preset, _lib.FLAG_NO_SQUEEZE_MEMO), o.R, o.B, 5)
            on.set_state(st0); off.set_state(st0)
            oracle.failed_frames(True)
            acts = actions(rng, o.R)
            if seed == 3:
                acts[:, 1] = acts[:, 0]   # robot 1 pushes too
            for a in acts:
                r_on = on.step(a)
                replays += emul.lib().emul_last_replays()
                whole += emul.lib().emul_last_whole_frames()
                r_off = off.step(a)
                r_o = o.step(a)
                s_on, s_off, s_o = on.get_state(), off.get_state(), o.get_state()
                for k in s_on:   # memo on == memo off, bit for bit
                    assert np.array_equal(s_on[k], s_off[k]), (it, k)
                assert r_on["err"] == r_off["err"] == r_o["err"]
                if r_on["err"]:   # the reference raised (e.g. a bystander robot placed on top of the pushing one)
                    break
                for k in ("rew", "obs_h", "obs_g"):
                    assert np.array_equal(r_on[k], r_off[k], equal_nan=True), (it, k)
                assert r_on["naughty"] == r_off["naughty"] == r_o["naughty"]
                for k in ("rflag", "step"):
                    assert np.array_equal(s_on[k], s_o[k]), (it, k)
                for k in ("rob", "rhist", "ball"):   # ... and both == the oracle
                    assert np.allclose(s_on[k], s_o[k], rtol=1e-9, atol=1e-9), (it, k)
                assert np.allclose(r_on["rew"], r_o["rew"], rtol=1e-9, atol=1e-9)
            failed += oracle.failed_frames(True)
    finally:
        emul.use_libm_sincos(False)
        oracle.scratch_mode(0)
    print(f"seed {seed}: {failed} pinned-ball frames in the oracle, {int(replays)} replayed by the memo, "
          f"{int(whole)} of them as whole frames")
    assert whole > 0.5 * replays
    assert failed > (1500 if seed < 3 else 200) and replays > (0.7 if seed < 2 else 0.4) * failed


def test_stuck_pair_memo_replay_is_exact(oracle):
    """Two robots driving into each other (both undone every frame, RR_EnvBase.py:303-333): the stuck-pair memo of
    rr_sim.cuh (resolve_bot_collisions / stuck_pair_replay) on, off, and the oracle."""
    from emul import emul
    from emul.emul import EmulEnv
    from roborugby_b200 import _lib
    rng = np.random.default_rng(11)
    replays = naughty = raised = 0
    oracle.scratch_mode(1)
    emul.use_libm_sincos(True)
    try:
        for it in range(80):
            o = stuck_pair_env(oracle, rng)
            st0 = o.get_state()
            on = EmulEnv(_cfg("GAME", 0), o.R, o.B, 5)
            off = EmulEnv(_cfg("GAME", _lib.FLAG_NO_SQUEEZE_MEMO), o.R, o.B, 5)
            on.set_state(st0); off.set_state(st0)
            for s in range(6):
                a = [int(x) for x in rng.integers(0, 8, o.R)]
                if s in (0, 1, 4):
                    a[0] = a[1] = 0          # both forward: head-on
                r_on = on.step(a)
                replays += emul.lib().emul_last_stuck_replays()
                r_off = off.step(a)
                r_o = o.step(a)
                assert r_on["err"] == r_off["err"] == r_o["err"]
                if r_on["err"]:
                    raised += 1
                    break
                s_on, s_off, s_o = on.get_state(), off.get_state(), o.get_state()
                for k in s_on:
                    assert np.array_equal(s_on[k], s_off[k]), (it, s, k)
                assert np.array_equal(r_on["rew"], r_off["rew"]) and r_on["naughty"] == r_off["naughty"] == r_o["naughty"]
                naughty += r_o["naughty"]
                for k in ("rflag", "step"):
                    assert np.array_equal(s_on[k], s_o[k]), (it, s, k)
                for k in ("rob", "rhist", "ball"):
                    assert np.allclose(s_on[k], s_o[k], rtol=1e-9, atol=1e-9), (it, s, k)
                assert np.allclose(r_on["rew"], r_o["rew"], rtol=1e-9, atol=1e-9)
    finally:
        emul.use_libm_sincos(False)
        oracle.scratch_mode(0)
    print(f"stuck pairs: {naughty} naughty robot-steps, {int(replays)} robot-robot phases replayed, {raised} scenarios raised")
    assert replays > 1000 and naughty > 300


def _mixed_scenario(oracle, rng, i, preset, K):
    """The scenario mix of the GPU test (tests/test_parity_gpu.py::test_gpu_squeeze_memo_replay_is_exact)."""
    if preset == "GAME" and i % 4 == 3:
        o = pincer_env(oracle, rng)
        a = actions(rng, o.R, K)
        a[:, 1] = a[:, 0]
    elif preset == "GAME" and i % 4 == 2:
        o = stuck_pair_env(oracle, rng)
        a = actions(rng, o.R, K)
        a[:, 1] = a[:, 0]
    else:
        spect = None
        if i % 3 == 1:
            ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(16, 60)
            spect = [(d * np.cos(ang), d * np.sin(ang))]
        o = oracle_env(oracle, preset, scenario(rng, preset), spectators=spect)
        a = actions(rng, o.R, K)
    return o, a


@pytest.mark.parametrize("preset", ["GAME", "TRAIN"])
def test_memos_persist_across_fused_steps_exactly(oracle, preset):
    """Inside a fused launch the memo state outlives the env-step (slots recorded under one action are hit again when
    the action comes back, whole-frame records are replayed across the step boundary).  emul_step_k runs K steps on
    one Env object like a launch does: memo on == memo off bit for bit, and == the oracle within 1e-9, on the GPU
    test's scenario mix (wall squeezes with and without a bystander ball, two-robot pincers, robot pairs stuck on
    each other).  A bystander that is itself at the wall found a real bug here: its own wall bounce carried it into
    the pushing robot, so whole frames are only replayed when nothing else is in contact."""
    from emul import emul
    from emul.emul import EmulEnv
    from roborugby_b200 import _lib
    rng = np.random.default_rng(5)
    K, n = 6, 700 if preset == "GAME" else 300
    tot = np.zeros(3)
    oracle.scratch_mode(1)
    emul.use_libm_sincos(True)
    try:
        for i in range(n):
            o, a = _mixed_scenario(oracle, rng, i, preset, K)
            st0 = o.get_state()
            on = EmulEnv(_cfg(preset, 0), o.R, o.B, 5)
            off = EmulEnv(_cfg(preset, _lib.FLAG_NO_SQUEEZE_MEMO), o.R, o.B, 5)
            on.set_state(st0); off.set_state(st0)
            e_on, r_on, cnt = on.step_k(a)
            e_off, r_off, cnt_off = off.step_k(a)
            tot += cnt
            assert not cnt_off.any()
            s_on, s_off = on.get_state(), off.get_state()
            assert e_on == e_off and np.array_equal(r_on, r_off), i
            for k in s_on:
                assert np.array_equal(s_on[k], s_off[k]), (i, k)
            if i % 7 == 0 and not e_on:
                for t in range(K):
                    r = o.step(a[t])
                    assert not r["err"] and np.allclose(r["rew"], r_on[t], rtol=1e-9, atol=1e-9), (i, t)
                s_o = o.get_state()
                for k in ("rob", "rhist", "ball"):
                    assert np.allclose(s_on[k], s_o[k], rtol=1e-9, atol=1e-9), (i, k)
    finally:
        emul.use_libm_sincos(False)
        oracle.scratch_mode(0)
    print(f"{preset}: {n} scenarios x {K} fused steps: {int(tot[0])} frames replayed ({int(tot[1])} as whole frames), "
          f"{int(tot[2])} robot-robot phases replayed")
    assert tot[1] > 1000 and (preset == "TRAIN" or tot[2] > 500)


def test_memos_are_exact_on_long_random_rollouts(oracle):
    """Random-action rollouts from the product's own resets (no hand-built contact): 48 GAME envs x 240 steps in
    fused chunks of 24, memo on vs off, bit for bit."""
    from emul.emul import EmulEnv
    from roborugby_b200 import _lib
    rng = np.random.default_rng(9)
    tot = np.zeros(3)
    for i in range(48):
        on = EmulEnv(_cfg("GAME", 0), 4, 8, 5)
        off = EmulEnv(_cfg("GAME", _lib.FLAG_NO_SQUEEZE_MEMO), 4, 8, 5)
        on.reset(100 + i, 0, construct=True)
        off.set_state(on.get_state())
        for chunk in range(10):
            a = rng.integers(0, 8, (24, 4))
            if chunk % 3 == 2:
                a[:, :] = a[:1, :]      # a held action: robots keep driving into whatever they hit
            e_on, r_on, cnt = on.step_k(a)
            e_off, r_off, _ = off.step_k(a)
            tot += cnt
            assert e_on == e_off and np.array_equal(r_on, r_off), (i, chunk)
            s_on, s_off = on.get_state(), off.get_state()
            for k in s_on:
                assert np.array_equal(s_on[k], s_off[k]), (i, chunk, k)
            if e_on:
                break
    print(f"random rollouts: {int(tot[0])} frames replayed ({int(tot[1])} whole), {int(tot[2])} robot-robot phases replayed")
    assert tot[2] > 0


@pytest.mark.parametrize("preset", ["GAME", "TRAIN"])
def test_relaxed_reset_matches_oracle(oracle, preset):
    """strict_reset = 0 (two extra rejection rules in the placement loop, rr_sim.cuh reset_env) against the oracle's
    restatement of that mode; and the two modes really differ somewhere in the sampled streams."""
    from emul.emul import EmulEnv
    differ = 0
    for strict in (1, 0):
        cfg = _cfg(preset, 0)
        cfg.strict_reset = strict
        cfg.seed = 77
        for i in range(400 if preset == "GAME" else 60):
            o = oracle.OracleEnv(preset, V2)
            o.reset_philox(77, 5000 + i, 3, relaxed=not strict)
            ref = o.get_state()
            e = EmulEnv(cfg, o.R, o.B, 5)
            e.reset(5000 + i, 3, construct=True)
            got = e.get_state()
            assert np.array_equal(ref["rob"][:, [0, 1, 6]], got["rob"][:, [0, 1, 6]]), (strict, i)
            assert np.array_equal(ref["ball"], got["ball"]), (strict, i)
            assert np.allclose(ref["rob"], got["rob"], rtol=0, atol=1e-11)
            if not strict:
                o2 = oracle.OracleEnv(preset, V2)
                o2.reset_philox(77, 5000 + i, 3)
                differ += not np.array_equal(o2.get_state()["ball"], ref["ball"]) or \
                    not np.array_equal(o2.get_state()["rob"], ref["rob"])
    print(f"{preset}: relaxed placement differs from the reference's in {differ} of the sampled resets")
