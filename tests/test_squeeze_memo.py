"""CPU tier: the squeeze memo of the kernel source (rr_sim.cuh squeeze_contacts), run on the host by tests/emul.

A replayed frame must be indistinguishable from a recomputed one: the same trajectories are run with the memo on and
with RR_FLAG_NO_SQUEEZE_MEMO, and against the oracle, on states that spend most of their frames pinned (the oracle's
failed-frame counter proves that the scenarios exercise the path, the emulator's replay counter that the memo fires).
Also pins the relaxed reset placement (strict_reset = 0) of the kernels against the oracle's restatement of it."""
import numpy as np
import pytest

from squeeze_util import V2, actions, oracle_env, pincer_env, scenario


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
