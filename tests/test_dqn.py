"""VecDQNAgent (roborugby_b200/dqn.py) against a plain restatement of the reference's learn()
(Training_DQN_pytorch.py:143-189) on CPU tensors, plus replay-ring and schedule semantics."""
import copy

import pytest
import torch

from roborugby_b200.dqn import VecDQNAgent


def _agent(**kw):
    kw.setdefault("device", "cpu")
    kw.setdefault("batch_size", 8)
    kw.setdefault("max_mem_size", 32)
    return VecDQNAgent(5, **kw)


def test_replay_ring_wraps_like_the_reference():
    ag = _agent(max_mem_size=10)
    for start in (0, 7):  # second store wraps around the end
        n = 7
        s = torch.arange(n * 5, dtype=torch.float32).reshape(n, 5) + 100 * start
        ag.store(s, torch.arange(n) % 8, torch.arange(n, dtype=torch.float32), s + 1, torch.arange(n) % 2 == 0)
    assert ag.mem_cntr == 14
    # transitions 7..13 went to slots 7,8,9,0,1,2,3 (i = mem_cntr % mem_size, :122)
    assert ag.state_memory[0, 0] == 700 + 3 * 5 and ag.state_memory[9, 0] == 700 + 2 * 5
    assert ag.state_memory[4, 0] == 4 * 5  # untouched slot from the first batch
    assert ag.terminal_memory[0].item() is False and ag.terminal_memory[9].item() is True


def test_learn_matches_reference_update():
    torch.manual_seed(3)
    ag = _agent(gamma=0.9, lr=1e-2, epsilon=1.0, eps_dec=0.5, eps_end=0.3, target_update_freq=16)
    n = 8
    s, s_ = torch.randn(n, 5), torch.randn(n, 5)
    a = torch.randint(0, 8, (n,)); r = torch.randn(n); t = torch.tensor([0, 1, 0, 0, 1, 0, 0, 0], dtype=torch.bool)
    ag.store(s, a, r, s_, t)
    # restatement of :149-183 on the same (whole) batch
    ref_eval, ref_target = copy.deepcopy(ag.Q_eval), copy.deepcopy(ag.Q_target)
    opt = torch.optim.Adam(ref_eval.parameters(), lr=1e-2)
    opt.zero_grad()
    q_eval = ref_eval(s)[torch.arange(n), a]
    q_next = ref_target(s_).detach()
    q_next[t] = 0.0
    q_target = r + 0.9 * torch.max(q_next, dim=1)[0]
    loss = torch.nn.MSELoss()(q_target, q_eval)
    loss.backward()
    opt.step()
    got = ag.learn()  # batch_size == stored transitions: randperm selects all of them
    assert torch.allclose(got, loss.detach(), rtol=1e-6, atol=1e-7)
    # Adam's first step is lr * sign-like(grad): entries whose gradient is ~0 can differ by 2*lr when the batch is
    # summed in another order (randperm), so require agreement on (almost) all entries rather than on every one
    for p, q in zip(ag.Q_eval.parameters(), ref_eval.parameters()):
        close = torch.isclose(p, q, rtol=1e-4, atol=1e-6)
        assert close.float().mean() > 0.99 and (p - q).abs().max() <= 2.5e-2
    assert ag.epsilon == 0.5
    # target copy only when the transition counter crosses a multiple of target_update_freq (:185-187)
    before = copy.deepcopy(ag.Q_target.state_dict())
    ag.learn()
    assert all(torch.equal(before[k], v) for k, v in ag.Q_target.state_dict().items())
    ag.store(s, a, r, s_, t)  # mem_cntr 8 -> 16: crosses
    ag.learn()
    assert all(torch.equal(v, ag.Q_eval.state_dict()[k]) for k, v in ag.Q_target.state_dict().items())
    assert ag.epsilon == 0.3  # floor (:189)


def test_no_learning_before_batch_is_full_and_eps_greedy():
    ag = _agent(batch_size=16, epsilon=0.0)
    ag.store(torch.zeros(4, 5), torch.zeros(4), torch.zeros(4), torch.zeros(4, 5), torch.zeros(4, dtype=torch.bool))
    assert ag.learn() is None  # :144-147
    obs = torch.randn(64, 5)
    greedy = torch.argmax(ag.Q_eval(obs), dim=1).to(torch.uint8)
    assert torch.equal(ag.choose_actions(obs), greedy)
    rand = ag.choose_actions(obs, epsilon_override=1.0)
    assert rand.dtype == torch.uint8 and int(rand.max()) < 8 and len(torch.unique(rand)) > 1


@pytest.mark.gpu
def test_dqn_loop_runs_on_gpu_vec_env():
    from roborugby_b200.dqn import train
    from roborugby_b200.vec_env import RoboRugbyVecEnv
    env = RoboRugbyVecEnv("RoboRugbySimpleDuel-v2", 512, preset="TRAIN", device="cuda:0", seed=2, n_actions=1)
    ag = VecDQNAgent(env.obs_dim, batch_size=1024, max_mem_size=100000, device="cuda:0")
    out = train(env, ag, 40, log_every=10)
    assert out["env_steps"] == 40 * 512 and out["env_steps_per_s"] > 0
    assert ag.mem_cntr == 40 * 512 and all(l == l for l in out["losses"])  # finite
    assert out["stats"]["steps"] == 40 * 512


@pytest.mark.gpu
@pytest.mark.parametrize("drive_grumpy", [False, True])
def test_dqn_replay_holds_both_teams_transitions_of_the_oracle(oracle, drive_grumpy):
    """Training_DQN_pytorch.py:333-353 on the GAME preset (which has grumpy robots): per step the replay receives the
    happy transition and then the grumpy one (its own observation, the action chosen for it, info.dblGrumpyScore,
    info.adblGrumpyState, the same done).  The stored rows of a few envs are replayed on the oracle from the env's
    own initial state with the stored actions."""
    import numpy as np
    from roborugby_b200.dqn import train
    from roborugby_b200.vec_env import RoboRugbyVecEnv
    V2, N, T, seed = "RoboRugbySimpleDuel-v2", 48, 6, 21
    env = RoboRugbyVecEnv(V2, N, preset="GAME", device="cuda:0", seed=seed, n_actions=1)
    ag = VecDQNAgent(env.obs_dim, batch_size=4 * N, max_mem_size=4 * N * T, epsilon=1.0, eps_end=1.0, device="cuda:0", seed=3)
    out = train(env, ag, T, drive_grumpy=drive_grumpy)
    assert ag.mem_cntr == 2 * N * T and out["transitions"] == 2 * N * T
    S, S_, A = ag.state_memory.cpu().numpy(), ag.new_state_memory.cpu().numpy(), ag.action_memory.cpu().numpy()
    Rw, Tm = ag.reward_memory.cpu().numpy(), ag.terminal_memory.cpu().numpy()
    # the random middle-robot actions of drive_grumpy come from the agent's generator: recover them from the env is not
    # possible, so that mode is checked on the rows that do not depend on them (observations before the first step,
    # layout, dones) and the faithful mode on everything
    oracle.scratch_mode(1)
    try:
        for i in range(0, N, 5):
            o = oracle.OracleEnv("GAME", V2, time_limit=True)
            o.reset_philox(seed, i, 0)     # rr_create
            o.reset_philox(seed, i, 1)     # env.reset() in train()
            for t in range(T):
                h, g = 2 * N * t + i, 2 * N * t + N + i
                oh, og = o.observe(1), o.observe(-1)
                if t == 0 or not drive_grumpy:
                    assert np.allclose(S[h], oh.astype(np.float32), rtol=1e-6, atol=1e-6), (i, t)
                    assert np.allclose(S[g], og.astype(np.float32), rtol=1e-6, atol=1e-6), (i, t)
                if drive_grumpy:
                    break
                r = o.step([A[h]])          # only the happy action is applied (:343)
                assert Tm[h] == Tm[g] == bool(r["done"])
                assert np.allclose(Rw[h], r["rew"][0], rtol=1e-5, atol=1e-6) and np.allclose(Rw[g], r["rew"][1], rtol=1e-5, atol=1e-6)
                assert np.allclose(S_[h], r["obs_h"].astype(np.float32), rtol=1e-6, atol=1e-6), (i, t)
                assert np.allclose(S_[g], r["obs_g"].astype(np.float32), rtol=1e-6, atol=1e-6), (i, t)
                assert 0 <= A[g] < 8
    finally:
        oracle.scratch_mode(0)
    if drive_grumpy:   # the grumpy robot really moves: its thrust is the commanded direction's
        st = env.get_state()
        assert (st["rflag"][:, env.preset.num_robots_happy, :2] != 0).any()
    else:              # reference behaviour: nothing but robot 0 is ever commanded
        assert (env.get_state()["rflag"][:, 1:, :2] == 0).all()


def test_og_twitchy_action_distribution():
    """Batched OG_Twitchy (RR_Players.py:14-30): 5 % left, 45 % forward, 45 % back, 5 % right, as GameEnv_Simple ids."""
    import torch
    from roborugby_b200.players import og_twitchy_actions
    g = torch.Generator().manual_seed(0)
    a = og_twitchy_actions((200000, 2), device="cpu", generator=g)
    assert a.dtype == torch.uint8 and a.shape == (200000, 2)
    frac = torch.bincount(a.flatten().long(), minlength=8).double() / a.numel()
    # ids: 0 forward (1,1), 1 back (-1,-1), 2 left (-1,1), 3 right (1,-1)   (RR_EnvBase.py:593-602)
    assert abs(frac[2] - 0.05) < 0.004 and abs(frac[0] - 0.45) < 0.006
    assert abs(frac[1] - 0.45) < 0.006 and abs(frac[3] - 0.05) < 0.004 and frac[4:].sum() == 0
