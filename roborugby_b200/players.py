"""Scripted opponents as batched action samplers (reference: robo_rugby/gym_env/RR_Players.py).

The reference asks one player object per robot for a (left, right) thrust pair every step on the host; here a whole batch
of robots gets its discrete action ids (GameEnv_Simple's table, RR_EnvBase.py:593-602) in one torch op on the env's device,
so a rollout with scripted opponents needs no host round trip."""
import torch

# GameEnv_Simple.thrust_from_direction ids of the four thrust pairs OG_Twitchy uses
_LEFT, _FORWARD, _BACKWARD, _RIGHT = 2, 0, 1, 3


def og_twitchy_actions(shape, device="cuda:0", generator=None):
    """OG_Twitchy.get_action (RR_Players.py:14-30) for `shape` robots: u = random();
    u <= 0.05 -> (-1, 1) turn left, u <= 0.5 -> (1, 1) straight, u < 0.95 -> (-1, -1) back, else (1, -1) turn right.
    Returns uint8 action ids usable as (part of) RoboRugbyVecEnv.step / step_k actions."""
    u = torch.rand(shape, device=device, generator=generator, dtype=torch.float64)
    a = torch.full(shape, _RIGHT, dtype=torch.uint8, device=device)
    a = torch.where(u < 0.95, torch.full_like(a, _BACKWARD), a)
    a = torch.where(u <= 0.5, torch.full_like(a, _FORWARD), a)
    a = torch.where(u <= 0.05, torch.full_like(a, _LEFT), a)
    return a


# GameEnv_Simple._dct_thrust_from_direction (RR_EnvBase.py:593-602) as a tensor table: action id -> (left, right)
_THRUST_TABLE = ((1, 1), (-1, -1), (-1, 1), (1, -1), (0, 1), (1, 0), (-1, 0), (0, -1))


def stephen_thrusts(env, robots, choose_action, epsilon=0.2):
    """The "Stephen" players (DQN_pytorch_player.py:9-92) for a whole batch, on the device.

    env: RoboRugbyVecEnv with one of the 6-way lidar observers; robots: the robot indices the players drive;
    choose_action(obs [N, D], epsilon_override=...) -> action ids [N] (e.g. VecDQNAgent.choose_actions: the reference
    loads a pickled DQNAgent that is not part of its tree, so the network is the caller's).
    Per step, as Stephen.__consult does (:63-72): one greedy nearest-ball assignment for all players
    (rr_assign_balls), each player's observation of ITS robot and ITS ball (rr_observe_entity), an eps-greedy action
    (epsilon_override=0.2), translated to a thrust pair; a player without a ball returns (0, 0).
    Returns float32 thrusts [N, len(robots), 2] and the assignment int32 [N, len(robots)]."""
    asg = env.assign_balls(robots)
    table = torch.tensor(_THRUST_TABLE, dtype=torch.float32, device=env.device)
    out = torch.zeros(env.num_envs, len(robots), 2, dtype=torch.float32, device=env.device)
    for j, r in enumerate(robots):
        has = asg[:, j] >= 0
        obs = torch.nan_to_num(env.observe_entity(r, asg[:, j]).float(), nan=0.0)
        act = choose_action(obs, epsilon_override=epsilon).long()
        out[:, j] = torch.where(has.unsqueeze(1), table[act], torch.zeros_like(table[act]))
    return out, asg
