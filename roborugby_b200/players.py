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
