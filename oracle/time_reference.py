#!/usr/bin/env python
"""Time the UNMODIFIED Python reference (read-only /root/reference, stub pygame/gym) on this machine.

TEST INFRASTRUCTURE ONLY; runs in the build container only.  One process per core, each stepping its
own env with uniformly random discrete actions (resets on done or on a reference exception).
    python oracle/time_reference.py GAME RoboRugbySimpleDuel-v2 150 8
Prints one JSON line.  The numbers are quoted in DESIGN.md / BASELINE notes; bench.py cannot run this
on the GPU box because the reference tree does not travel.
"""
import json
import multiprocessing as mp
import os
import random
import signal
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(args):
    preset, env_id, n, seed = args
    import contextlib
    import io
    import ref_harness as H
    with contextlib.redirect_stdout(io.StringIO()):
        const = H.load_reference(preset)
        env = H.make_env(env_id, through_gym=True)
    random.seed(seed)
    n_act = const.NUM_ROBOTS_TOTAL

    def alarm(sig, frm):
        raise TimeoutError

    signal.signal(signal.SIGALRM, alarm)
    env.reset()
    done_steps, t0 = 0, None
    for i in range(n + 20):
        if i == 20:
            t0 = time.perf_counter()
        signal.alarm(30)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                _, _, d, _ = env.step([random.randrange(8) for _ in range(n_act)])
        except Exception:
            d = True
        finally:
            signal.alarm(0)
        if d:
            env.reset()
        if i >= 20:
            done_steps += 1
    return done_steps, time.perf_counter() - t0


def main():
    preset, env_id, n, cores = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_worker, [(preset, env_id, n, 100 + i) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    per_core = [r[0] / r[1] for r in res]
    print(json.dumps({"preset": preset, "env_id": env_id, "cores": cores, "steps_per_s": steps / wall,
                      "per_core_mean": sum(per_core) / len(per_core), "wall_s": wall}))


if __name__ == "__main__":
    main()
