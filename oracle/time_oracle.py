#!/usr/bin/env python
"""Time the C oracle (oracle/rr_oracle.c) on the local host cores: random-action rollouts,
one process per core.  TEST/BENCH INFRASTRUCTURE ONLY.  Prints one JSON line.

    python oracle/time_oracle.py GAME RoboRugbySimpleDuel-v2 3000 8
"""
import json
import multiprocessing as mp
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from oracle import rr_oracle
    preset, env_id, n, cores = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rr_oracle.build()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(rr_oracle._rollout_worker, [(preset, env_id, n, 1000 + i) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    print(json.dumps({"steps_per_s": steps / wall, "wall_s": wall, "cores": cores, "steps": steps}))


if __name__ == "__main__":
    main()
