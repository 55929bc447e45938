#!/usr/bin/env python
"""BASELINE configs[0]: make_snake(num_envs=1, 20x20, 4 snakes, length 3, vision 5) driven exactly like the
reference's scripts drive it (Python list of actions in, NumPy observation out, reset() when all snakes are
done) -- the latency-bound drop-in mode.  Prints env-steps/s incl. resets; the reference's own CPU path does
~2 200 env-steps/s step-only and ~830 incl. resets on one core."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_snake_b200 import make_snake  # noqa: E402

env, _, _, props = make_snake(num_envs=1, num_snakes=4, height=20, width=20, snake_length=3, vision_range=5)
rng = np.random.RandomState(0)
obs = env.reset()
steps, resets = int(os.environ.get('BENCH_STEPS', 20000)), 0
acts = rng.randint(0, 3, size=(steps, 4)).tolist()
for a in acts[:200]:
    obs, rew, done, info = env.step(a)
    if all(done):
        obs = env.reset()
t0 = time.perf_counter()
for a in acts:
    obs, rew, done, info = env.step(a)
    if all(done):
        obs = env.reset()
        resets += 1
dt = time.perf_counter() - t0
print(json.dumps(dict(config='cfg1 (num_envs=1 drop-in)', env_steps_per_sec=steps / dt, agent_steps_per_sec=4 * steps / dt,
                      us_per_step=1e6 * dt / steps, steps=steps, resets=resets)))
