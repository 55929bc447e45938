#!/usr/bin/env python
"""Single-thread CPU context number for bench.py: the device rule source compiled for the host
(tests/hostsim, TEST INFRASTRUCTURE) stepping 256 cfg5-shaped envs.  Prints one JSON line.  Run as a
subprocess of bench.py so that no test code is loaded into the measuring process."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

ENV_KW = dict(height=20, width=20, num_snakes=4, snake_length=3, vision_range=5, frame_stack=1)


def main(seconds=4.0, n_envs=256):
    from hostsim_util import HostSim
    hs = HostSim(n_envs, ENV_KW, rng_mode=0, auto_reset=1, seed=1)
    hs.reset()
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 3, size=(64, n_envs, ENV_KW['num_snakes'])).astype(np.uint8)
    for t in range(8):
        hs.step(acts[t])
    t0, steps = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        hs.step(acts[steps % 64])
        steps += 1
    wall = time.perf_counter() - t0
    print(json.dumps({'value': steps * n_envs * ENV_KW['num_snakes'] / wall, 'unit': 'agent-steps/s', 'cores': 1,
                      'kind': 'native C++ host build of the rule source (tests/hostsim), incl. NumPy marshalling',
                      'sample': f'{n_envs} envs x {steps} steps, {wall:.1f} s'}))


if __name__ == '__main__':
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 4.0)
