"""CPU timing of the oracle's Python restatement (TEST/BENCH INFRASTRUCTURE).

Used only by bench.py's `cpu_baseline` leg and its `--impl reference` arm: the reference is pure
Python and cannot travel to the GPU box, so its stand-in is oracle/snake_oracle.py (kind "port"), run
the way the reference vectorises -- one OS process per environment (wrappers.py:211-212), reset on
all(done) (wrappers.py:141-143), random actions, module-global NumPy RNG.
"""
import multiprocessing as mp
import os
import time

import numpy as np


def usable_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _worker(rank, kw, conn):
    from oracle.snake_oracle import OracleSnakeEnv
    np.random.seed(1000 + rank)
    env = OracleSnakeEnv(**kw)
    ns = env.num_snakes
    env.reset()
    act = np.random.RandomState(rank).randint(0, 3, size=(4096, ns))
    t = 0
    while True:
        msg = conn.recv()
        if msg is None:
            break
        n_steps = msg
        t0 = time.perf_counter()
        for _ in range(n_steps):
            _, _, done, _ = env.step([int(a) for a in act[t & 4095]])
            t += 1
            if all(done):
                env.reset()
        conn.send(time.perf_counter() - t0)
    conn.close()


class CpuPool:
    """P processes, one oracle env each; run(n) advances every env by n steps and returns wall seconds."""

    def __init__(self, kw, procs=None):
        self.procs = procs or usable_cores()
        ctx = mp.get_context('spawn')
        self.pipes, self.ps = [], []
        for r in range(self.procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(r, kw, b), daemon=True)
            p.start()
            self.pipes.append(a)
            self.ps.append(p)
        self.num_snakes = kw.get('num_snakes', 4)
        self.run(1)                      # wait until every worker has imported and reset

    def run(self, n_steps):
        t0 = time.perf_counter()
        for c in self.pipes:
            c.send(n_steps)
        for c in self.pipes:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        for c in self.pipes:
            c.send(None)
        for p in self.ps:
            p.join(timeout=5)
