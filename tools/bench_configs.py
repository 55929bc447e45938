#!/usr/bin/env python
"""Device-timed throughput of every BASELINE.json config shape on one GPU (not the headline bench line;
bench.py measures cfg5).  Prints one JSON object per config: agent-steps/s, ms/step, achieved GB/s of
algorithmic traffic, fraction of the measured HBM peak; cfg2 (launch-latency bound) is also timed with the
steps captured in a CUDA graph."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_snake_b200 import SnakeBatch  # noqa: E402

WANT_OBS = os.environ.get('BENCH_NO_OBS', '0') != '1'      # BENCH_NO_OBS=1: rules + record traffic only
CFG4_REW = {'fruit': 10.0, 'kill': 1.0, 'lose': -1.0, 'win': 0.1, 'time': -0.001}
CONFIGS = {
    'cfg2': dict(num_envs=4096, height=20, width=20, num_snakes=4, snake_length=3),
    'cfg3': dict(num_envs=65536, height=20, width=20, num_snakes=4, snake_length=3, vision_range=5, frame_stack=4),
    'cfg4': dict(num_envs=16384, height=64, width=64, num_snakes=16, snake_length=5, vision_range=7,
                 reward_dict=CFG4_REW),
    'wide8': dict(num_envs=32768, height=32, width=32, num_snakes=8, snake_length=4, vision_range=7),
    'cfg5_shard': dict(num_envs=131072, height=20, width=20, num_snakes=4, snake_length=3, vision_range=5),
    'cfg5_256k': dict(num_envs=262144, height=20, width=20, num_snakes=4, snake_length=3, vision_range=5),
    'cfg5_512k': dict(num_envs=524288, height=20, width=20, num_snakes=4, snake_length=3, vision_range=5),
    'cfg5_full': dict(num_envs=1048576, height=20, width=20, num_snakes=4, snake_length=3, vision_range=5),
}


def run(name, kw, steps, graph=False):
    kw = dict(kw)
    N = kw.pop('num_envs')
    ns = kw['num_snakes']
    b = SnakeBatch(N, seed=0, **kw)
    g = torch.Generator(device='cuda').manual_seed(1)
    npool = 251
    pool = torch.randint(0, 3, (npool, N, ns), dtype=torch.uint8, device='cuda', generator=g)
    b.reset()
    for t in range(int(os.environ.get('BENCH_BURN', 300))):
        b.step(pool[t % npool], want_info=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        T = 50
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for t in range(3):
                b.step(pool[t], want_info=False)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, stream=s):
                for t in range(T):
                    b.step(pool[t], want_info=False)
        torch.cuda.synchronize()
        reps = max(1, steps // T)
        e0.record()
        for _ in range(reps):
            cg.replay()
        e1.record()
        steps = reps * T
    else:
        e0.record()
        for t in range(steps):
            b.step(pool[(300 + t) % npool], want_info=False, want_obs=WANT_OBS)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    peak = 6549.4
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    bpe = b.algorithmic_bytes_per_env_step()
    gbs = bpe * N * steps / (ms * 1e-3) / 1e9
    out = dict(config=name, graph=graph, num_envs=N, steps=steps, ms_per_step=ms / steps,
               agent_steps_per_sec=N * ns * steps / (ms * 1e-3), bytes_per_agent_step=bpe / ns,
               achieved_gbs=gbs, frac_of_measured_peak=gbs / peak, device_errors=b.device_errors(),
               tile_envs=os.environ.get('SNK_TILE_ENVS', 'auto'), threads=os.environ.get('SNK_THREADS', 'auto'),
               coop=os.environ.get('SNK_COOP', 'auto'))
    print(json.dumps(out), flush=True)
    b.close()


if __name__ == '__main__':
    which = sys.argv[1:] or ['cfg2', 'cfg3', 'cfg4', 'cfg5_shard', 'cfg5_full']
    for name in which:
        run(name, CONFIGS[name], int(os.environ.get('BENCH_STEPS', 400 if name != 'cfg2' else 2000)))
        if name == 'cfg2' and 'BENCH_STEPS' not in os.environ:
            run(name, CONFIGS[name], 2000, graph=True)
