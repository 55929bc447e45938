#!/usr/bin/env python
"""Per-CTA phase timeline of snk_tile_kernel from the profiling build (make -C marl-snake_b200/csrc prof):

    SNK_LIB_PATH=marl-snake_b200/libsnk_prof.so python tools/phase_timing.py cfg4 cfg5_shard cfg2

Every warp of every CTA of ONE launch stamps %globaltimer (ns) at its phase boundaries into a trace buffer (no
atomics).  Prints per config: mean ns per phase for the rule warp and the other warps, CTA life, the launch's
span from first CTA start to last CTA end, and how many CTAs were resident over time.  frame_stack 1 only."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bench_configs as bc  # noqa: E402
import marl_snake_b200 as m  # noqa: E402

NAMES = ['prologue', 'load_wait', 'rules', 'barrier', 'encode', 'drain']

for name in sys.argv[1:] or ['cfg4', 'cfg5_shard', 'cfg2']:
    kw = dict(bc.CONFIGS[name.split(':')[0]])
    if ':' in name:                      # e.g. cfg5_shard:max_episode_steps=1 -> every step ends the episode
        k, v = name.split(':')[1].split('=')
        kw[k] = int(v)
    N = kw.pop('num_envs')
    ns = kw['num_snakes']
    b = m.SnakeBatch(N, seed=0, **kw)
    g = torch.Generator(device='cuda').manual_seed(1)
    pool = torch.randint(0, 3, (64, N, ns), dtype=torch.uint8, device='cuda', generator=g)
    b.reset()
    for t in range(200):
        b.step(pool[t % 64], want_info=False)
    max_ctas = N + 8
    trace = torch.zeros((max_ctas, 8, 8), dtype=torch.int64, device='cuda')
    trace2 = torch.zeros((max_ctas, 8), dtype=torch.int64, device='cuda')
    m.lib.snk_prof_set_trace(C.c_void_p(trace.data_ptr()))
    m.lib.snk_prof_set_trace2(C.c_void_p(trace2.data_ptr()))
    torch.cuda.synchronize()
    b.step(pool[7], want_info=False)
    torch.cuda.synchronize()
    m.lib.snk_prof_set_trace(C.c_void_p(0))
    m.lib.snk_prof_set_trace2(C.c_void_p(0))
    t2 = trace2.cpu().numpy()
    t2 = t2[t2[:, 4] != 0]
    if len(t2):
        dd = np.diff(t2[:, :5].astype(np.float64), axis=1)
        reset = {k: round(float(dd[:, i].mean())) for i, k in enumerate(['walls', 'spawn', 'snakes', 'fruits'])}
        reset.update(resets=len(t2), attempts_mean=float(t2[:, 5].mean()), attempts_max=int(t2[:, 5].max()),
                     total_p50_p99=[int(np.percentile(t2[:, 4] - t2[:, 0], q)) for q in (50, 99)])
    else:
        reset = None
    tr = trace.cpu().numpy()
    used = tr[:, 0, 0] != 0
    tr = tr[used]
    nct = tr.shape[0]
    nwarps = int((tr[0, :, 0] != 0).sum())
    t0 = tr[:, :nwarps, 0].min()
    rec = {'config': name, 'ctas': nct, 'warps_per_cta': nwarps, 'reset_ns': reset}
    for role, sl in (('rule_warp', slice(0, 1)), ('other_warps', slice(1, nwarps))):
        if sl.start >= nwarps:
            continue
        x = tr[:, sl, :7].astype(np.float64)
        d = np.diff(x, axis=2)
        rec[role] = {k: round(float(d[:, :, i].mean())) for i, k in enumerate(NAMES)}
        rec[role]['life_ns'] = round(float((x[:, :, 6] - x[:, :, 0]).mean()))
    start = tr[:, 0, 0] - t0
    end = tr[:, :nwarps, 6].max(axis=1) - t0
    rec['launch_span_ns'] = int(end.max())
    rec['first_cta_done_ns'] = int(end.min())
    rec['last_cta_start_ns'] = int(start.max())
    # resident CTAs sampled every 5% of the span
    ts = np.linspace(0, end.max(), 21)
    rec['resident_ctas_over_time'] = [int(((start <= t) & (end > t)).sum()) for t in ts]
    rules = (tr[:, 0, 3] - tr[:, 0, 2]).astype(np.float64)
    life = (tr[:, :nwarps, 6].max(axis=1) - tr[:, 0, 0]).astype(np.float64)
    rec['rules_ns_percentiles_50_90_99_max'] = [int(np.percentile(rules, q)) for q in (50, 90, 99, 100)]
    rec['life_ns_percentiles_50_90_99_max'] = [int(np.percentile(life, q)) for q in (50, 90, 99, 100)]
    late = start > 0.7 * end.max()
    rec['rules_ns_p50_max_of_ctas_started_in_last_30pct'] = [int(np.percentile(rules[late], 50)), int(rules[late].max())] if late.any() else None
    sm = tr[:, 0, 7]
    rec['ctas_per_sm_min_max'] = [int(np.bincount(sm).min()), int(np.bincount(sm).max())]
    print(json.dumps(rec), flush=True)
    b.close()
