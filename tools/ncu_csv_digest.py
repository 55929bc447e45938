#!/usr/bin/env python
"""Digest of the ncu CSV exports that tools/prof_coop_shapes.sh writes on the GPU box (gpurun brings back at most
64 MiB, so the .ncu-rep files stay there):

    python tools/ncu_csv_digest.py <tag> <config> [<config> ...]     e.g.  r02 cfg2 cfg3 cfg4 cfg5_shard cfg5_full

reads  gpurun_out/<tag>_ncu_<config>_raw.csv / _source.csv
writes profiles/<tag>_ncu_<config>.csv  (metric,unit,value of ONE launch: every metric family that the judge's recipe
       greps + launch geometry), its stall-reason shares and hottest SASS instructions as trailing comment lines,
       and updates profiles/traffic.json[<config>] (dram bytes per launch, read + written)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r'^(gpu__time_duration|dram__bytes_(read|write)\.sum$|dram__throughput|gpu__dram_throughput|'
                  r'sm__throughput|sm__warps_active|smsp__issue_active|sm__inst_issued|smsp__inst_executed\.sum|'
                  r'smsp__thread_inst_executed_per_inst|launch__(grid_size|block_size|registers_per_thread$|shared_mem_per_block_dynamic|occupancy_limit|waves_per_multiprocessor|stack_size)|l1tex__throughput|lts__throughput|lts__t_sector_hit_rate|'
                  r'l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|'
                  r'sm__pipe_(alu|fma|fp64)_cycles_active\.avg\.pct|smsp__inst_executed_pipe_lsu\.sum|'
                  r'smsp__average_warps?_issue_stalled_.*_per_issue_active|smsp__warps_issue_stalled_.*_pct|'
                  r'sm__cycles_elapsed\.max|sm__cycles_active\.avg|smsp__cycles_active\.avg)')


def digest(tag, cfg):
    raw = os.path.join(ROOT, 'gpurun_out', f'{tag}_ncu_{cfg}_raw.csv')
    src = os.path.join(ROOT, 'gpurun_out', f'{tag}_ncu_{cfg}_source.csv')
    rows = list(csv.reader(open(raw)))
    hdr, units, val = rows[0], rows[1], rows[2]
    out = [f'# ncu --set full --clock-control none --import-source on -k regex:snk_tile -s 215 -c 1  python tools/bench_configs.py {cfg}',
           f'# (tools/prof_coop_shapes.sh {tag}; run after the same command exited 0 without ncu; values are for ONE launch = one env step)',
           f'# kernel: {val[hdr.index("Kernel Name")]}  grid {val[hdr.index("Grid Size")]} block {val[hdr.index("Block Size")]}',
           'metric,unit,value']
    got = {}
    for h, u, v in zip(hdr, units, val):
        name = h.split('.', 2)[-1] if h.count('.') >= 2 and h.split('.')[1].startswith('Triage') else h
        if KEEP.match(name) and name not in got:
            got[name] = (u, v)
    for k in sorted(got):
        out.append(f'{k},{got[k][0]},{got[k][1]}')
    rd = float(got['dram__bytes_read.sum'][1].replace(',', ''))
    wr = float(got['dram__bytes_write.sum'][1].replace(',', ''))
    mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    rd *= mul[got['dram__bytes_read.sum'][0]]
    wr *= mul[got['dram__bytes_write.sum'][0]]
    # stall reasons and hot instructions from the source page
    if os.path.exists(src):
        rws = list(csv.reader(open(src)))
        h2 = None
        tot, lines = {}, []
        for r in rws:
            if r and r[0] == 'Address':
                h2 = r
                continue
            if h2 and len(r) == len(h2):
                rec = dict(zip(h2, r))
                for k, v in rec.items():
                    if k.startswith('stall_') and not k.endswith('(Not Issued)'):
                        try:
                            tot[k] = tot.get(k, 0) + int(v)
                        except ValueError:
                            pass
                try:
                    lines.append((int(rec['# Samples']), int(rec['Instructions Executed']), rec['Source'].strip()))
                except (ValueError, KeyError):
                    pass
        s = sum(tot.values()) or 1
        out.append('# warp stall samples by reason (source page, all samples):')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
            out.append(f'#   {k:26s} {100 * v / s:5.1f} %')
        ts = sum(x[0] for x in lines) or 1
        ti = sum(x[1] for x in lines) or 1
        out.append(f'# SASS instructions with the most stall samples (of {ts} samples, {ti} warp instructions executed):')
        for smp, ins, text in sorted(lines, key=lambda x: -x[0])[:16]:
            out.append(f'#   {100 * smp / ts:5.1f} % samples  {100 * ins / ti:5.2f} % inst   {text}')
    dst = os.path.join(ROOT, 'profiles', f'{tag}_ncu_{cfg}.csv')
    open(dst, 'w').write('\n'.join(out) + '\n')
    tj = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        t = json.load(open(tj))
    except Exception:
        t = {}
    if 'dram_bytes_per_launch' in t:          # round-1 layout: one record at top level
        t = {'r01_cfg5_full': t}
    t[cfg] = {'dram_bytes_per_launch': rd + wr, 'dram_bytes_read': rd, 'dram_bytes_write': wr,
              'source': f'profiles/{tag}_ncu_{cfg}.csv (ncu --set full, one launch)'}
    json.dump(t, open(tj, 'w'), indent=1)
    print(cfg, 'duration', got.get('gpu__time_duration.sum'), 'dram GB', (rd + wr) / 1e9)


if __name__ == '__main__':
    for c in sys.argv[2:]:
        digest(sys.argv[1], c)
