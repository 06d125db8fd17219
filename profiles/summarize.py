#!/usr/bin/env python
"""Turn ncu exports into the tracked summaries under profiles/.

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
  python profiles/summarize.py raw.csv launches.csv profiles/r01 10m

writes <prefix>_ncu_summary.md (per-kernel metrics) and updates profiles/traffic.json (DRAM bytes per launch, used by
bench.py's roofline.traffic)."""
import collections
import csv
import json
import os
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__grid_size', 'grid'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / instruction'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared bank conflicts'),
    ('sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'ADU pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main():
    raw, launches, prefix, workload = sys.argv[1:5]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    out = [f'# ncu --set full summary ({workload} workload, one step, --clock-control none)\n',
           'Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n']
    traffic = {}
    for r in rows[2:]:
        name = r[ki].split('(')[0].replace('void ', '').replace('lgr::', '')
        out.append(f'\n## {name}\n\n| metric | value |\n|---|---|')
        for k, label in KEYS:
            if k in hdr and r[hdr.index(k)]:
                out.append(f'| {label} (`{k}`) | {r[hdr.index(k)]} {units[hdr.index(k)]} |')
        stalls = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in hdr
                         if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')), reverse=True)[:4]
        out.append('| top stall reasons (warps per issue) | ' + ', '.join(f"{h.split('stalled_')[1].split('_per')[0]} {v:.2f}" for v, h in stalls) + ' |')
        rd = to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
        wr = to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
        key = name.split('<')[0].replace('_kernel', '')
        traffic[key] = traffic.get(key, 0) + rd + wr
    # launch list
    lrows = list(csv.reader(open(launches)))
    h = next(i for i, r in enumerate(lrows) if 'Kernel Name' in r)
    lh = lrows[h]
    agg = collections.OrderedDict()
    for r in lrows[h + 1:]:
        if len(r) <= lh.index('Metric Value'):
            continue
        v = float(r[lh.index('Metric Value')].replace(',', ''))
        u = r[lh.index('Metric Unit')]
        v = v / 1e6 if u == 'ns' else v / 1e3 if u in ('us', 'usecond') else v
        name = r[lh.index('Kernel Name')].split('(')[0].replace('void ', '')
        if 'lgr::' not in name:      # torch's own fill / copy kernels of the harness
            continue
        agg.setdefault(name.replace('lgr::', ''), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out.append('\n## launch list (gpu__time_duration.sum, every launch of the captured steps)\n\n| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f'| {k[:70]} | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | {sum(v) / tot * 100:.1f}% |')
    open(prefix + '_ncu_summary.md', 'w').write('\n'.join(out) + '\n')
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'traffic.json')
    allt = json.load(open(tp)) if os.path.exists(tp) else {}
    t = {'project_fwd': traffic.get('project_fwd', 0), 'project_bwd': traffic.get('project_bwd', 0),
         'blend_fwd': traffic.get('blend_fwd', 0), 'blend_bwd': traffic.get('blend_bwd', 0),
         'bin_sort': traffic.get('tile_scan', 0) + traffic.get('bin_scatter', 0) + traffic.get('tile_sort', 0)}
    allt[workload] = t
    json.dump(allt, open(tp, 'w'), indent=1)
    print('wrote', prefix + '_ncu_summary.md', tp)


if __name__ == '__main__':
    main()
