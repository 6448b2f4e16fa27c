"""Summarise an ncu report (.ncu-rep, read here with `ncu -i`) and a launch list into profiles/<name>.md|json.
usage: python scripts/summarize_ncu.py <rep | raw-page.csv> <launches.csv|-> <out-prefix> [algorithmic-bytes-json]"""
import csv, io, json, subprocess, sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_peak'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu_dram_pct_of_peak'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_throughput_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved_occupancy_pct'),
    ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue_active_pct'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_active_pct'),
    ('launch__registers_per_thread', 'registers_per_thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem_per_block'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__waves_per_multiprocessor', 'waves_per_sm'),
    ('inst_executed', 'warp_instructions'),
    ('sass__inst_executed_global_loads', 'global_load_instructions'),
    ('sass__inst_executed_shared_loads', 'shared_load_instructions'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_scoreboard_per_issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_scoreboard_per_issue'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall_not_selected_per_issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier_per_issue'),
]


def to_bytes(val, unit):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit)
    return float(val) * mult if mult else float(val)


def to_ms(val, unit):
    return float(val) * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'second': 1e3, 's': 1e3}.get(unit, 1.0)


def main():
    rep, launches, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
    if rep.endswith('.csv'):      # `ncu -i <rep> --page raw --csv` already exported on the GPU box (reports > 64 MiB do not travel)
        raw = open(rep).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {'report': rep, 'kernels': []}
    for r in data:
        k = {'name': r[hdr.index('Kernel Name')]}
        for key, short in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v, u = r[i], units[i]
                try:
                    if short in ('dram_read', 'dram_write'):
                        k[short + '_bytes'] = to_bytes(v, u)
                    elif short == 'duration':
                        k['duration_ms'] = to_ms(v, u)
                    else:
                        k[short] = float(v)
                except ValueError:
                    k[short] = v
        if 'dram_read_bytes' in k:
            k['dram_bytes'] = k['dram_read_bytes'] + k['dram_write_bytes']
            k['dram_gbs'] = k['dram_bytes'] / (k['duration_ms'] * 1e-3) / 1e9
        out['kernels'].append(k)
    if launches != '-':
        ls = []
        for r in csv.DictReader(l for l in open(launches) if l.startswith('"')):
            if r.get('Metric Name') == 'gpu__time_duration.sum':
                ls.append({'kernel': r['Kernel Name'].split('(')[0].replace('void <unnamed>::', '').replace('<unnamed>::', ''),
                           'grid': r['Grid Size'], 'ns': float(r['Metric Value'])})
        out['launches'] = ls
    json.dump(out, open(prefix + '.json', 'w'), indent=1)
    with open(prefix + '.md', 'w') as f:
        f.write('# ncu summary: %s\n\n' % rep)
        f.write('`ncu --set full --clock-control none --import-source on` capture; numbers per launch.\n\n')
        for k in out['kernels']:
            f.write('## %s\n\n' % k['name'].split('(')[0])
            for kk, vv in k.items():
                if kk != 'name':
                    f.write('- %s: %s\n' % (kk, ('%.4g' % vv) if isinstance(vv, float) else vv))
            f.write('\n')
        if 'launches' in out:
            f.write('## launch list (`--metrics gpu__time_duration.sum`, serialised, cold cache)\n\n| # | kernel | grid | us |\n|---|---|---|---|\n')
            for i, l in enumerate(out['launches']):
                f.write('| %d | %s | %s | %.1f |\n' % (i, l['kernel'][:90], l['grid'], l['ns'] / 1e3))
            tot = {}
            for l in out['launches']:
                tot[l['kernel'][:60]] = tot.get(l['kernel'][:60], 0) + l['ns']
            s = sum(tot.values())
            f.write('\nShare of listed device time:\n\n')
            for kname, v in sorted(tot.items(), key=lambda kv: -kv[1]):
                f.write('- %s: %.1f %%\n' % (kname, 100 * v / s))
    print(open(prefix + '.md').read()[:3000])


main()
