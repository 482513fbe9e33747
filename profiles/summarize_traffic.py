"""Summarise an `ncu --set full` capture: per-kernel time, DRAM traffic, L2 hit rate, registers, occupancy.

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_traffic.py /tmp/raw.csv <workload> [--json profiles/r02_ncu_traffic.json] > profiles/<name>.md

The JSON keeps dram__bytes_read.sum + dram__bytes_write.sum per launch of the three aggregation kernels (what bench.py
reports as roofline.traffic next to the algorithmic bytes); entries of other workloads in the file are kept."""
import csv
import json
import sys

NAMES = {'agg_lean_kernel<0': 'agg_fwd', 'agg_lean_kernel<1': 'agg_bwd_src', 'agg_lean_kernel<2': 'agg_bwd_rel'}
COLS = [('gpu__time_duration.sum', 'ms'), ('dram__bytes_read.sum', 'read GB'), ('dram__bytes_write.sum', 'write GB'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
        ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM busy %')]


def to_unit(value, unit, want):
    v = float(value.replace(',', ''))
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    if want == 'GB':
        return v * scale[unit] / 1e9
    if want == 'ms':
        return v * scale[unit]
    return v


def main():
    path, workload = sys.argv[1], sys.argv[2]
    out_json = sys.argv[sys.argv.index('--json') + 1] if '--json' in sys.argv else None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('| kernel | ' + ' | '.join(c[1] for c in COLS) + ' | DRAM TB/s |')
    print('|---|' + '---:|' * (len(COLS) + 1))
    traffic = {}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        short = name.split('(')[0].split('::')[-1]
        vals = []
        for m, label in COLS:
            if m not in idx:
                vals.append(float('nan'))
                continue
            want = 'GB' if label.endswith('GB') else ('ms' if label == 'ms' else '')
            vals.append(to_unit(r[idx[m]], units[idx[m]], want))
        ms, rd, wr = vals[0], vals[1], vals[2]
        print('| `{}` | '.format(short[:40]) + ' | '.join('{:.3f}'.format(v) if i < 3 else '{:.1f}'.format(v) for i, v in enumerate(vals))
              + ' | {:.2f} |'.format((rd + wr) / ms))
        for pat, key in NAMES.items():
            if pat in name and key not in traffic:
                traffic[key] = int((rd + wr) * 1e9)
    if out_json:
        try:
            with open(out_json) as f:
                data = json.load(f)
        except (OSError, ValueError):
            data = {}
        data.setdefault(workload, {}).update(traffic)
        with open(out_json, 'w') as f:
            json.dump(data, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
