#!/usr/bin/env python
"""
Condenses an `ncu --set full` report into the tracked summaries under profiles/:

  python tools/ncu_summary.py <report.ncu-rep> --name r02_ncu_los_c3 [--kernel los_rl_kernel]
         [--los-entry c3:1:64]     # also (re)writes the LOS entry bench.py reads `roofline.traffic` from

Writes profiles/<name>.csv (metric, unit, value of the first matching launch, a fixed list of the
metrics DESIGN.md quotes) and, with --los-entry workload:n_gpus:n_local, updates
profiles/r02_ncu_los_summary.json.
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
]
BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--name', required=True)
    ap.add_argument('--kernel', default='')
    ap.add_argument('--los-entry', default='')
    args = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', args.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    row = next(r for r in rows[2:] if args.kernel in r[kn])
    out = os.path.join(ROOT, 'profiles', args.name + '.csv')
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', row[kn]])
        w.writerow(['report', os.path.basename(args.report)])
        w.writerow(['metric', 'unit', 'value'])
        for m in KEEP:
            if m in hdr:
                w.writerow([m, units[hdr.index(m)], row[hdr.index(m)]])
    print('wrote', out)
    if args.los_entry:
        wl, n_gpus, n_local = args.los_entry.split(':')

        def b(metric):
            i = hdr.index(metric)
            return float(row[i]) * BYTES[units[i]]
        entry = {'workload': wl, 'n_gpus': int(n_gpus), 'n_local': int(n_local), 'kernel': row[kn],
                 'dram_bytes': int(b('dram__bytes_read.sum') + b('dram__bytes_write.sum')),
                 'dram_bytes_read': int(b('dram__bytes_read.sum')), 'dram_bytes_write': int(b('dram__bytes_write.sum')),
                 'l1_global_load_sectors': float(row[hdr.index('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum')]),
                 'duration_ms_under_ncu': float(row[hdr.index('gpu__time_duration.sum')]) *
                 {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6}.get(
                     units[hdr.index('gpu__time_duration.sum')], 1.0),
                 'source': 'profiles/%s.csv (ncu --set full --clock-control none of the committed kernel)' % args.name}
        path = os.path.join(ROOT, 'profiles', 'r02_ncu_los_summary.json')
        try:
            prof = json.load(open(path))
        except Exception:
            prof = {'captures': []}
        prof['captures'] = [e for e in prof['captures']
                            if (e['workload'], e['n_gpus'], e['n_local']) != (wl, int(n_gpus), int(n_local))] + [entry]
        json.dump(prof, open(path, 'w'), indent=1)
        print('updated', path)


if __name__ == '__main__':
    main()
