"""Summarise an .ncu-rep (read here, on the CPU box) into the handful of numbers DESIGN.md cites.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for r in rows[2:]:
        print('=' * 100)
        print(r[idx['Kernel Name']])
        for w in WANT:
            if w in idx:
                print(f'  {w:75s} {r[idx[w]]:>16s} {units[idx[w]]}')
        vals = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:7]
        print('  top stall reasons (warps stalled per issue):')
        for v, h in vals:
            print(f'      {v:8.3f}  ' + h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))


SLOTS = (('rows_forward', 'rows_forward'), ('columns', 'columns'), ('rows_inverse', 'rows_inverse'),
         ('finalize', 'finalize'), ('fused', 'fused'))


def traffic(path, workload, out_json):
    """profiles/traffic.json[workload][kernel slot] = mean (dram read + write) bytes per launch."""
    import json
    import os
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        slot = next((s for key, s in SLOTS if key in name), None)
        if slot is None:
            continue
        b = sum(float(r[idx[m]]) * scale[units[idx[m]]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        acc.setdefault(slot, []).append(b)
    data = {}
    if os.path.exists(out_json):
        data = json.load(open(out_json))
    data[workload] = {k: sum(v) / len(v) for k, v in acc.items()}
    json.dump(data, open(out_json, 'w'), indent=1, sort_keys=True)
    print(json.dumps(data[workload]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[2] == '--traffic':
        traffic(sys.argv[1], sys.argv[3], sys.argv[4])
    else:
        main(sys.argv[1])
