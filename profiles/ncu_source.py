"""Per-kernel SASS hot spots of an .ncu-rep (read here, on the CPU box): samples by opcode and the
top instructions by stall samples.

    python profiles/ncu_source.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=25):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == 'Kernel Name':
            cur = {'name': row[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = row
        elif cur is not None:
            cur['rows'].append(row)
    seen = set()
    for b in blocks:
        if b['name'] in seen:
            continue
        seen.add(b['name'])
        h = {k: i for i, k in enumerate(b['hdr'])}
        stall_cols = [k for k in b['hdr'] if k.startswith('stall_') and 'Not Issued' not in k]
        tot = sum(int(r[h['# Samples']] or 0) for r in b['rows'])
        inst = sum(int(r[h['Instructions Executed']] or 0) for r in b['rows'])
        print('=' * 110)
        print(b['name'], ' samples', tot, ' warp-instructions', inst, ' SASS lines', len(b['rows']))
        by_op = collections.Counter(); n_op = collections.Counter()
        by_stall = collections.Counter()
        for r in b['rows']:
            op = r[h['Source']].split()[0] if r[h['Source']].split() else '?'
            if op.startswith('@'):
                op = r[h['Source']].split()[1]
            op = op.split('.')[0]
            by_op[op] += int(r[h['# Samples']] or 0)
            n_op[op] += int(r[h['Instructions Executed']] or 0)
            for k in stall_cols:
                by_stall[k] += int(r[h[k]] or 0)
        print('  samples by stall reason:', ', '.join(f'{k[6:]} {100*v/max(tot,1):.1f}%' for k, v in by_stall.most_common(9)))
        print('  opcode: executed share / sample share')
        for op, v in n_op.most_common(16):
            print(f'    {op:10s} {100*v/max(inst,1):6.1f}%  {100*by_op[op]/max(tot,1):6.1f}%')
        print(f'  top {top} instructions by samples:')
        rows = sorted(b['rows'], key=lambda r: -int(r[h['# Samples']] or 0))[:top]
        for r in rows:
            st = sorted(((int(r[h[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
            print(f"    {int(r[h['# Samples']]):6d}  {r[h['Source']].strip()[:70]:70s} {st}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
