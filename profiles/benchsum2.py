"""One line per bench JSON: python profiles/benchsum2.py gpurun_out/r2jobs/bench_*_n*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
    except Exception as exc:
        print(f, 'unreadable', exc)
        continue
    e = d.get('e2e') or {}
    r = d.get('run_info') or {}
    print(f"{f.split('/')[-1]:38s} N={d['n_gpus']} value={d['value']:11.0f} ms/step={d['ms_per_step']:9.2f} e2e={e.get('value', 0):11.0f} "
          f"jobs/s={r.get('jobs_per_s') or 0:8.1f} xcorr_share={(d['roofline'] or {}).get('xcorr_share_of_step') or 0:.3f} "
          f"pipe_frac={((d['roofline'] or {}).get('pipeline') or {}).get('frac') or 0:.3f} clk={(d.get('clocks') or {}).get('sm_mhz')}")
