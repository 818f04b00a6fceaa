python -m pytest tests/test_xcorr_gpu.py tests/test_abi.py -x -q -k "warp_fused or golden_small or plan_info" 2>&1 | tail -25
for f in "" "--force fused_smem"; do python bench.py --workload stitch_fine --steps 30 --no-cpu-baseline --no-e2e $f 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('stitch_fine', d['run_info'].get('path'), 'value', round(d['value']), 'ms/step', d['ms_per_step'])
    else: print(l.rstrip()[-300:])
"; done
