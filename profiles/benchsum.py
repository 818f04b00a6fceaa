import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('%.0f matches/s  frac %.3f  e2e %s'%(d["value"], d["roofline"]["pipeline"]["frac"], d["e2e"] and round(d["e2e"]["value"])), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()}, d['clocks'])
