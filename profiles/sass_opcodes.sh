#!/bin/bash
# Opcode histogram of the shipped library (cuobjdump -sass), per kernel family: evidence for TMA / bulk copies /
# mbarriers / packed FP32 in the SASS.   usage: bash profiles/sass_opcodes.sh > profiles/r2/sass_opcodes.txt
SO=feabas_b200/csrc/libfeabas_cuda.so
echo "# cuobjdump -sass $SO (sm_100a), $(date -u +%Y-%m-%dT%H:%MZ), $(stat -c %s $SO) bytes"
cuobjdump -sass $SO > /tmp/fb_sass.txt
echo "# kernels: $(grep -c 'Function :' /tmp/fb_sass.txt)"
echo
echo "## whole library: opcodes of interest (count)"
for op in UTMASTG UTMALDG UBLKCP UBLKPF SYNCS FADD2 FFMA2 FMUL2 UTCMMA LDTM HMMA LDS.128 LDS.64 STS.64 STS.128 LDGSTS BAR.SYNC SHFL DADD DFMA DMUL; do
  printf "%-10s %8d\n" $op $(grep -c "[ @]$op" /tmp/fb_sass.txt)
done
echo
echo "## per kernel family: instructions, and the 12 most frequent opcodes"
python3 - <<'PY'
import collections, re
fam = collections.OrderedDict()
cur = None
for line in open('/tmp/fb_sass.txt'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        key = next((k for k in ('fbk_wf', 'fbk_fast_columns', 'fbk_fast_rows_forward', 'fbk_fast_rows_inverse_tma', 'fbk_fast_rows_inverse', 'fbk_fused',
                                'fbk_rows_forward', 'fbk_columns', 'fbk_rows_inverse', 'fbk_finalize', 'fbk_channel_mean', 'fbk_gauss2d_reg', 'fbk_gauss2d_f32',
                                'fbk_gauss2d', 'fbk_crop_blocks', 'fbk_resize', 'fbk_minmax') if k in name), 'other')
        cur = fam.setdefault(key, [0, collections.Counter()])
        cur[0] += 1
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur is not None:
        cur[1][m.group(1).split('.')[0] if not m.group(1).startswith(('UTMA', 'UBLK', 'SYNCS', 'LDS', 'STS')) else m.group(1)] += 1
for k, (n, c) in fam.items():
    tot = sum(c.values())
    print(f'{k:28s} kernels {n:4d}  instructions {tot:9d}  ' + ', '.join(f'{o} {v}' for o, v in c.most_common(12)))
PY
