#!/bin/bash
# development: env-knob sweep of conv_patch_tc_kernel over a few layers
for cfg in "" "SGK_PATCH_RW_KB=0" "SGK_PATCH_RW_KB=0 SGK_PATCH_SA=8" "SGK_PATCH_RW_KB=0 SGK_PATCH_SA=8 SGK_PATCH_MT=1" "SGK_PATCH_RW_KB=70 SGK_PATCH_SA=8"; do
  for case in "$@"; do
    echo "[$cfg] $(env SGK_PATCH=2 $cfg python tools/layer_bench.py $case 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['case'], ' '.join('%s %.1f %s'%(k,d[k]['us'],d[k]['kernels'][:12]) for k in ('fwd','dgrad')), max(d['err'].values()))
")"
  done
done
