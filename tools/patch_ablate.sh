#!/bin/bash
# ablation sweep of conv_patch_tc_kernel on one layer (development tool): tools/patch_ablate.sh <case tag>
case=${1:-C32-64_257_N16}
for dbg in 0 1 8 4 16 2 12 20 28; do
  echo "DBG=$dbg $(SGK_PATCH=2 SGK_PATCH_DBG=$dbg python tools/layer_bench.py $case 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['case'], ' '.join('%s %.1f %s'%(k,d[k]['us'],d[k]['kernels'][:16]) for k in ('fwd','dgrad')))
")"
done
