#!/bin/bash
# development: wgrad knob sweep over a few layers
for cfg in "" "SGK_WTMA_PATCH2=1" "SGK_WTMA_STAGES=3" "SGK_WTMA_STAGES=4" "SGK_WTMA_WAVES=2" "SGK_WTMA_PATCH2=1 SGK_WTMA_STAGES=3"; do
  for case in "$@"; do
    echo "[$cfg] $(env $cfg python tools/layer_bench.py $case 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:160]); continue
    print(d['case'], 'wgrad %.1f'%d['wgrad']['us'], d['err']['wgrad'])
")"
  done
done
