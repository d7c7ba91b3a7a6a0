#!/bin/bash
# development helper: bench.py under different weight-gradient split settings
for w in 1 2 3 4; do
  timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --debug-flags $((w*256)) 2>/dev/null | tail -1 > /tmp/b_$w.json
  python -c "
import json,sys
b=json.loads(open('/tmp/b_$w.json').read().strip().splitlines()[-1])
print('wgrad_ctas_per_sm', $w, round(b['value']), round(b['ms_per_step'],3), round(b['kernel_classes']['tc_mlp']['ms_per_step'],3))"
done
