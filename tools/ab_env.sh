#!/bin/bash
# same-box A/B of environment switches on the in-tree library: tools/ab_env.sh "DM2_KERNEL_RAYS=0" "DM2_KERNEL_RAYS=1" ...
for r in 1 2; do for e in "$@"; do
env $e python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stages_ms_per_step']; print('$e', round(d['ms_per_step'],4), round(s['raster_forward'],4), round(s['raster_backward'],4))"
done; done
