#!/bin/bash
# same-box A/B of library variants (tools/build_variant.py), two rounds: tools/ab_lib.sh base v1 v2 ...
for r in 1 2; do for v in "$@"; do
if [ "$v" = base ]; then unset DM2_LIB; else export DM2_LIB=$PWD/dmesh2_renderer_b200/lib/variants/lib_$v.so; fi
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stages_ms_per_step']; print('$v', round(d['ms_per_step'],4), round(s['raster_forward'],4), round(s['raster_backward'],4))"
done; done
