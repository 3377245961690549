#!/bin/bash
# A/B of library variants (tools/build_variant.py) on the GPU box: parity subset + bench per variant.
#   tools/ab_bench.sh base unroll3 ...      ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset DM2_LIB; else export DM2_LIB=$PWD/dmesh2_renderer_b200/lib/variants/lib_$v.so; fi
  t=$(python -m pytest tests -m gpu -x -q -k "small_scene or cfg1 or goldens or fused_path or cache" 2>&1 | tail -1)
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" "$t" <<'PY'
import json, sys
v, t = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json"))
    s = d["stages_ms_per_step"]
    print(f"{v:14s} {d['ms_per_step']:.4f} ms  fwd {s['raster_forward']:.4f}  bwd {s['raster_backward']:.4f}  classify {s['classify']:.4f}  | {t}")
except Exception as e:
    print(v, "FAILED", e, t)
PY
done
