# Round-2 evidence capture on one B200 (run through gpurun from the repo root; ~4 GPU-minutes):
#   gpurun --timeout 1700 -- 'bash tools/final_capture_r2.sh'
# Every number quoted in DESIGN.md / profiles/r02_* comes from the files this writes under gpurun_out/.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo bench rc=$?
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_1gpu.json 2> gpurun_out/bench_cfg4_1gpu.err; echo cfg4 rc=$?
python tools/perf_configs.py cfg1 cfg3 cfg4 cfg5 --check --stages > gpurun_out/perf_configs.log 2> gpurun_out/perf_configs.err; echo perf rc=$?
python tools/sort_ab.py > gpurun_out/sort_ab.log 2>&1; echo sort rc=$?
python tools/probe_ray_recipe.py > gpurun_out/ray_recipe.log 2>&1; echo rays rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --clock-window-s 0 > gpurun_out/ncu_l.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 27 --launch-count 9 -f -o gpurun_out/prof_r2_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --clock-window-s 0 > gpurun_out/ncu.log 2>&1; echo full rc=$?
python tools/perf_configs.py cfg3 --no-ref --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_layers --launch-count 3 -f -o gpurun_out/prof_layers_final python tools/perf_configs.py cfg3 --no-ref --iters 1 > gpurun_out/ncu_layers.log 2>&1; echo layers rc=$?
ls -la gpurun_out/prof_r2_final.ncu-rep gpurun_out/prof_layers_final.ncu-rep
