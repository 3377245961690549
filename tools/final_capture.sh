set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo bench rc=$?
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
python tools/perf_configs.py --check > gpurun_out/perf_configs.log 2> gpurun_out/perf_configs.err; echo perf rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 27 --launch-count 9 -f -o gpurun_out/prof_r12 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo full rc=$?
ls -la gpurun_out/prof_r12.ncu-rep
