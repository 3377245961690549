N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_peer.json 2> gpurun_out/bench_${N}gpu_peer.err; echo bench rc=$?
timeout 400 $TR --master-port 29523 bench.py --gpus $N --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${N}gpu.json 2> gpurun_out/bench_cfg4_${N}gpu.err; echo cfg4 rc=$?
python - <<PY
import json
for f in ("bench_${N}gpu_peer", "bench_cfg4_${N}gpu"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d.get("allreduce_check"), d["config"]["parallelism"])
PY
