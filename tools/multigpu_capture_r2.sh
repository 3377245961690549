# Round-2 multi-GPU evidence (run through gpurun --gpus N): tools/multigpu_capture_r2.sh N
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/check_peer_allreduce.py > gpurun_out/peer${N}_r2.json 2> gpurun_out/peer${N}_r2.err; echo peer rc=$?
for ex in peer peer_unfused nccl; do
  timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --exchange $ex > gpurun_out/bench_${N}gpu_${ex}.json 2> gpurun_out/bench_${N}gpu_${ex}.err; echo bench $ex rc=$?
done
timeout 400 $TR --master-port 29523 bench.py --gpus $N --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${N}gpu.json 2> gpurun_out/bench_cfg4_${N}gpu.err; echo cfg4 rc=$?
timeout 400 $TR --master-port 29524 bench.py --gpus $N --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo ref rc=$?
python - <<PY
import json
for f in ("bench_${N}gpu_peer", "bench_${N}gpu_peer_unfused", "bench_${N}gpu_nccl", "bench_cfg4_${N}gpu", "bench_ref_${N}gpu"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d.get("allreduce_check"), d["config"]["parallelism"], d.get("stages_ms_per_step", {}).get("peer_allreduce"))
    except Exception as e:
        print(f, "FAILED", e)
PY
