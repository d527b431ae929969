#!/bin/bash
# Scaling on ONE 8-GPU box: N = 8, 4, 2, 1 back to back, then the N = 8 diagnostic without the all-reduce (per-rank times).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale8.sh r02'
mkdir -p gpurun_out
tag=${1:-r02}
for n in 8 4 2; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  echo "n=$n rc=$?"
done
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 --no-allreduce > gpurun_out/${tag}_scale_n8_no_allreduce.json 2> gpurun_out/${tag}_scale_n8_no_allreduce.err
python - <<PY
import json
base = None
for n in (1, 2, 4, 8, "8_no_allreduce"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_scale_n{n}.json").read().strip().splitlines()[-1])
        if n == 1: base = d["value"]
        k = 8 if n == "8_no_allreduce" else n
        print(n, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "eff", round(d["value"] / (k * base), 4) if base else None, d["impl_detail"].get("ms_per_step_per_rank"))
    except Exception as e:
        print(n, "unreadable:", e)
PY
