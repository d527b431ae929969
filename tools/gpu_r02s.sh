#!/bin/bash
# scaling on ONE 8-GPU box: N = 8, 4, 2, 1 back to back (run with gpurun --gpus 8)
mkdir -p gpurun_out
tag=${1:-r02s}
for n in 8 4 2; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
  echo "n=$n rc=$?"
done
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
echo "n=1 rc=$?"
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_n{n}.json").read().strip().splitlines()[-1])
        if n == 1: base = d["value"]
        print(n, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "eff", round(d["value"] / (n * base), 4) if base else None)
    except Exception as e:
        print(n, "unreadable:", e)
PY
tail -3 gpurun_out/${tag}_n8.err
