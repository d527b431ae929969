#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02t}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 --no-allreduce > gpurun_out/${tag}_n8_noallreduce.json 2> gpurun_out/${tag}_n8_noallreduce.err
echo "rc=$?"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_n8.json 2> gpurun_out/${tag}_n8.err
echo "rc=$?"
python - <<PY
import json
for f in ("n8_noallreduce", "n8"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "per rank", d["impl_detail"]["ms_per_step_per_rank"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
