#!/bin/bash
# 2-GPU check of the graph-captured step with the NCCL all-reduce inside (run with gpurun --gpus 2)
mkdir -p gpurun_out
tag=${1:-r02j}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "rc=$?"
tail -5 gpurun_out/${tag}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --schedule eager > gpurun_out/${tag}_bench_n2_eager.json 2> gpurun_out/${tag}_bench_n2_eager.err
echo "rc=$?"
tail -3 gpurun_out/${tag}_bench_n2_eager.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
for f in ("n1", "n2", "n2_eager"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"], d["impl_detail"]["schedule"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
