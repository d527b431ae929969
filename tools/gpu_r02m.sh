#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02m}
timeout 1200 python -m pytest tests -m gpu -q -rfEs --no-header -p no:cacheprovider -x > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -8 gpurun_out/${tag}_gpu_tests.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
tail -2 gpurun_out/${tag}_n1.err
if [ "$2" == "n2" ]; then
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_n2.json 2> gpurun_out/${tag}_n2.err
echo "n2 rc=$?"; tail -4 gpurun_out/${tag}_n2.err
fi
python - <<PY
import json
for f in ("n1", "n2"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"], d["impl_detail"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
