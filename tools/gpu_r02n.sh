#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02n}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline"
for i in 1 2; do
timeout 300 $B > gpurun_out/${tag}_on$i.json 2> gpurun_out/${tag}_on$i.err
RSB_FPROP_STREAM=0 timeout 300 $B > gpurun_out/${tag}_off$i.json 2> gpurun_out/${tag}_off$i.err
done
timeout 900 python -m pytest tests/test_widen_gpu.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
python - <<PY
import json
for f in ("on1", "off1", "on2", "off2"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "igemm", round(d["kernels"]["conv3_igemm"]["ms_per_step"], 3), d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
