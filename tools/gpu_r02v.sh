#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02v}
timeout 900 python -m pytest tests/test_report_losses_gpu.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --config cfg3 --no-cpu-baseline --no-torch-gpu-baseline --trace gpurun_out/${tag}_trace_cfg3.txt > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err
tail -2 gpurun_out/${tag}_bench_cfg3.err
python - <<PY
import json
for f in ("cfg3",):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
grep -E "ball_|dilate|rank" gpurun_out/${tag}_trace_cfg3.txt | head -4
