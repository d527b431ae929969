#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02i}
timeout 900 python -m pytest tests/test_report_losses_gpu.py tests/test_kernels_gpu.py -m gpu -q -rfEs --no-header -p no:cacheprovider > gpurun_out/${tag}_report_tests.log 2>&1
tail -15 gpurun_out/${tag}_report_tests.log
for c in cfg3 cfg5; do
timeout 600 python bench.py --steps 5 --config $c --no-cpu-baseline --no-torch-gpu-baseline --trace gpurun_out/${tag}_trace_$c.txt > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
tail -3 gpurun_out/${tag}_bench_$c.err
done
python - <<PY
import json
for f in ("cfg3", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"], d["impl_detail"]["schedule"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
grep -v "conv3_\|norm_act\|instnorm\|upsample\|maxpool" gpurun_out/${tag}_trace_cfg3.txt | head -24
grep -v "conv3_\|norm_act\|instnorm\|upsample\|maxpool" gpurun_out/${tag}_trace_cfg5.txt | head -12
