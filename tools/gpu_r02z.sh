#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02z}
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -2
timeout 120 python tools/probe_sat.py upsample > gpurun_out/${tag}_sat.log 2>&1; cat gpurun_out/${tag}_sat.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
PY
