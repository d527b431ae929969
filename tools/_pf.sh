#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_widen_gpu.py -m gpu -q -s -k "prefetch or graphed" 2>&1 | grep "^\[prefetch\|passed\|failed\|Error" | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/r02_pf_cfg2.json 2> gpurun_out/r02_pf_cfg2.err
timeout 600 python bench.py --steps 10 --config cfg3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/r02_pf_cfg3.json 2> gpurun_out/r02_pf_cfg3.err
python - <<PY
import json
for f in ("cfg2", "cfg3"):
    d = json.loads(open(f"gpurun_out/r02_pf_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 2))
PY
tail -2 gpurun_out/r02_pf_cfg2.err
