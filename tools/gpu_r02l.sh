#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02l}
B="python bench.py --steps 10 --no-cpu-baseline --no-torch-gpu-baseline"
timeout 300 $B > gpurun_out/${tag}_default.json 2> gpurun_out/${tag}_default.err
RSB_FPROP_STREAM=1 timeout 300 $B > gpurun_out/${tag}_stream.json 2> gpurun_out/${tag}_stream.err
timeout 300 $B --no-side-stream > gpurun_out/${tag}_noside.json 2> gpurun_out/${tag}_noside.err
RSB_FPROP_STREAM=1 timeout 300 $B --no-side-stream > gpurun_out/${tag}_stream_noside.json 2> gpurun_out/${tag}_stream_noside.err
python - <<PY
import json
for f in ("default", "stream", "noside", "stream_noside"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "igemm", round(d["kernels"]["conv3_igemm"]["ms_per_step"], 3), "clocks", d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
