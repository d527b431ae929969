#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02g}
RSB_SIDE_STREAM=1 timeout 400 python bench.py --no-cpu-baseline --steps 10 --cuda-graph > gpurun_out/${tag}_bench_graph_side.json 2> gpurun_out/${tag}_bench_graph_side.err
timeout 400 python bench.py --no-cpu-baseline --steps 10 --cuda-graph --trace gpurun_out/${tag}_trace.txt > gpurun_out/${tag}_bench_graph.json 2> gpurun_out/${tag}_bench_graph.err
python - <<PY
import json
for f in ("graph_side", "graph"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/${tag}_bench_graph_side.err
cat gpurun_out/${tag}_trace.txt
