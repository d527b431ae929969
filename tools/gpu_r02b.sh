#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02b}
timeout 600 python -m pytest tests/test_widen_gpu.py -m gpu -q -rfEs --no-header -p no:cacheprovider -x -k "graphed or resume or capturable or sliding" > gpurun_out/${tag}_graph_tests.log 2>&1
tail -30 gpurun_out/${tag}_graph_tests.log
timeout 600 python tools/probe_parity_depth.py 32 64 128 > gpurun_out/${tag}_parity_depth.log 2>&1
cat gpurun_out/${tag}_parity_depth.log | tail -40
timeout 400 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/${tag}_bench_stock_glue.json 2> gpurun_out/${tag}_bench_stock_glue.err
timeout 400 python bench.py --no-cpu-baseline --steps 10 --cuda-graph > gpurun_out/${tag}_bench_cuda_graph.json 2> gpurun_out/${tag}_bench_cuda_graph.err
python - <<PY
import json
for f in ("stock_glue", "cuda_graph"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2),
              "h2d", d["e2e"]["h2d_bytes_per_step"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/${tag}_bench_cuda_graph.err
