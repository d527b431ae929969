#!/bin/bash
# Round-2 first GPU visit: full GPU suite with tracebacks, then bench A/B (stock glue / fused optimizer / CUDA graph),
# stock-PyTorch-on-B200 baseline and the CUDA-graph probe.
mkdir -p gpurun_out
tag=${1:-r02a}
timeout 900 python -m pytest tests -m gpu -q -rfEs --no-header -p no:cacheprovider --tb=long > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "pytest exit=$?" >> gpurun_out/${tag}_gpu_tests.log
timeout 200 python __graft_entry__.py smoke >> gpurun_out/${tag}_gpu_tests.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/${tag}_bench_stock_glue.json 2> gpurun_out/${tag}_bench_stock_glue.err
timeout 400 python bench.py --no-cpu-baseline --steps 10 --fused-optimizer --packed-labels > gpurun_out/${tag}_bench_fused_glue.json 2> gpurun_out/${tag}_bench_fused_glue.err
timeout 400 python bench.py --no-cpu-baseline --steps 10 --cuda-graph > gpurun_out/${tag}_bench_cuda_graph.json 2> gpurun_out/${tag}_bench_cuda_graph.err
timeout 400 python tools/probe_torch_gpu_baseline.py > gpurun_out/${tag}_torch_gpu_baseline.log 2>&1
timeout 300 python tools/probe_cuda_graph.py > gpurun_out/${tag}_cuda_graph.log 2>&1
grep -E "passed|failed|FAILED|ERROR|smoke" gpurun_out/${tag}_gpu_tests.log | tail -40
python - <<PY
import json
for f in ("stock_glue", "fused_glue", "cuda_graph"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2),
              "h2d", d["e2e"]["h2d_bytes_per_step"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/${tag}_bench_cuda_graph.err; tail -8 gpurun_out/${tag}_torch_gpu_baseline.log; tail -8 gpurun_out/${tag}_cuda_graph.log
