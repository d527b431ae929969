#!/bin/bash
# FIRST GPU visit of round 2: run the staged tests (tests/test_widen_gpu.py: XPASS = green on hardware, XFAIL = traceback in the
# log under -rxX), then A/B the staged fused optimizer and packed-label upload in the bench, then the stock-PyTorch baseline.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r02a'
mkdir -p gpurun_out
tag=${1:-r02a}
timeout 600 python -m pytest tests -m gpu -q -rxX --no-header -p no:cacheprovider > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "pytest exit=$?" >> gpurun_out/${tag}_gpu_tests.log
timeout 300 python -m pytest tests/test_widen_gpu.py -m gpu -q -s --runxfail --no-header -p no:cacheprovider > gpurun_out/${tag}_staged_strict.log 2>&1
echo "pytest (staged, strict) exit=$?" >> gpurun_out/${tag}_staged_strict.log
timeout 200 python __graft_entry__.py smoke >> gpurun_out/${tag}_gpu_tests.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/${tag}_bench_stock_glue.json 2> gpurun_out/${tag}_bench_stock_glue.err
timeout 400 python bench.py --no-cpu-baseline --steps 10 --fused-optimizer --packed-labels > gpurun_out/${tag}_bench_fused_glue.json 2> gpurun_out/${tag}_bench_fused_glue.err
timeout 400 python bench.py --no-cpu-baseline --steps 10 --cuda-graph > gpurun_out/${tag}_bench_cuda_graph.json 2> gpurun_out/${tag}_bench_cuda_graph.err
timeout 400 python tools/probe_torch_gpu_baseline.py > gpurun_out/${tag}_torch_gpu_baseline.log 2>&1
timeout 300 python tools/probe_cuda_graph.py > gpurun_out/${tag}_cuda_graph.log 2>&1
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clip_adamw_ema -c 2 -o gpurun_out/${tag}_clip_adamw_ema_full \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fused-optimizer > gpurun_out/${tag}_ncu_optimizer.log 2>&1
fi
grep -E "passed|failed|FAILED|XPASS|XFAIL|xpassed|xfailed|smoke" gpurun_out/${tag}_gpu_tests.log | tail -60
tail -5 gpurun_out/${tag}_staged_strict.log
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
tail -3 gpurun_out/${tag}_bench_cuda_graph.err; tail -8 gpurun_out/${tag}_torch_gpu_baseline.log; tail -8 gpurun_out/${tag}_cuda_graph.log
