#!/bin/bash
# The record run of a round: full GPU suite, smoke(), bench lines for every BASELINE.json GPU config and for the parity mode.
#   gpurun --timeout 2400 -- 'bash tools/gpu_records.sh r02'      then copy gpurun_out/<tag>_* into profiles/
mkdir -p gpurun_out
tag=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -q -rfEsxX --no-header -p no:cacheprovider > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -4 gpurun_out/${tag}_gpu_tests.log
timeout 200 python __graft_entry__.py smoke >> gpurun_out/${tag}_gpu_tests.log 2>&1
grep smoke gpurun_out/${tag}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --trace gpurun_out/${tag}_trace_cfg2.txt > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
timeout 600 python bench.py --steps 10 --precision fp32 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_bench_cfg2_fp32.json 2> gpurun_out/${tag}_bench_cfg2_fp32.err
for c in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --steps 10 --config $c --no-cpu-baseline --no-torch-gpu-baseline --trace gpurun_out/${tag}_trace_$c.txt > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
done
python - <<PY
import json
for f in ("cfg2", "cfg2_fp32", "cfg3", "cfg4", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "flop-roofline", round(d["conv3d_flop_roofline_frac"], 3),
              "dominant kernel frac", round(d["roofline"]["frac"], 3), "|", d["impl_detail"]["schedule"][:24], "| cpu", (d.get("cpu_baseline") or {}).get("value"),
              "torch gpu", (d.get("torch_gpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
# MedFormer (row N1): not a bench.py workload (the north-star metric is the UNet step) — its own timing script
timeout 900 python tools/bench_medformer.py --batch 2 --side 128 --schedule graph > gpurun_out/${tag}_medformer_bench_b2.json 2> gpurun_out/${tag}_medformer_bench_b2.err
tail -c 900 gpurun_out/${tag}_medformer_bench_b2.json
# the reference arm exactly as the driver launches it
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
tail -c 600 gpurun_out/${tag}_bench_reference_arm.json
