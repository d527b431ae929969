#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02h}
timeout 600 python bench.py --steps 10 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
tail -3 gpurun_out/${tag}_bench_cfg2.err
for c in cfg3 cfg4 cfg5; do
timeout 600 python bench.py --steps 5 --config $c --no-cpu-baseline --no-torch-gpu-baseline --trace gpurun_out/${tag}_trace_$c.txt > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
tail -3 gpurun_out/${tag}_bench_$c.err
done
python - <<PY
import json
for f in ("cfg2", "cfg3", "cfg4", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"], d["impl_detail"]["schedule"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
        print("cpu", d.get("cpu_baseline")); print("torch gpu", d.get("torch_gpu_baseline"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
grep -v conv3_ gpurun_out/${tag}_trace_cfg3.txt | head -40
