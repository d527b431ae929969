#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02u}
timeout 900 python -m pytest tests/test_report_losses_gpu.py tests/test_unet_gpu.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
for c in cfg3 cfg5; do
timeout 600 python bench.py --steps 10 --config $c --no-cpu-baseline --no-torch-gpu-baseline --trace gpurun_out/${tag}_trace_$c.txt > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
tail -2 gpurun_out/${tag}_bench_$c.err
done
timeout 600 python bench.py --steps 10 --config cfg3 --schedule eager --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/${tag}_bench_cfg3_eager.json 2> gpurun_out/${tag}_bench_cfg3_eager.err
python - <<PY
import json
for f in ("cfg3", "cfg3_eager", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"], d["impl_detail"]["schedule"][:30])
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k in ("report", "dilate", "seg_loss", "head")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
grep -E "ball_|dilate|rank" gpurun_out/${tag}_trace_cfg3.txt | head
