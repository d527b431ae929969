#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02x}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  echo "$name rc=$?"
}
run ctas8 NCCL_MAX_CTAS=8
run ctas4 NCCL_MAX_CTAS=4
run dbg NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=gpurun_out/${tag}_nccl_%h_%p.log
python - <<PY
import json
for f in ("ctas8", "ctas4", "dbg"):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3))
    except Exception as e:
        print(f, "unreadable:", e)
PY
ls gpurun_out/${tag}_nccl_* | head -2
f=$(ls gpurun_out/${tag}_nccl_* | head -1)
grep -E "NVLS|Channel|channels|Algo|algo|AllReduce" $f | head -20
# keep only one NCCL log, trimmed
head -c 200000 $f > gpurun_out/${tag}_nccl_rank.log; rm -f gpurun_out/${tag}_nccl_*_*.log
