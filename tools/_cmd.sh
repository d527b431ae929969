mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
for v in "" "res" "nostats" "mask"; do python tools/probe_one.py 2 128 128 128 32 32 $v; done
for v in "" "mask"; do python tools/probe_one.py 2 128 128 128 64 96 $v; python tools/probe_one.py 2 128 128 128 96 64 $v; done
python tools/probe_one.py 2 64 64 64 128 192 mask
python tools/probe_one.py 2 64 64 64 64 64 res
python tools/probe_one.py 2 32 32 32 128 128 res
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --trace gpurun_out/trace_h.txt
} > gpurun_out/probe_c.log 2>&1
cat gpurun_out/probe_c.log
