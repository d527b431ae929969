mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/probe_sat.py stem
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --trace gpurun_out/trace_t.txt
} > gpurun_out/run_t.log 2>&1
cat gpurun_out/run_t.log | cut -c1-250
