mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
timeout 200 python tools/probe_wgrad_timing.py 2>&1 | grep -v "max_ctas [1-9]"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --trace gpurun_out/trace_k.txt
} > gpurun_out/run_k.log 2>&1
cat gpurun_out/run_k.log | cut -c1-330
