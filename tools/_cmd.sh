mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --trace gpurun_out/trace_l.txt
} > gpurun_out/run_l.log 2>&1
cat gpurun_out/run_l.log | cut -c1-330
grep -E "upsample|pack|stem|head|maxpool" gpurun_out/trace_l.txt
