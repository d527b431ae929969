mkdir -p gpurun_out
{
for g in 0 1; do
  if [ $g = 1 ]; then export RSB_FPROP_GENERIC=1; echo "--- generic"; else echo "--- compile-time NT"; fi
  python tools/probe_one.py 2 128 128 128 32 32 res
  python tools/probe_one.py 2 128 128 128 32 32 mask
  python tools/probe_one.py 2 128 128 128 96 64
  python tools/probe_one.py 2 128 128 128 64 96 mask
  python tools/probe_one.py 2 64 64 64 64 64 res
  python tools/probe_one.py 2 64 64 64 128 192 mask
  python tools/probe_one.py 2 32 32 32 128 128 res
  python tools/probe_one.py 2 16 16 16 256 256 res
  python tools/probe_one.py 2 8 8 8 320 320 res
done
unset RSB_FPROP_GENERIC
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
} > gpurun_out/run_p.log 2>&1
cat gpurun_out/run_p.log | cut -c1-300
