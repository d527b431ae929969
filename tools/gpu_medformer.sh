#!/bin/bash
# MedFormer (row N1) on one B200: its GPU tests (+ the graph / split / UNet variants that share code with it), the bench line and
# the per-kernel trace.   gpurun --timeout 1800 -- bash tools/gpu_medformer.sh   then copy gpurun_out/r02_medformer_* into profiles/
mkdir -p gpurun_out
python -m pytest tests/test_medformer_gpu.py tests/test_widen_gpu.py::test_split_schedule_matches_eager_on_report_batches tests/test_unet_gpu.py -m gpu -q -s 2>&1 | grep -v "Saved to" > gpurun_out/r02_medformer_tests.log
grep "^\[medformer\|^\[split\|passed\|failed\|Error" gpurun_out/r02_medformer_tests.log | cut -c1-300 | tail -40
timeout 900 python tools/bench_medformer.py --batch 1 --side 128 --schedule both --trace gpurun_out/r02_medformer_trace.txt > gpurun_out/r02_medformer_bench_b1.json 2> gpurun_out/r02_medformer_bench_b1.err
tail -5 gpurun_out/r02_medformer_bench_b1.err; cat gpurun_out/r02_medformer_bench_b1.json; head -24 gpurun_out/r02_medformer_trace.txt
