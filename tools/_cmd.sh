mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -p no:cacheprovider -k configs 2>&1 | tail -30 > gpurun_out/run_o.log
cat gpurun_out/run_o.log | cut -c1-400
