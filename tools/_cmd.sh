mkdir -p gpurun_out
for S in 64 128; do
RSB_TEST_SINGLE_S=$S timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -p no:cacheprovider -k "singleconv and fp32" 2>&1 | grep "single\] \|passed\|failed" | grep -v print
done > gpurun_out/run_z.log
cat gpurun_out/run_z.log | cut -c1-1000
