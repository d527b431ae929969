mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "upsample or pool or stem" 2>&1 | tail -3
timeout 300 python tools/probe_sat.py
} > gpurun_out/run_r.log 2>&1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --clock-control none -k regex:"maxpool|upsample|stem_|pack_conv3" -c 24 --csv --page raw --log-file gpurun_out/sat_ncu.csv python tools/probe_sat.py > /dev/null 2>&1
cat gpurun_out/run_r.log | cut -c1-200; ls -la gpurun_out/sat_ncu.csv
