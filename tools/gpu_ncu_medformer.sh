#!/bin/bash
# ncu --set full over one launch of every MedFormer kernel (tools/ncu_kernels.py mf); summary: tools/summarize_ncu.py roofline gpurun_out/r02_mf.csv
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:rsb:: -f -o /tmp/r02_mf python tools/ncu_kernels.py mf > gpurun_out/r02_ncu_mf.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_ncu_mf.log
ncu -i /tmp/r02_mf.ncu-rep --page raw --csv > gpurun_out/r02_mf.csv 2>/dev/null
ls -la /tmp/r02_mf.ncu-rep gpurun_out/r02_mf.csv
