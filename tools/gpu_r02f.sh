#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02f}
timeout 300 python tools/probe_grad_error.py 8 64 > gpurun_out/${tag}_grad_error_8_64.log 2>&1
cat gpurun_out/${tag}_grad_error_8_64.log
RSB_FPROP_NO_GROUPS=1 timeout 300 python tools/probe_grad_error.py 8 64 > gpurun_out/${tag}_grad_error_8_64_nogroups.log 2>&1
cat gpurun_out/${tag}_grad_error_8_64_nogroups.log | head -12
timeout 300 python tools/probe_grad_error.py 32 64 > gpurun_out/${tag}_grad_error_32_64.log 2>&1
cat gpurun_out/${tag}_grad_error_32_64.log
