mkdir -p gpurun_out
timeout 300 python tools/probe_cpu_overhead.py > gpurun_out/run_v.log 2>&1
head -60 gpurun_out/run_v.log | cut -c1-180
