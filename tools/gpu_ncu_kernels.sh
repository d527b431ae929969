#!/bin/bash
# ncu --set full over ONE launch of every kernel family (tools/ncu_kernels.py); only rsb:: kernels are profiled, the raw
# page is exported as CSV on the box and the (large) .ncu-rep is not brought back.
mkdir -p gpurun_out
tag=${1:-r02}
for part in sat opt conv; do
  timeout 420 ncu --set full --clock-control none --kernel-name-base demangled -k regex:rsb:: -f -o /tmp/${tag}_$part python tools/ncu_kernels.py $part > gpurun_out/${tag}_ncu_$part.log 2>&1
  echo "$part rc=$?"; tail -2 gpurun_out/${tag}_ncu_$part.log
  ncu -i /tmp/${tag}_$part.ncu-rep --page raw --csv > gpurun_out/${tag}_$part.csv 2>/dev/null
  ls -la /tmp/${tag}_$part.ncu-rep gpurun_out/${tag}_$part.csv
done
du -sh gpurun_out
