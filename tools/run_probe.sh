#!/bin/bash
# Runs probe cases one process each (a trap in one case must not poison the next).
# usage: tools/run_probe.sh <script.py> <logname> case1 case2 ...
script=$1; shift
log=gpurun_out/$1; shift
mkdir -p gpurun_out
: > $log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $log 2>&1
for c in "$@"; do
  echo "=== $c" >> $log
  timeout 300 python $script $c >> $log 2>&1
  echo "    exit=$?" >> $log
done
tail -n 200 $log
