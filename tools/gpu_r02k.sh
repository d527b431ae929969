#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02k}
timeout 300 python tools/probe_epilogue.py > gpurun_out/${tag}_epilogue.log 2>&1
cat gpurun_out/${tag}_epilogue.log
