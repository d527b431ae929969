#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err
tail -c 400 gpurun_out/r02_bench_2gpu_final.json; tail -3 gpurun_out/r02_bench_2gpu_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_medformer.py --batch 1 --side 128 > gpurun_out/r02_medformer_bench_2gpu.json 2> gpurun_out/r02_medformer_bench_2gpu.err
cat gpurun_out/r02_medformer_bench_2gpu.json; tail -3 gpurun_out/r02_medformer_bench_2gpu.err
python -m pytest tests/test_widen_gpu.py -m gpu -q -k "graphed or split" 2>&1 | tail -2
