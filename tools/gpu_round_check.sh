#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, ncu --set full of the top kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.log
python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
python tools/gpu_kernel_loop.py ${1:-gelu2} > gpurun_out/plain_kernel.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${2:-gemm_tn} -c 3 -f -o gpurun_out/prof_${1:-gelu2} \
    python tools/gpu_kernel_loop.py ${1:-gelu2} > gpurun_out/ncu_kernel.log 2>&1
ls -la gpurun_out
