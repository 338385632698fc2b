#!/bin/bash
# GPU tests + bench line + one-step ncu launch list + ncu --set full of the dominant kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 2400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
bash tools/gpu_ncu_kernels.sh "$@"
