#!/bin/bash
# bench line + one-step ncu launch list + ncu --set full of the dominant kernels
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err
python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
bash tools/gpu_ncu_kernels.sh wgrad:gemm_wgrad lnbwd:gemm_tn
