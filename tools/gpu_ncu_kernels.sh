#!/bin/bash
# ncu --set full for a list of "<which>:<kernel regex>" pairs (tools/gpu_kernel_loop.py cases); 2 launches each.
mkdir -p gpurun_out
for pair in "$@"; do
  which=${pair%%:*}; rx=${pair##*:}
  python tools/gpu_kernel_loop.py $which > gpurun_out/plain_$which.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 2 -f -o gpurun_out/prof_$which \
      python tools/gpu_kernel_loop.py $which > gpurun_out/ncu_$which.log 2>&1
  tail -2 gpurun_out/ncu_$which.log
done
ls -la gpurun_out/*.ncu-rep
