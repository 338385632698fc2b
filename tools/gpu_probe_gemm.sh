#!/bin/bash
# Runs every GEMM probe case in its own process with a timeout; logs to gpurun_out/probe_gemm.log
mkdir -p gpurun_out
L=gpurun_out/probe_gemm.log
: > $L
for c in "store" "gelu_mul" "ln" "wgrad 8192 1024" "wgrad 1024 8192" "wgrad 8192 128" "wgrad 128 1024"; do
  echo "=== case $c" >> $L
  timeout 240 python tools/gpu_probe_gemm.py $c >> $L 2>&1
  echo "=== exit $?" >> $L
done
tail -n 120 $L
