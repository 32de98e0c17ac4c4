#!/bin/bash
# A/B of the tower kernel variants on the GPU box: correctness first, then throughput
for m in single pair; do
  echo "=== AZ_TC_MODE=$m"
  AZ_TC_MODE=$m timeout -k 10 180 python -m pytest tests/test_nn_gpu.py -m gpu -x -q 2>&1 | tail -3
  AZ_TC_MODE=$m timeout -k 10 180 python tools/nn_bench.py 5 4096 16384 2>&1 | grep bf16
done
