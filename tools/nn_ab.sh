#!/bin/bash
# A/B of the tower kernel's cluster size (AZ_TC_CLUSTER) on the GPU box: correctness first, then throughput
for c in 1 2 4; do
  echo "=== AZ_TC_CLUSTER=$c"
  AZ_TC_CLUSTER=$c timeout 300 python -m pytest tests/test_nn_gpu.py -m gpu -x -q 2>&1 | tail -3
  AZ_TC_CLUSTER=$c timeout 300 python tools/nn_bench.py 5 4096 16384 2>&1 | grep bf16
done
