#!/bin/bash
# diagnostic: is the tower kernel bound by chip-wide L2 delivery (fewer CTAs -> faster tiles) or by its own pipeline depth?
for g in 148 111 74 37; do
  echo "=== AZ_TC_GRID=$g cluster=1"
  AZ_TC_CLUSTER=1 AZ_TC_GRID=$g timeout 300 python tools/nn_bench.py 5 4096 2>&1 | grep bf16
done
echo "=== rebuild with 3 weight stages"
AZ_B200_NVCC_FLAGS="-DTC_STAGES=3" python -m alphazero_risk_b200.build --force > /dev/null 2>&1
AZ_TC_CLUSTER=1 timeout 300 python tools/nn_bench.py 5 4096 2>&1 | grep bf16
echo "=== rebuild with 2 weight stages"
AZ_B200_NVCC_FLAGS="-DTC_STAGES=2" python -m alphazero_risk_b200.build --force > /dev/null 2>&1
AZ_TC_CLUSTER=1 timeout 300 python tools/nn_bench.py 5 4096 2>&1 | grep bf16
