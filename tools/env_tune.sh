#!/bin/bash
# correctness of the restructured rollout kernel, then a sweep of its parking thresholds (run on the GPU box)
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-selfplay --no-cpu-baseline"
echo "v1:"; AZ_ENV_ROLLOUT=v1 $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'])"
for pf in ${PFS:-2 4 6 8 12 16}; do for pr in ${PRS:-1 2}; do
  echo -n "park_f=$pf park_r=$pr: "; AZ_ENV_PARK_F=$pf AZ_ENV_PARK_R=$pr $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'])"
done; done
