#!/bin/bash
# env kernel block-size sweep on the GPU box (65536 games = 13.8 warps per SM: block size decides how evenly they spread)
for b in ${BLOCKS:-128 224 256 320 448}; do
  AZ_B200_NVCC_FLAGS="-DENV_BLOCK=$b" python -m alphazero_risk_b200.build --force > /dev/null 2>&1
  echo -n "ENV_BLOCK=$b: "; timeout -k 10 200 python bench.py --steps 10 --no-cpu-baseline --no-selfplay --play-games 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9)"
done
