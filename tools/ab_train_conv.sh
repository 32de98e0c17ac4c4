for mode in gemm tower gemm; do AZ_TRAIN_CONV=$mode python bench.py --steps 3 --warmup 3 --no-selfplay --play-games 0 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['train']['modes']['bf16_tcgen05']['ms_per_step'], json.dumps(d['train']['epoch']))"; done
