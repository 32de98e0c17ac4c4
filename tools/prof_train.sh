python -m pytest tests/test_train_gpu.py tests/test_nn_gpu.py tests/test_loop_gpu.py -x -q 2>&1 | tail -5
python tools/train_once.py 512 bf16 > gpurun_out/train_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_train_bf16_launches.csv python tools/train_once.py 512 bf16 > gpurun_out/train_ncu.log 2>&1; tail -3 gpurun_out/train_plain.log
for mode in tower gemm; do AZ_TRAIN_CONV=$mode python bench.py --steps 3 --warmup 3 --no-selfplay --play-games 0 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', json.dumps(d['train']['modes']['bf16_tcgen05']), d['train']['epoch'].get('samples_per_sec'))"; done
