# interleaved A/B of the training step's tensor-core routes: AZ_TRAIN_CONV=gemm (im2col + k_tc_gemm for every convolution),
# AZ_TRAIN_WGRAD=gemm (transposed-im2col GEMM for the weight gradients only), default (tower kernel + k_tc_wgrad)
for mode in "AZ_TRAIN_CONV=gemm" "AZ_TRAIN_WGRAD=gemm" "AZ_TRAIN_X=default" "AZ_TRAIN_WGRAD=gemm" "AZ_TRAIN_X=default"; do
env $mode python bench.py --steps 3 --warmup 3 --no-selfplay --play-games 0 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['train']['modes']['bf16_tcgen05']['ms_per_step'], d['train']['epoch'].get('samples_per_sec'), d['train']['modes']['bf16_tcgen05']['last_losses'])"; done
