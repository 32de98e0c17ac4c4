# full GPU suite, one bench line, then (each only after its plain command exited 0) the ncu captures of the training GEMM
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo bench rc $?
python tools/train_once.py 512 bf16 > gpurun_out/train_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 40 -c 3 -o gpurun_out/r1_tc_gemm python tools/train_once.py 512 bf16 > gpurun_out/train_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
