# round-end ncu evidence for the default bench command (short self-play settings so that the launch list stays small);
# every ncu run comes after the same command exited 0 without ncu
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sp-moves 1 --sp-steps 2 --play-games 0 --train-batch 0"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_final.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_env_rollout -s 3 -c 1 -o gpurun_out/r1_env_rollout_final $CMD > gpurun_out/prof_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_nn_conv_tc3|k_mcts_sim|k_nn_heads_tc|k_nn_conv_tc<' -s 400 -c 14 -o gpurun_out/r1_selfplay_final $CMD > gpurun_out/prof_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
