# ncu evidence for the default bench command (short self-play settings so that the launch list stays small); every ncu run comes
# after the same command exited 0 without ncu.  Usage on the GPU box:  bash tools/round_profile.sh r2   (tag = file prefix)
# Afterwards HERE:  python tools/ncu_summary.py gpurun_out/<tag>_env_rollout.ncu-rep profiles/<tag>_env_rollout_summary.csv  (same for
# _selfplay), then tools/traffic_from_ncu.py --sha gpurun_out/<tag>_source_sha.json ... -> profiles/traffic.json
set -x
TAG=${1:-r2}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sp-moves 1 --sp-steps 2 --play-games 0 --train-batch 0 --cfg5-games 0 --no-blocks20 --env6-games 0 --cfg4-games 0"
python tools/source_sha.py > gpurun_out/${TAG}_source_sha.json
$CMD > gpurun_out/${TAG}_prof_plain.json 2> gpurun_out/${TAG}_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_prof_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_env_rollout -s 3 -c 1 -o gpurun_out/${TAG}_env_rollout $CMD > gpurun_out/${TAG}_prof_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_nn_conv_tc3|k_mcts_sim|k_nn_heads_tail|k_nn_conv_tc<' -s 400 -c 14 -o gpurun_out/${TAG}_selfplay $CMD > gpurun_out/${TAG}_prof_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
