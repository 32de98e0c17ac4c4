#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds oracle/_ref/AlphaZero_risk: the reference's OWN program (src/alphazero_risk.cpp main, settings.h CLI,
# executePlay -> GameGroup::playGames with its thread-per-game loop, AlphaZeroPlayer, AlphaZeroMCTS, ScriptPlayer, its own rng.h)
# from the sources WHERE THEY LIE under /root/reference, linked against libaz_b200.so through the NN facade adapter
# (alphazero_risk_b200/host/az_nn_service.hpp).  This is BASELINE configs[0] verbatim: `AlphaZero_risk -m play --mcts=16 --cg=1000`.
# Not the reference's build system (CMake + prebuilt TensorFlow, unbuildable here).  What is replaced in the scratch copy:
#   neural_network/alphazero_gpu_cluster.h  <- oracle/ref/overlay_gpu_cluster_b200_main.h  (the binding INTEGRATION.md describes)
#   neural_network/alphazero_gpu_cluster.cpp   not compiled (its classes are the adapter's templates)
#   board/board_gui.h                       <- oracle/ref/board_gui_stub.h   (Windows-only GUI)
#   <tensorflow/...> includes               -> oracle/ref/tf_stub; alphazero_nn.cpp with -D_DEBUG (its built-in fake backend; only
#                                              executeAnalysis would touch it)
# Everything else — rng.h, the unordered_map in the search, game loop, players, trainer — is compiled unmodified.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${AZ_REFERENCE_DIR:-/root/reference}"
OUT="$REPO/oracle/_ref"
B="$OUT/build_main"
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF — keeping prebuilt oracle/_ref" >&2; exit 0; fi
AZLIB="$REPO/alphazero_risk_b200/libaz_b200.so"
if [ ! -f "$AZLIB" ]; then echo "libaz_b200.so not built yet" >&2; exit 1; fi
rm -rf "$B"; mkdir -p "$B/obj" "$OUT"
cp -r "$REF/src" "$B/src"
mkdir -p "$B/libs"; cp -r "$REF/libs/xxhash" "$B/libs/xxhash"; cp -r "$REF/libs/cxxopts" "$B/libs/cxxopts"
AZ="$B/src/risk_game/player/alpha_zero"
cp "$HERE/overlay_gpu_cluster_b200_main.h" "$AZ/neural_network/alphazero_gpu_cluster.h"
cp "$HERE/overlay_gpu_cluster_b200.h" "$AZ/neural_network/overlay_gpu_cluster_b200.h"
cp "$HERE/board_gui_stub.h" "$B/src/risk_game/board/board_gui.h"
CXXFLAGS="-std=gnu++2a -O3 -w -pthread -DINPUT_VECTOR_TYPE_2 -DSTATE_SIMPLE_CARDS -DFAST_ATTACK_MOBILIZATION -DFAST_REINFORCEMENT -I$B/src -I$B/libs -I$HERE/tf_stub -I$REPO/include -I$REPO/alphazero_risk_b200/host"
R="$B/src/risk_game"
SRCS=("$R/land/land.cpp" "$R/state/state.cpp" "$R/land/land_set.cpp" "$AZ/alphazero_moves.cpp" "$R/player/game_helper.cpp"
      "$AZ/neural_network/alphazero_nn_data.cpp" "$AZ/alphazero_mcts.cpp" "$AZ/alphazero_player.cpp" "$AZ/alphazero_trainer.cpp"
      "$R/game/game.cpp" "$R/player/base/player.cpp" "$R/player/script/script_player.cpp" "$R/player/random/random_player.cpp"
      "$B/src/alphazero_risk.cpp")
pids=(); i=0; OBJS=()
for s in "${SRCS[@]}"; do
  o="$B/obj/$(printf '%02d' $i)_$(basename "$s" .cpp).o"; OBJS+=("$o"); i=$((i+1))
  g++ $CXXFLAGS -c "$s" -o "$o" & pids+=($!)
done
g++ $CXXFLAGS -D_DEBUG -c "$AZ/neural_network/alphazero_nn.cpp" -o "$B/obj/90_alphazero_nn.o" & pids+=($!)
gcc -O3 -w -c "$B/libs/xxhash/xxhash.c" -o "$B/obj/92_xxhash.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -pthread -o "$OUT/AlphaZero_risk" "${OBJS[@]}" "$B/obj/90_alphazero_nn.o" "$B/obj/92_xxhash.o" \
    -L"$REPO/alphazero_risk_b200" -laz_b200 -Wl,-rpath,'$ORIGIN/../../alphazero_risk_b200'
rm -rf "$B"
echo "built $OUT/AlphaZero_risk"
