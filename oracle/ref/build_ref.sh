#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds oracle/_ref/libref_oracle.so from the reference sources
# WHERE THEY LIE under /root/reference (never copied into the repo's history).
#
# The reference's own build system (CMake + prebuilt TensorFlow, CMakeLists.txt:137-141) is
# not run: TensorFlow is un-vendored and absent.  Instead the TF-free translation units are
# compiled directly with the CMake default switches (CMakeLists.txt:15-21).  Three files
# must be replaced, and because `#include "rng.h"` resolves relative to the including file
# the only way to do that without editing /root/reference is a scratch copy of src/ that
# lives inside oracle/_ref/build/ for the duration of the build and is deleted afterwards:
#   src/rng.h                                   <- oracle/ref/overlay_rng.h     (reproducible RNG contract)
#   neural_network/alphazero_gpu_cluster.h      <- oracle/ref/overlay_gpu_cluster.h (TF-free NN facade)
#   alphazero_mcts.h / .cpp: std::unordered_map<LandIndex,SimulationValue> -> std::map
#       (PUCT tie-break = iteration order, alphazero_mcts.cpp:78,87; libstdc++'s hash order is
#        an artefact, the contract is ascending move index — SURVEY.md §7 hard part 3)
# alphazero_nn.cpp is compiled with -D_DEBUG against oracle/ref/tf_stub so the reference's own
# tensor encoder (setInStateTensor, alphazero_nn.cpp:31-67) is available.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${AZ_REFERENCE_DIR:-/root/reference}"
OUT="$REPO/oracle/_ref"
B="$OUT/build"
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF — keeping prebuilt oracle/_ref" >&2; exit 0; fi
rm -rf "$B"; mkdir -p "$B/obj"
cp -r "$REF/src" "$B/src"
mkdir -p "$B/libs"; cp -r "$REF/libs/xxhash" "$B/libs/xxhash"; cp -r "$REF/libs/cxxopts" "$B/libs/cxxopts"
AZ="$B/src/risk_game/player/alpha_zero"
cp "$HERE/overlay_rng.h" "$B/src/rng.h"
cp "$HERE/overlay_gpu_cluster.h" "$AZ/neural_network/alphazero_gpu_cluster.h"
sed -i 's/#include <unordered_map>/#include <unordered_map>\n#include <map>/' "$AZ/alphazero_mcts.h"
sed -i 's/std::unordered_map<LandIndex, SimulationValue>/std::map<LandIndex, SimulationValue>/g' "$AZ/alphazero_mcts.h" "$AZ/alphazero_mcts.cpp"
grep -q 'std::map<LandIndex, SimulationValue> moveValues' "$AZ/alphazero_mcts.h" || { echo "std::map overlay did not apply" >&2; exit 1; }

CXXFLAGS="-std=gnu++2a -O3 -w -fPIC -pthread -DINPUT_VECTOR_TYPE_2 -DSTATE_SIMPLE_CARDS -DFAST_ATTACK_MOBILIZATION -DFAST_REINFORCEMENT -I$B/src -I$B/libs -I$HERE/tf_stub -I$REPO/include"
R="$B/src/risk_game"
# land.cpp must precede land_set.cpp at link time (static-init order, land_set.cpp:3-9)
SRCS=("$R/land/land.cpp" "$R/state/state.cpp" "$R/land/land_set.cpp" "$AZ/alphazero_moves.cpp" "$R/player/game_helper.cpp"
      "$AZ/neural_network/alphazero_nn_data.cpp" "$AZ/alphazero_mcts.cpp" "$R/game/game.cpp" "$R/player/base/player.cpp"
      "$R/player/script/script_player.cpp" "$R/player/random/random_player.cpp")
pids=()
i=0; OBJS=()
for s in "${SRCS[@]}"; do
  o="$B/obj/$(printf '%02d' $i)_$(basename "$s" .cpp).o"; OBJS+=("$o"); i=$((i+1))
  g++ $CXXFLAGS -c "$s" -o "$o" & pids+=($!)
done
g++ $CXXFLAGS -D_DEBUG -c "$AZ/neural_network/alphazero_nn.cpp" -o "$B/obj/90_alphazero_nn.o" & pids+=($!)
g++ $CXXFLAGS -fno-access-control -c "$HERE/ref_shim.cpp" -o "$B/obj/91_ref_shim.o" & pids+=($!)
gcc -O3 -w -fPIC -c "$B/libs/xxhash/xxhash.c" -o "$B/obj/92_xxhash.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -pthread -o "$OUT/libref_oracle.so" "${OBJS[@]}" "$B/obj/90_alphazero_nn.o" "$B/obj/91_ref_shim.o" "$B/obj/92_xxhash.o"
echo "built $OUT/libref_oracle.so"

# ---- second variant: the same UNMODIFIED reference search / player / game loop, NN facade bound to the B200
# library through alphazero_risk_b200/host/az_nn_service.hpp (drop-in test, tests/test_dropin_gpu.py)
AZLIB="$REPO/alphazero_risk_b200/libaz_b200.so"
if [ -f "$AZLIB" ]; then
  cp "$HERE/overlay_gpu_cluster_b200.h" "$AZ/neural_network/alphazero_gpu_cluster.h"
  GFLAGS="$CXXFLAGS -I$REPO/alphazero_risk_b200/host"
  pids=()
  g++ $GFLAGS -c "$AZ/alphazero_mcts.cpp" -o "$B/obj/g_mcts.o" & pids+=($!)
  g++ $GFLAGS -c "$AZ/alphazero_player.cpp" -o "$B/obj/g_player.o" & pids+=($!)
  g++ $GFLAGS -fno-access-control -c "$HERE/ref_shim_gpu.cpp" -o "$B/obj/g_shim.o" & pids+=($!)
  for p in "${pids[@]}"; do wait "$p"; done
  GOBJS=()
  for o in "${OBJS[@]}"; do case "$o" in *alphazero_mcts.o) ;; *) GOBJS+=("$o");; esac; done
  g++ -shared -pthread -o "$OUT/libref_gpusvc.so" "${GOBJS[@]}" "$B/obj/g_mcts.o" "$B/obj/g_player.o" "$B/obj/g_shim.o" "$B/obj/92_xxhash.o" \
      -L"$REPO/alphazero_risk_b200" -laz_b200 -Wl,-rpath,'$ORIGIN/../../alphazero_risk_b200'
  echo "built $OUT/libref_gpusvc.so"
else
  echo "libaz_b200.so not built yet: skipping libref_gpusvc.so" >&2
fi
rm -rf "$B"
