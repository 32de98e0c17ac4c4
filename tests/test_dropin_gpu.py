"""GPU suite: the DROP-IN.  oracle/_ref/libref_gpusvc.so is the UNMODIFIED reference search / player / game loop
(AlphaZeroMCTS, AlphaZeroPlayer, GameGroup, ScriptPlayer) compiled against the host adapter
alphazero_risk_b200/host/az_nn_service.hpp, i.e. the reference's own code served by the B200 network.
  * reference MCTS (one host thread per game, leaf requests batched across threads by the adapter) must
    produce the same visit counts and moves as the B200 lockstep MCTS with the same network;
  * BASELINE config 1's shape (`-m play`: AlphaZero vs ScriptPlayer through GameGroup::playGames) must run."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_gpusvc.so")
SEED = 0x5EED0001


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_gpusvc.so not built (needs /root/reference at build time)")
    L = C.CDLL(LIB)
    L.refgpu_new.restype = C.c_void_p
    L.refgpu_new.argtypes = [C.c_int, C.c_int, C.c_uint64]
    L.refgpu_free.argtypes = [C.c_void_p]
    L.refgpu_last_error.restype = C.c_char_p
    vp = C.c_void_p
    L.refgpu_selfplay_threads.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.refgpu_play_vs_script.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_double)]
    return L


def test_reference_mcts_on_b200_network_equals_b200_mcts(ref):
    from alphazero_risk_b200 import api
    from oracle import pyoracle as po
    n, first, sims, cap, blocks = 8, 300, 12, 48, 2
    h = ref.refgpu_new(blocks, api.FP32, 1234)
    assert h, ref.refgpu_last_error()
    visits = np.zeros((n, cap, 43), np.uint32)
    moves = np.zeros((n, cap), np.uint8)
    nmoves = np.zeros(n, np.int32)
    final = np.zeros((n, 160), np.uint8)
    rc = ref.refgpu_selfplay_threads(h, n, first, SEED, sims, cap, visits.ctypes.data, moves.ctypes.data, nmoves.ctypes.data,
                                     final.ctypes.data, None)
    assert rc == 0, ref.refgpu_last_error()
    ref.refgpu_free(h)
    assert (nmoves == cap).all()
    # the same games on the B200 lockstep MCTS with the same random-init network (az_nn_init_random(1234))
    env = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1), first_game_id=first)
    env.reset(SEED)
    net = api.Net(blocks=blocks, seed=1234)
    mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.FP32)
    for ply in range(cap):
        res = mc.search(pick_mode=api.PICK_SELFPLAY, apply_move=True)
        assert (res["N"] == visits[:, ply]).all(), ply
        assert (res["move"] == moves[:, ply]).all(), ply
    mask = po.data_byte_mask()
    assert (env.export_aos()[:, mask] == final[:, mask]).all()
    assert mc.counters()["errors"] == 0
    mc.close(); net.close(); env.close()


def test_config1_play_az_vs_script_through_reference_game_loop(ref):
    from alphazero_risk_b200 import api
    h = ref.refgpu_new(5, api.BF16, 1234)
    assert h, ref.refgpu_last_error()
    out, secs = np.zeros(5, np.int32), C.c_double(0)
    rc = ref.refgpu_play_vs_script(h, 8, 16, 2, 4, out.ctypes.data, C.byref(secs))
    assert rc == 0, ref.refgpu_last_error()
    ref.refgpu_free(h)
    count, draw, az, script, _ = [int(v) for v in out]
    assert count == 8 and az + script + draw == 8
    print("config-1 shape: 8 games AZ(16 sims, T=2, random-init 5-block net on B200) vs ScriptPlayer: az %d script %d draw %d in %.1f s"
          % (az, script, draw, secs.value))


MAIN = os.path.join(ROOT, "oracle", "_ref", "AlphaZero_risk")


def test_reference_main_runs_config0_verbatim(tmp_path):
    """BASELINE configs[0] verbatim: the reference's OWN program — main(), settings.h command line, executePlay, GameGroup::playGames
    with one host thread per game, AlphaZeroPlayer / AlphaZeroMCTS with 2 search threads per tree, ScriptPlayer, its own rng.h —
    compiled unmodified (oracle/ref/build_ref_main.sh) against the B200 network through the NN facade adapter, run as
    `AlphaZero_risk -m play --mcts=16 --cg=<n>` and read from the result lines it prints (src/alphazero_risk.cpp:46)."""
    import re
    import subprocess
    if not os.path.exists(MAIN):
        pytest.skip("oracle/_ref/AlphaZero_risk not built (needs /root/reference at build time)")
    (tmp_path / "log").mkdir()                     # LOG.init / Settings::init write log/*.txt relative to the working directory
    games = 64
    out = subprocess.run([MAIN, "-m", "play", "--mcts=16", "--cg=%d" % games], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    m = re.search(r"Games: (\d+)\s+Draws:(\d+)\s+Player 1:(\d+)\s+Player 2:(\d+)", out.stdout)
    assert m, out.stdout[-2000:]
    count, draws, p1, p2 = [int(v) for v in m.groups()]
    assert count == games and draws + p1 + p2 == games
    assert "Starting program with GPUs: 1" in out.stdout and "MCTS simulations 16" in out.stdout
    assert (tmp_path / "checkpoints" / "latest-checkpoint.bin.index").exists()      # AlphaZeroNN::loadCheckpoint: missing -> init + save
    assert (tmp_path / "log" / "settings.txt").exists()
    print("AlphaZero_risk -m play --mcts=16 --cg=%d on B200: AlphaZero %d, ScriptPlayer %d, draws %d" % (games, p1, p2, draws))
