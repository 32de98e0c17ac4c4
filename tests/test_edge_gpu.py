"""GPU suite: sizes and inputs at the edges of the C ABI — one game, game counts that are not a multiple of any block size,
zero-length work, bad arguments, the 800-simulation node-pool capacity of BASELINE configs[4]."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    return a


@pytest.mark.parametrize("n", [1, 31, 33, 129, 1000])
def test_ragged_game_counts_rollout_equals_oracle(api, n):
    env = api.Env(n, first_game_id=5)
    env.reset(SEED)
    env.rollout(0)                                   # zero moves: a no-op that must not touch anything
    before = env.export_aos()
    steps = 150
    env.rollout(steps)
    dev = env.export_aos()
    assert env.counters()["steps"] == n * steps
    for g in sorted(set([0, n // 2, n - 1])):
        o = po.OracleGame()
        o.new_game(SEED, 5 + g, 0)
        assert (before[g] == o.data()).all()
        for ply in range(steps):
            if o.status() != -1:
                o.new_game(SEED, 5 + g, ply)
            assert o.move(o.random_action(SEED, 5 + g, ply), SEED, 5 + g, ply) == 0
        assert (dev[g] == o.data()).all()
    env.close()


def test_bad_arguments_are_rejected_with_a_message(api):
    L = api.lib()
    h = C.c_void_p()
    assert L.az_env_create(0, None, 0, 0, C.byref(h)) == -1 and b"n_games" in L.az_last_error()
    assert L.az_env_create(4, None, 99, 0, C.byref(h)) == -1
    assert L.az_nn_create(0, 0, C.byref(h)) == -1 and L.az_nn_create(27, 0, C.byref(h)) == -1
    env = api.Env(4)
    with pytest.raises(api.AzError):
        api.Mcts(env, net=None, evaluator=api.EVAL_NN)                    # network evaluator without a network
    with pytest.raises(api.AzError):
        api.Env(4, rules=api.default_rules(mcts_simulations=1, threads_per_mcts=2)) and api.Mcts(
            api.Env(4, rules=api.default_rules(mcts_simulations=1, threads_per_mcts=2)), evaluator=api.EVAL_PSEUDO)   # 1 - 1 % 2 = 0 sims
    mc = api.Mcts(env, evaluator=api.EVAL_PSEUDO)
    with pytest.raises(api.AzError):
        mc.samples()                                                      # recording was never enabled
    with pytest.raises(api.AzError):
        api.Arena(mc, opponent=7)
    # recorder of scripted / random turns: drained before it was enabled, enabled twice, bad player kinds, a scripted side without
    # its member array; a game with more samples than its staging area is dropped whole and counted
    with pytest.raises(api.AzError):
        env.turn_samples()
    with pytest.raises(api.AzError):
        env.record_turns(0)
    env.record_turns(capacity_samples=1000, max_samples_per_game=8)
    with pytest.raises(api.AzError):
        env.record_turns(1000)
    with pytest.raises(api.AzError):
        env.play_turn(api.OPPONENT_SCRIPT, api.OPPONENT_ALPHAZERO, np.full((4, 2), api.SCRIPT_INIT, np.uint32))
    with pytest.raises(api.AzError):
        env.play_turn(api.OPPONENT_SCRIPT, api.OPPONENT_RANDOM, None)
    env.reset(SEED)
    for _ in range(200):
        if (env.play_turn(api.OPPONENT_RANDOM, api.OPPONENT_RANDOM) != -1).all():
            break
    recs, dropped = env.turn_samples()
    assert len(recs) == 0 and dropped > 4 * 8                          # every game overflowed its 8-sample staging area
    mc.close(); env.close()


def test_network_batch_sizes_at_tile_boundaries(api):
    """the tcgen05 tower tiles 128 padded rows (2.29 boards): batches of 1, 2, 3, 7, 16, 37 boards cross every tile / pair boundary case"""
    net = api.Net(blocks=2, seed=11)
    rng = np.random.default_rng(3)
    big = rng.random((37, 546), dtype=np.float32)
    p_all, v_all = net.forward(big, api.FP32)
    for n in (1, 2, 3, 7, 16, 37):
        p32, v32 = net.forward(big[:n], api.FP32)
        assert (p32 == p_all[:n]).all() and (v32 == v_all[:n]).all()      # fp32 path is batch invariant
        p16, v16 = net.forward(big[:n], api.BF16)
        assert np.abs(p16 - p32).max() < 2e-3 and np.abs(v16 - v32).max() < 1e-2, n
    net.close()


def test_800_simulations_fit_the_node_pools(api):
    """BASELINE configs[4] runs 800 simulations per move: the pools (3 * (sims + 1) + 64 nodes) must not overflow, visit counts add up"""
    n, sims = 8, 800
    env = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1), first_game_id=0)
    env.reset(SEED)
    mc = api.Mcts(env, evaluator=api.EVAL_PSEUDO)
    for _ in range(3):
        res = mc.search(pick_mode=api.PICK_SELFPLAY, apply_move=True)
    assert mc.counters()["errors"] == 0
    peak, cap, per_game = mc.pool_stats()                     # measured occupancy next to the worst-case sizing
    assert cap == 3 * (sims + 1) + 64 and sims < peak <= cap and per_game >= 2 * cap * 640
    rs = mc.root_stats()
    assert (res["N"].sum(axis=1) == rs["sumN"]).all() and (rs["sumN"] >= sims).all()
    # one game replayed on the oracle
    rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    o, t = po.OracleGame(rules), po.OracleMcts(rules, "pseudo")
    o.new_game(SEED, 0, 0)
    for ply in range(3):
        a = t.search(o, SEED, 0, ply)
        mv = t.pick(a["pi"], True, SEED, 0, ply)
        if ply == 2:
            assert (res["N"][0] == a["N"]).all() and res["move"][0] == mv
        assert o.move(mv, SEED, 0, ply) == 0
    mc.close(); env.close()
