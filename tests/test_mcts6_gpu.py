"""GPU suite of the six-player SEARCH (az_mcts6_*, SIXPLAYER.md).  No reference parity exists for this game; the checker is
oracle/risk6_oracle.c's r6_mcts_search.  Visit counts, Q, P, pi, table sizes and chosen moves are compared as BIT PATTERNS with the
exactly representable pseudo evaluator on both sides; the network path is checked for its encoding (bit-exact against the oracle's
r6_encode) and, with the oracle search calling the SAME CUDA forward, for identical searches."""
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


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def run_lockstep6(api, n, first, sims, play_mode, max_moves, rules_kw=None, net=None, precision=None, start_rollout=0):
    rules_kw = rules_kw or {}
    rules_o = po.default_rules(mcts_simulations=sims, threads_per_mcts=1, **rules_kw)
    env = api.Env6(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1, **rules_kw), first_game_id=first)
    env.reset(SEED)
    if start_rollout:
        env.rollout(start_rollout)
    eval_fn = None
    if net is not None:
        L = po.oracle_lib()
        L.r6_encode.argtypes = [C.POINTER(po.R6State), np.ctypeslib.ndpointer(np.float32, flags="C")]

        @C.CFUNCTYPE(None, C.POINTER(po.R6State), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)
        def eval_fn(sp, policy, value, user):
            x = np.zeros(po.INPUT_FLOATS, np.float32)
            L.r6_encode(sp, x)
            p, v = net.forward(x.reshape(1, -1), precision)
            C.memmove(policy, p.ctypes.data, 43 * 4)
            value[0] = float(v[0])

    mc = api.Mcts6(env, net=net, evaluator=api.EVAL_NN if net is not None else api.EVAL_PSEUDO,
                   precision=api.FP32 if precision is None else precision)
    img = env.export()
    games = [po.Oracle6Game(rules_o) for _ in range(n)]
    trees = [po.Oracle6Mcts(rules_o, eval_fn) for _ in range(n)]
    for g, o in enumerate(games):
        o.set_image(img[g])
    ply = np.full(n, start_rollout, int)
    moves_done = 0
    for step in range(max_moves):
        res = mc.search(pick_mode=api.PICK_ARGMAX if play_mode else api.PICK_SELFPLAY, apply_move=True)
        rs = mc.root_stats()
        for g, o in enumerate(games):
            if o.status() != -1:
                assert res["move"][g] == 43 and res["status"][g] == o.status()
                continue
            a = trees[g].search(o, SEED, first + g, int(ply[g]))
            assert (res["N"][g] == a["N"]).all(), (step, g, res["N"][g], a["N"])
            assert (bits(rs["Q"][g]) == bits(a["Q"])).all(), (step, g)
            assert (bits(rs["P"][g]) == bits(a["P"])).all(), (step, g)
            assert (bits(res["pi"][g]) == bits(a["pi"])).all(), (step, g)
            assert rs["sumN"][g] == a["sumN"] and rs["table"][g] == trees[g].table_size(), (step, g)
            sample = (not play_mode) and o.s.round <= rules_o.temperature_threshold
            mv = trees[g].pick(a["pi"], sample, SEED, first + g, int(ply[g]))
            assert res["move"][g] == mv, (step, g)
            assert o.move(mv, SEED, first + g, int(ply[g])) == 0
            ply[g] += 1
            moves_done += 1
            assert res["status"][g] == o.status()
        if step % 16 == 0 or step == max_moves - 1:
            dev = env.export()
            for g, o in enumerate(games):
                assert (dev[g] == o.image()).all(), (step, g)
    cnt = mc.counters()
    assert cnt["errors"] == 0 and cnt["sims"] == moves_done * sims and cnt["steps"] == moves_done
    mc.close(); env.close()
    return moves_done


def test_selfplay_pseudo_16_sims_from_the_deal(api):
    assert run_lockstep6(api, n=12, first=300, sims=16, play_mode=False, max_moves=110) > 1200


@pytest.mark.parametrize("sims,play_mode,start,kw", [(24, False, 600, {}), (64, True, 900, {}), (200, False, 1500, {}),
                                                      (20, False, 800, dict(cpuct=2.5, dir_noise_epsi=0.5, dir_noise_value=0.05, temperature_threshold=20)),
                                                      (16, True, 1200, dict(limit_attack=1, limit_reinforcement=0, min_unit_move=1, allow_yield=0, max_game_rounds=40))])
def test_midgame_search_vs_oracle(api, sims, play_mode, start, kw):
    """searches from mid-game positions (reached by a rollout): attacks with dice inside the descents, eliminated seats skipped,
    game ends inside the tree, the -v / 5 rule on paths that change seats; non-default hyper-parameters and rules"""
    assert run_lockstep6(api, n=6, first=40, sims=sims, play_mode=play_mode, max_moves=40, rules_kw=kw, start_rollout=start) > 100


def test_encoding_matches_oracle(api):
    n = 200
    env = api.Env6(n, first_game_id=9)
    env.reset(SEED)
    env.rollout(1700)
    img, x = env.export(), env.encode()
    o, t = po.Oracle6Game(), po.Oracle6Mcts()
    for g in range(n):
        o.set_image(img[g])
        assert (bits(x[g].reshape(-1)) == bits(t.encode(o))).all(), g
    env.close()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_search_with_network_matches_oracle_given_same_outputs(api, prec):
    net = api.Net(blocks=2, seed=4321)
    assert run_lockstep6(api, n=4, first=70, sims=10, play_mode=False, max_moves=30, net=net, precision=api.FP32 if prec == "fp32" else api.BF16,
                         start_rollout=700) > 60
    net.close()


def test_selfplay6_device_loop_and_full_size_properties(api):
    """az_selfplay6_run at configs[3]'s game count with the network: counters add up, the run is reproducible, a half-size shard
    with the second half's global ids reproduces the second half"""
    n, sims, moves = 16384, 8, 3
    net = api.Net(blocks=1, seed=5)
    rules = api.default_rules(mcts_simulations=sims, threads_per_mcts=1)

    def run(count, first):
        env = api.Env6(count, rules=rules, first_game_id=first)
        env.reset(SEED)
        mc = api.Mcts6(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
        mc.selfplay(moves)
        cnt, img = mc.counters(), env.export().copy()
        mc.close(); env.close()
        return cnt, img

    cnt, img = run(n, 0)
    assert cnt["errors"] == 0 and cnt["sims"] == n * sims * moves and cnt["steps"] == n * moves and cnt["evals"] >= n * moves
    cnt2, img2 = run(n, 0)
    assert cnt2 == cnt and (img2 == img).all()
    _, half = run(n // 2, n // 2)
    assert (half == img[n // 2:]).all()
    net.close()
