"""GPU suite: the CUDA MCTS through the C ABI against the CPU oracle (itself pinned bit-for-bit to the
compiled reference's AlphaZeroMCTS) and the reference's golden search traces.  Visit counts, Q, P, pi,
table sizes and chosen moves are compared as BIT PATTERNS given identical evaluator outputs."""
import ctypes as C
import os

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


def run_lockstep(api, n, first, sims, T, play_mode, max_moves, evaluator="pseudo", net=None, oracle_eval=None, K=1, rules_kw=None,
                 precision=None):
    rules_kw = rules_kw or {}
    rules_o = po.default_rules(mcts_simulations=sims, threads_per_mcts=T, **rules_kw)
    rules_d = api.default_rules(mcts_simulations=sims, threads_per_mcts=T, concurrent_descents=K, **rules_kw)
    env = api.Env(n, rules=rules_d, first_game_id=first)
    env.reset(SEED)
    ev = {"pseudo": api.EVAL_PSEUDO, "uniform": api.EVAL_UNIFORM, "nn": api.EVAL_NN}[evaluator]
    mc = api.Mcts(env, net=net, evaluator=ev, precision=api.FP32 if precision is None else precision)
    assert mc.simulations() == sims - sims % T
    games = [po.OracleGame(rules_o) for _ in range(n)]
    trees = []
    for g, o in enumerate(games):
        o.new_game(SEED, first + g, 0)
        t = po.OracleMcts(rules_o, "pseudo" if evaluator == "nn" else evaluator)
        if oracle_eval is not None:
            t.L.ro_mcts_free(t.h)
            t.h = t.L.ro_mcts_new(C.cast(oracle_eval, C.c_void_p), None)
        trees.append(t)
    ply = np.zeros(n, int)
    last_cur = [None] * n
    moves_done = 0
    for step in range(max_moves):
        extra = np.zeros(n, np.uint8)
        if play_mode:
            for g, o in enumerate(games):
                if o.status() == -1 and o.s.cur != last_cur[g]:
                    extra[g] = 1
        res = mc.search(pick_mode=api.PICK_ARGMAX if play_mode else api.PICK_SELFPLAY, apply_move=True,
                        extra_trim=extra if play_mode else None)
        rs = mc.root_stats()
        for g, o in enumerate(games):
            if o.status() != -1:
                assert res["move"][g] == 43 and res["status"][g] == o.status()
                continue
            if play_mode and extra[g]:
                trees[g].trim(); last_cur[g] = o.s.cur
            a = trees[g].search(o, SEED, first + g, int(ply[g]), lockstep=K)
            assert (res["N"][g] == a["N"]).all(), (step, g, res["N"][g], a["N"])
            assert (bits(rs["Q"][g]) == bits(a["Q"])).all(), (step, g)
            assert (bits(rs["P"][g]) == bits(a["P"])).all(), (step, g)
            assert (bits(res["pi"][g]) == bits(a["pi"])).all(), (step, g)
            assert rs["sumN"][g] == a["sumN"] and bits(rs["value"][g:g + 1])[0] == bits(np.float32([a["value"]]))[0]
            assert rs["table"][g] == trees[g].table_size(), (step, g, rs["table"][g], trees[g].table_size())
            sample = (not play_mode) and o.s.round <= rules_o.temperature_threshold
            mv = trees[g].pick(a["pi"], sample, SEED, first + g, int(ply[g]))
            assert res["move"][g] == mv, (step, g)
            assert o.move(mv, SEED, first + g, int(ply[g])) == 0
            ply[g] += 1
            moves_done += 1
            assert res["status"][g] == o.status()
        if step % 16 == 0 or step == max_moves - 1:
            dev = env.export_aos()
            for g, o in enumerate(games):
                assert (dev[g] == o.data()).all(), (step, g)
        if all(o.status() != -1 for o in games):
            break
    cnt = mc.counters()
    assert cnt["errors"] == 0 and cnt["illegal"] == 0
    assert cnt["sims"] == moves_done * (sims - sims % T)
    if K > 1:      # the active_N rule must have fired: moves passed over and duplicate requests
        assert sum(t.vl_counts()[0] for t in trees) > 0
    mc.close(); env.close()
    return moves_done


def test_selfplay_pseudo_net_16_sims(api):
    assert run_lockstep(api, n=24, first=500, sims=16, T=1, play_mode=False, max_moves=130) > 2500


def test_selfplay_full_games_64_sims(api):
    assert run_lockstep(api, n=3, first=9000, sims=64, T=1, play_mode=False, max_moves=400) > 500


def test_play_mode_turn_start_trim_and_thread_rounding(api):
    # MCTS_SIMULATIONS = 33 with THREADS_PER_MCTS = 2 -> 32 simulations (alphazero_mcts.cpp:265)
    assert run_lockstep(api, n=6, first=77, sims=33, T=2, play_mode=True, max_moves=150) > 500


@pytest.mark.parametrize("sims,T,K,play_mode,evaluator", [(16, 2, 2, True, "pseudo"), (48, 1, 4, False, "pseudo"), (32, 1, 8, False, "uniform"),
                                                          (64, 1, 16, False, "pseudo")])
def test_virtual_loss_concurrent_descents(api, sims, T, K, play_mode, evaluator):
    """az_rules.concurrent_descents = K: K descents per tree per leaf batch with the reference's active_N rule
    (getNextBestMoveAndSetVisited, alphazero_mcts.cpp:91-111) == the oracle's lockstep schedule, which
    tests/test_oracle_vs_ref.py pins to the reference's own search threads taking turns"""
    assert run_lockstep(api, n=8, first=4100, sims=sims, T=T, play_mode=play_mode, max_moves=140, evaluator=evaluator, K=K) > 700


# the sets tests/test_oracle_vs_ref.py::test_mcts_lockstep pins oracle-vs-reference (CPUCT, DIR_NOISE_VALUE, DIR_NOISE_EPSI,
# TEMPERATURE_THRESHOLD of settings.h:40-62 -> alphazero_mcts.cpp:67-119, alphazero_trainer.cpp:98-106; game rules inside the
# search -> alphazero_moves.cpp:18,36,59), here CUDA-vs-oracle
MCTS_RULE_SETS = [dict(cpuct=2.5, dir_noise_value=0.05, dir_noise_epsi=0.5, temperature_threshold=6),
                  dict(cpuct=0.4, dir_noise_epsi=0.0, temperature_threshold=0, limit_attack=1, limit_reinforcement=0),
                  dict(dir_noise_value=1.0, dir_noise_epsi=1.0, allow_yield=0, max_game_rounds=40, min_unit_move=1)]


@pytest.mark.parametrize("sims,K,play_mode,kw", [(24, 1, False, 0), (20, 1, False, 1), (16, 1, True, 2), (32, 1, False, 2),
                                                 (32, 4, False, 0), (16, 2, True, 1)])
def test_non_default_hyper_parameters_and_rules(api, sims, K, play_mode, kw):
    assert run_lockstep(api, n=6, first=2600 + 10 * kw, sims=sims, T=1, play_mode=play_mode, max_moves=130, K=K,
                        rules_kw=MCTS_RULE_SETS[kw]) > 500


def test_concurrent_descents_must_divide_the_simulation_count(api):
    env = api.Env(2, rules=api.default_rules(mcts_simulations=16, threads_per_mcts=1, concurrent_descents=3))
    with pytest.raises(api.AzError):
        api.Mcts(env, evaluator=api.EVAL_PSEUDO)
    env.close()


def test_uniform_evaluator(api):
    assert run_lockstep(api, n=4, first=31, sims=24, T=1, play_mode=False, max_moves=60, evaluator="uniform") > 200


@pytest.mark.parametrize("name", ["selfplay16", "selfplay64", "play32t2", "selfplay200", "selfplay800"])
def test_reference_golden_search_traces(api, golden_dir, name):
    """the compiled reference's own search traces (tests/golden/mcts_trace*.npz), one full game each; 200 and 800 simulations
    per move are the depths of BASELINE configs[3] / [4] (pool migration, long paths)"""
    t = np.load(os.path.join(golden_dir, "mcts_trace_deep.npz" if name in ("selfplay200", "selfplay800") else "mcts_trace.npz"))
    sims, T, play_mode, g = [int(v) for v in t[name + "_cfg"]]
    seed = int(t["seed"])
    env = api.Env(1, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=T), first_game_id=g)
    env.reset(seed)
    mc = api.Mcts(env, evaluator=api.EVAL_PSEUDO)
    n = len(t[name + "_move"])
    for ply in range(n):
        assert (env.export_aos()[0] == t[name + "_root"][ply]).all(), ply
        extra = np.array([1 if t[name + "_trimmed"][ply] else 0], np.uint8)
        res = mc.search(pick_mode=api.PICK_ARGMAX if play_mode else api.PICK_SELFPLAY, apply_move=True, extra_trim=extra)
        rs = mc.root_stats()
        assert (res["N"][0] == t[name + "_N"][ply]).all(), ply
        assert (bits(rs["Q"][0]) == t[name + "_Q"][ply]).all() and (bits(rs["P"][0]) == t[name + "_P"][ply]).all()
        assert (bits(res["pi"][0]) == t[name + "_pi"][ply]).all()
        assert rs["sumN"][0] == t[name + "_sumN"][ply] and bits(rs["value"])[0] == t[name + "_value"][ply]
        assert rs["table"][0] == t[name + "_table"][ply]
        assert res["move"][0] == t[name + "_move"][ply]
    assert (env.export_aos()[0] == t[name + "_final"]).all()
    assert env.status()[0] == int(t[name + "_status"])
    assert mc.counters()["errors"] == 0
    mc.close(); env.close()


@pytest.mark.parametrize("K,prec", [(1, "fp32"), (3, "fp32"), (2, "bf16"), (4, "bf16")])
def test_search_with_network_matches_oracle_given_same_outputs(api, K, prec):
    """evaluator = the network (fp32 path, or the bf16 tcgen05 tower).  The oracle search calls the SAME network (batch of one, the
    kernels are batch-invariant) for its leaf evaluations, so both sides see identical outputs.  K > 1 with the tensor-core path
    also covers the root round, which evaluates one descent slot per game out of the K-slot leaf arrays."""
    net = api.Net(blocks=2, seed=1234)
    L = po.oracle_lib()
    precision = api.FP32 if prec == "fp32" else api.BF16

    @C.CFUNCTYPE(None, C.POINTER(po.RoState), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)
    def evaluator(sp, policy, value, user):
        x = np.zeros(po.INPUT_FLOATS, np.float32)
        L.ro_encode(sp, x)
        p, v = net.forward(x.reshape(1, -1), precision)
        C.memmove(policy, p.ctypes.data, 43 * 4)
        value[0] = float(v[0])

    assert run_lockstep(api, n=5, first=4242, sims=12, T=1, play_mode=False, max_moves=70, evaluator="nn", net=net,
                        oracle_eval=evaluator, K=K, precision=precision) > 300
    net.close()


def test_selfplay_run_device_loop(api):
    """az_selfplay_run: many moves with no host synchronisation, finished games re-dealt, trees cleared"""
    n, moves = 64, 500
    env = api.Env(n, rules=api.default_rules(mcts_simulations=8, threads_per_mcts=1), first_game_id=0)
    env.reset(SEED)
    mc = api.Mcts(env, evaluator=api.EVAL_PSEUDO)
    mc.selfplay(moves)
    cnt = mc.counters()
    assert cnt["errors"] == 0 and cnt["illegal"] == 0
    assert cnt["steps"] == n * moves and cnt["sims"] == n * moves * 8
    assert cnt["games"] == cnt["wins"][0] + cnt["wins"][1] + cnt["draws"] and cnt["games"] > n // 2
    # replay game 0 on the oracle, including the re-deals and the table clears
    rules = po.default_rules(mcts_simulations=8, threads_per_mcts=1)
    o, t = po.OracleGame(rules), po.OracleMcts(rules, "pseudo")
    o.new_game(SEED, 0, 0)
    for ply in range(moves):
        a = t.search(o, SEED, 0, ply)
        mv = t.pick(a["pi"], o.s.round <= rules.temperature_threshold, SEED, 0, ply)
        assert o.move(mv, SEED, 0, ply) == 0
        if o.status() != -1:
            o.new_game(SEED, 0, ply + 1)
            t.clear()
    assert (env.export_aos()[0] == o.data()).all()
    mc.close(); env.close()


def test_full_size_search_properties_and_shard_invariance(api):
    """BASELINE configs[2] at full size (4096 games x 64 simulations, 5-block network, bf16 tcgen05 forward): properties that need no
    oracle replay — visit counts live on legal moves only and add up to the root's sumN >= simulations, pi is the normalised count
    vector, the move played is legal (no illegal / overflow counters), every simulation is counted, the search is reproducible bit for
    bit, and the second half of the games searched as its own shard (first_game_id = 2048, the multi-GPU partitioning) reproduces the
    second half of the full run exactly."""
    n, sims, moves = 4096, 64, 3
    rules = api.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    net = api.Net(blocks=5, seed=1234)

    def run(count, first):
        env = api.Env(count, rules=rules, first_game_id=first)
        env.reset(SEED)
        mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
        out = []
        for _ in range(moves):
            valid = env.valid_moves()
            res = mc.search(pick_mode=api.PICK_SELFPLAY, apply_move=True)
            out.append((valid, res["N"].copy(), res["pi"].copy(), res["move"].copy()))
        cnt = mc.counters()
        stats = mc.root_stats()
        final = env.export_aos().copy()
        mc.close(); env.close()
        return out, cnt, final, stats

    full, cnt, final, _ = run(n, 0)
    assert cnt["errors"] == 0 and cnt["illegal"] == 0 and cnt["sims"] == n * sims * moves and cnt["steps"] == n * moves
    for valid, N, pi, move in full:
        legal = ((valid[:, None] >> np.arange(43, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        assert (N[~legal] == 0).all()
        tot = N.sum(1)
        assert (tot >= sims - 1).all()                       # the root's own expansion is not a visit of a child
        assert np.allclose(pi, N / tot[:, None], atol=1e-6)
        assert legal[np.arange(n), move].all()
    # 64 sampled global game ids replayed on the oracle search; its evaluator calls the SAME bf16 tcgen05 forward on a batch of
    # one (the tower is batch-invariant), so both sides see identical network outputs: N, pi and moves bit for bit at full size
    L = po.oracle_lib()

    @C.CFUNCTYPE(None, C.POINTER(po.RoState), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)
    def evaluator(sp, policy, value, user):
        x = np.zeros(po.INPUT_FLOATS, np.float32)
        L.ro_encode(sp, x)
        p, v = net.forward(x.reshape(1, -1), api.BF16)
        C.memmove(policy, p.ctypes.data, 43 * 4)
        value[0] = float(v[0])

    rules_o = po.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    for g in np.random.default_rng(11).choice(n, 64, replace=False):
        g = int(g)
        o, t = po.OracleGame(rules_o), po.OracleMcts(rules_o, "pseudo")
        t.L.ro_mcts_free(t.h)
        t.h = t.L.ro_mcts_new(C.cast(evaluator, C.c_void_p), None)
        o.new_game(SEED, g, 0)
        for ply, (valid, N, pi, move) in enumerate(full):
            assert int(valid[g]) == o.valid()
            a = t.search(o, SEED, g, ply)
            assert (N[g] == a["N"]).all(), (g, ply, N[g], a["N"])
            assert (bits(pi[g]) == bits(a["pi"])).all(), (g, ply)
            mv = t.pick(a["pi"], o.s.round <= rules_o.temperature_threshold, SEED, g, ply)
            assert move[g] == mv and o.move(mv, SEED, g, ply) == 0
        assert (final[g] == o.data()).all(), g
    again, cnt2, final2, _ = run(n, 0)
    assert cnt2 == cnt and (final2 == final).all()
    for (v1, n1, p1, m1), (v2, n2, p2, m2) in zip(full, again):
        assert (n1 == n2).all() and (m1 == m2).all() and (p1.view(np.uint32) == p2.view(np.uint32)).all()
    half, _, final_h, _ = run(n // 2, n // 2)
    assert (final_h == final[n // 2:]).all()
    for (v1, n1, p1, m1), (v2, n2, p2, m2) in zip(full, half):
        assert (n1[n // 2:] == n2).all() and (m1[n // 2:] == m2).all()
    net.close()


def test_fused_state_packing_matches_the_tensor_route(api):
    """the leaf batch of a bf16 search reaches the stem through ONE kernel that packs bf16 operands straight from the game states
    (k_nn_pack_state_tc); az_nn_forward on an fp32 input tensor takes the other route (k_nn_pack_input_tc on what the reference's
    encoder produces).  Same expressions, same roundings: a device search (first route) and an oracle search whose evaluator calls
    the network on the oracle's own encoding (second route, batch of one) must agree on visit counts, pi and moves, bit for bit.
    Mid-game positions (every channel of the encoding is live), 70 positions = more than one 128-row tile, ragged."""
    n, first, sims, start = 70, 900, 12, 90
    rules = api.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    rules_o = po.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    net = api.Net(blocks=2, seed=77)
    L = po.oracle_lib()

    @C.CFUNCTYPE(None, C.POINTER(po.RoState), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)
    def evaluator(sp, policy, value, user):
        x = np.zeros(po.INPUT_FLOATS, np.float32)
        L.ro_encode(sp, x)
        p, v = net.forward(x.reshape(1, -1), api.BF16)
        C.memmove(policy, p.ctypes.data, 43 * 4)
        value[0] = float(v[0])

    env = api.Env(n, rules=rules, first_game_id=first)
    env.reset(SEED)
    env.rollout(start)                                            # out of the set-up phase
    img = env.export_aos()
    mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
    dev = []
    for _ in range(4):
        r = mc.search(pick_mode=api.PICK_SELFPLAY, apply_move=True)
        dev.append((r["N"].copy(), r["pi"].copy(), r["move"].copy()))
    assert mc.counters()["errors"] == 0
    final = env.export_aos()
    for g in range(0, n, 3):
        o, t = po.OracleGame(rules_o), po.OracleMcts(rules_o, "pseudo")
        t.L.ro_mcts_free(t.h)
        t.h = t.L.ro_mcts_new(C.cast(evaluator, C.c_void_p), None)
        assert o.set_data(img[g]) == 0
        for k, (N, pi, move) in enumerate(dev):
            if o.status() != -1:
                assert move[g] == 43
                continue
            a = t.search(o, SEED, first + g, start + k)
            assert (N[g] == a["N"]).all(), (g, k)
            assert (bits(pi[g]) == bits(a["pi"])).all(), (g, k)
            mv = t.pick(a["pi"], o.s.round <= rules_o.temperature_threshold, SEED, first + g, start + k)
            assert move[g] == mv and o.move(mv, SEED, first + g, start + k) == 0
        assert (final[g] == o.data()).all(), g
    mc.close(); env.close(); net.close()


@pytest.mark.parametrize("n", [70, 2500])
def test_two_cohorts_change_nothing(api, n):
    """az_mcts_set_cohorts: the games of a search split into two cohorts on two CUDA streams (one cohort's tree kernel / packing /
    stem under the other's tower, each with its own set of tower work buffers).  Games are independent: visit counts, pi, moves,
    counters and final states must equal the one-stream search bit for bit.  70 games force the cohort path below its automatic
    threshold (ragged halves, less than one tile pair each); 2500 games take it automatically."""
    sims = 10
    rules = api.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    net = api.Net(blocks=2, seed=31)
    out = {}
    for mode in (1, 2 if n < 2048 else 0):
        env = api.Env(n, rules=rules, first_game_id=5000)
        env.reset(SEED)
        env.rollout(120)
        mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
        mc.set_cohorts(mode)
        res = []
        for _ in range(3):
            r = mc.search(pick_mode=api.PICK_SELFPLAY, apply_move=True)
            res.append((r["N"].copy(), r["pi"].copy(), r["move"].copy()))
        mc.selfplay(2)
        out[mode] = (res, mc.counters(), env.export_aos().copy())
        assert out[mode][1]["errors"] == 0
        mc.close(); env.close()
    (a, ca, fa), (b, cb, fb) = out.values()
    for (n1, p1, m1), (n2, p2, m2) in zip(a, b):
        assert (n1 == n2).all() and (bits(p1) == bits(p2)).all() and (m1 == m2).all()
    assert ca == cb and (fa == fb).all()
    net.close()
