"""GPU suite for SURVEY.md 8f N1 + N2: the device-side scripted opponent against the oracle's ScriptPlayer restatement (itself pinned
to the compiled reference, tests/test_oracle_vs_ref.py) and the `-m play` arena (AlphaZero vs Script, mirror pairs, GameResults)
against an oracle replay of the same match."""
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


def test_script_vs_script_lockstep(api):
    """every game: both sides scripted, one whole turn per call, all 160 Data bytes equal to the oracle after every turn"""
    n, first = 96, 7
    env = api.Env(n, first_game_id=first)
    env.reset(SEED)
    script = np.full((n, 2), api.SCRIPT_INIT, np.uint32)
    games = [po.OracleGame() for _ in range(n)]
    sps = [[po.new_script(), po.new_script()] for _ in range(n)]
    for g, o in enumerate(games):
        o.new_game(SEED, first + g, 0)
    turns = 0
    for ply in range(120):
        st = env.script_turn(script)
        dev = env.export_aos()
        for g, o in enumerate(games):
            if o.status() != -1:
                assert st[g] == -4
                continue
            assert o.script_turn(sps[g][o.s.cur], SEED, first + g, ply) == 0
            turns += 1
            assert st[g] == o.status()
            assert (dev[g] == o.data()).all(), "game %d differs after turn %d" % (g, ply)
    assert turns > n * 30 and all(o.status() != -1 for o in games)
    env.close()


def test_random_vs_random_lockstep(api):
    """RandomPlayer::takeTurn on the device, both sides, against the oracle restatement: all Data bytes after every turn"""
    n, first = 64, 2000
    env = api.Env(n, first_game_id=first)
    env.reset(SEED)
    games = [po.OracleGame() for _ in range(n)]
    for g, o in enumerate(games):
        o.new_game(SEED, first + g, 0)
    turns = 0
    for ply in range(110):
        st = env.random_turn()
        dev = env.export_aos()
        for g, o in enumerate(games):
            if o.status() != -1:
                assert st[g] == -4
                continue
            assert o.random_turn(SEED, first + g, ply) == 0
            turns += 1
            assert st[g] == o.status()
            assert (dev[g] == o.data()).all(), "game %d differs after turn %d" % (g, ply)
    assert turns > n * 40
    env.close()


def records_of(staged, status, rules):
    """the 265-byte records of one game's staged (state, pi) pairs once the game ended with `status` (NNTrainDataStorage::updateValues)"""
    out = []
    for st, pi in staged:
        o2 = po.OracleGame(rules)
        o2.s = st
        out.append(o2.sample_record(pi, status))
    return np.stack(out).tobytes()


def match_record_stream(recs, expected):
    """every game's records are contiguous in the device queue; games arrive in a data-dependent order"""
    assert sum(len(e) for e in expected) == recs.size
    blob, left, pos = recs.tobytes(), sorted(expected, key=len, reverse=True), 0
    while pos < len(blob):
        hit = [e for e in left if blob.startswith(e, pos)]
        assert hit, "record stream at byte %d matches no expected game" % pos
        left.remove(hit[0]); pos += len(hit[0])
    assert not left


def oracle_pair(slot_game, sims, ply0=0, opponent="script", oracle_eval=None, K=1):
    """one claimed pair on one slot, replayed on the oracle: AlphaZero (player 0, play mode, pseudo evaluator) vs Script (player 1),
    fresh deal then the mirror game (Game::newGame, game/game.cpp:170-191); returns GameResults-style tallies, the final Data
    image and the ply counter"""
    rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    o, tree, sp = po.OracleGame(rules), po.OracleMcts(rules, "pseudo"), po.new_script()
    if oracle_eval is not None:            # the search's evaluator = a caller-supplied function (e.g. the CUDA network, batch of one)
        tree.L.ro_mcts_free(tree.h)
        tree.h = tree.L.ro_mcts_new(C.cast(oracle_eval, C.c_void_p), None)
    res = dict(count=0, draw=0, win=[0, 0], was=[0, 0], az_moves=0, opp_turns=0, samples=[])
    ply = ply0
    start = None
    for player_start in (0, 1):
        if player_start == 0:
            o.new_game(SEED, slot_game, ply)
            start = po.RoState.from_buffer_copy(o.s)
        else:
            o.s = po.RoState.from_buffer_copy(start)
            o.invert_players()
            o.s.cur = 1
        tree.clear()
        last, staged = None, []
        while o.status() == -1:
            if o.s.cur == 1:
                assert (o.script_turn(sp, SEED, slot_game, ply) if opponent == "script" else o.random_turn(SEED, slot_game, ply)) == 0
                res["opp_turns"] += 1
                last = 1
            else:
                if last != 0:
                    tree.trim()            # AlphaZeroPlayer::takeTurn: trimNodes when its turn starts
                a = tree.search(o, SEED, slot_game, ply, lockstep=K)
                staged.append((po.RoState.from_buffer_copy(o.s), a["pi"].copy()))      # AlphaZeroPlayer::takeTurn pushes before the move
                mv = tree.pick(a["pi"], False, SEED, slot_game, ply)
                assert o.move(mv, SEED, slot_game, ply) == 0
                res["az_moves"] += 1
                last = 0
            ply += 1
        st = o.status()
        if staged:
            res["samples"].append(records_of(staged, st, rules))
        res["count"] += 1
        if st == -2:
            res["draw"] += 1
        else:
            res["win"][st] += 1
            res["was"][st] += int(st == player_start)
    return res, o.data(), ply


@pytest.mark.parametrize("opponent", ["script", "random"])
def test_arena_one_pair_per_slot_matches_oracle(api, opponent):
    """n slots, 2n games: every slot claims exactly one mirror pair, so the whole match is deterministic and must equal the oracle replay"""
    n, first, sims = (10, 40, 8) if opponent == "script" else (6, 90, 8)
    env = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1), first_game_id=first)
    mc = api.Mcts(env, evaluator=api.EVAL_PSEUDO)
    arena = api.Arena(mc, api.OPPONENT_SCRIPT if opponent == "script" else api.OPPONENT_RANDOM, mirror_games=True)
    mc.record(capacity_samples=n * 2 * 700, max_moves_per_game=1024)      # playGames(pg1, pg2, games, trainStorage): AlphaZero's samples
    r = arena.play(2 * n, SEED)
    assert r["errors"] == 0 and r["count"] == 2 * n
    dev = env.export_aos()
    recs, dropped = mc.samples()
    assert dropped == 0
    tot = dict(count=0, draw=0, win=[0, 0], was=[0, 0], az_moves=0, opp_turns=0)
    expected = []
    for g in range(n):
        res, data, _ = oracle_pair(first + g, sims, opponent=opponent)
        expected += res["samples"]
        assert (dev[g] == data).all(), "slot %d: final position differs from the oracle replay" % g
        for k in ("count", "draw", "az_moves", "opp_turns"):
            tot[k] += res[k]
        for i in range(2):
            tot["win"][i] += res["win"][i]; tot["was"][i] += res["was"][i]
    assert (r["count"], r["draw"], r["win"], r["win_and_started"]) == (tot["count"], tot["draw"], tot["win"], tot["was"])
    assert r["az_moves"] == tot["az_moves"] and r["opponent_turns"] == tot["opp_turns"] and r["az_sims"] == tot["az_moves"] * sims
    # games that end on the opponent's turn are flushed by the arena (Player::gameFinished), the others by the searcher itself
    match_record_stream(recs, expected)
    arena.close(); mc.close(); env.close()


def oracle_versus_pair(slot_game, sims, evaluators, K=1):
    """one claimed pair on one slot with an AlphaZeroPlayer on each side (GameGroup::playGames(trainAZPG, generateAZPG, ...),
    alphazero_trainer.cpp:147-166), replayed on the oracle: each side owns a search table (cleared at a new game, trimmed when its
    turn starts), play mode (argmax) on both sides"""
    rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    o = po.OracleGame(rules)
    trees = [po.OracleMcts(rules, evaluators[0]), po.OracleMcts(rules, evaluators[1])]
    res = dict(count=0, draw=0, win=[0, 0], was=[0, 0], moves=[0, 0], samples=[[], []])
    ply, start = 0, None
    for player_start in (0, 1):
        if player_start == 0:
            o.new_game(SEED, slot_game, ply)
            start = po.RoState.from_buffer_copy(o.s)
        else:
            o.s = po.RoState.from_buffer_copy(start)
            o.invert_players()
            o.s.cur = 1
        for t in trees:
            t.clear()
        last, staged = None, [[], []]
        while o.status() == -1:
            side = int(o.s.cur)
            if last != side:
                trees[side].trim()
            a = trees[side].search(o, SEED, slot_game, ply, lockstep=K)
            staged[side].append((po.RoState.from_buffer_copy(o.s), a["pi"].copy()))
            mv = trees[side].pick(a["pi"], False, SEED, slot_game, ply)
            assert o.move(mv, SEED, slot_game, ply) == 0
            res["moves"][side] += 1
            last = side
            ply += 1
        st = o.status()
        for side in (0, 1):
            if staged[side]:
                res["samples"][side].append(records_of(staged[side], st, rules))
        res["count"] += 1
        if st == -2:
            res["draw"] += 1
        else:
            res["win"][st] += 1
            res["was"][st] += int(st == player_start)
    return res, o.data()


@pytest.mark.parametrize("n,first,sims,K,evaluators", [(6, 300, 8, 1, ("pseudo", "uniform")), (5, 17, 6, 2, ("uniform", "pseudo")),
                                                         (4, 950, 8, 1, ("pseudo", "pseudo"))])
def test_arena_alphazero_vs_alphazero_matches_oracle(api, n, first, sims, K, evaluators):
    """the trainer's comparison match on the device: two searchers (own tables, own evaluators) over one set of game states; every slot
    plays one mirror pair, so final positions, tallies and per-side move counts must equal the oracle replay bit for bit"""
    kinds = dict(pseudo=api.EVAL_PSEUDO, uniform=api.EVAL_UNIFORM)
    env = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1, concurrent_descents=K), first_game_id=first)
    mc0, mc1 = api.Mcts(env, evaluator=kinds[evaluators[0]]), api.Mcts(env, evaluator=kinds[evaluators[1]])
    arena = api.Arena(mc0, mirror_games=True, opponent_mcts=mc1)
    for mc in (mc0, mc1):                              # INCLUDE_COMPARE_GAMES_TRAIN_SAMPLES: both players collect their samples
        mc.record(capacity_samples=n * 2 * 700, max_moves_per_game=1024)
    r = arena.play(2 * n, SEED)
    assert r["errors"] == 0 and r["count"] == 2 * n
    dev = env.export_aos()
    got = [mc0.samples(), mc1.samples()]
    tot = dict(count=0, draw=0, win=[0, 0], was=[0, 0], moves=[0, 0])
    expected = [[], []]
    for g in range(n):
        res, data = oracle_versus_pair(first + g, sims, evaluators, K)
        for side in (0, 1):
            expected[side] += res["samples"][side]
        assert (dev[g] == data).all(), "slot %d: final position differs from the oracle replay" % g
        tot["count"] += res["count"]; tot["draw"] += res["draw"]
        for i in range(2):
            tot["win"][i] += res["win"][i]; tot["was"][i] += res["was"][i]; tot["moves"][i] += res["moves"][i]
    assert (r["count"], r["draw"], r["win"], r["win_and_started"]) == (tot["count"], tot["draw"], tot["win"], tot["was"])
    assert r["az_moves"] == tot["moves"][0] and r["opponent_turns"] == tot["moves"][1] and r["az_sims"] == tot["moves"][0] * sims
    for side in (0, 1):
        assert got[side][1] == 0
        match_record_stream(got[side][0], expected[side])
    # the searchers are usable on their own again afterwards (the side selection is lifted)
    env.reset(SEED)
    assert mc0.search(pick_mode=api.PICK_ARGMAX, apply_move=False)["N"].sum() > 0
    other = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1), first_game_id=first)
    mc2 = api.Mcts(other, evaluator=api.EVAL_UNIFORM)
    with pytest.raises(api.AzError):
        api.Arena(mc0, opponent_mcts=mc2)              # different env
    with pytest.raises(api.AzError):
        api.Arena(mc0, opponent_mcts=mc0)              # the same handle twice
    arena.close(); mc0.close(); mc1.close(); mc2.close(); other.close(); env.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_arena_with_network_matches_oracle_replay(api, precision):
    """with a network evaluator the arena sends only the slots that search this tick through the network (gather -> forward ->
    scatter: the leaf batch shrinks and changes composition as slots finish at different times).  The oracle replays every slot's
    pair with an evaluator that calls the SAME network on a batch of one (the kernels are batch-invariant), so the match must be
    identical move for move: final positions, GameResults tallies, move and simulation counts.  K = 2 descents per tree."""
    prec = api.FP32 if precision == "fp32" else api.BF16
    n, first, sims, K = 9, 70, 6, 2
    rules = api.default_rules(mcts_simulations=sims, threads_per_mcts=1, concurrent_descents=K)
    net = api.Net(blocks=1, seed=5)
    L = po.oracle_lib()

    @C.CFUNCTYPE(None, C.POINTER(po.RoState), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)
    def evaluator(sp, policy, value, user):
        x = np.zeros(po.INPUT_FLOATS, np.float32)
        L.ro_encode(sp, x)
        p, v = net.forward(x.reshape(1, -1), prec)
        C.memmove(policy, p.ctypes.data, 43 * 4)
        value[0] = float(v[0])

    env = api.Env(n, rules=rules, first_game_id=first)
    mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=prec)
    arena = api.Arena(mc, api.OPPONENT_SCRIPT, mirror_games=True)
    r = arena.play(2 * n, SEED)
    assert r["errors"] == 0 and r["count"] == 2 * n
    dev = env.export_aos()
    tot = dict(count=0, draw=0, win=[0, 0], was=[0, 0], az_moves=0, opp_turns=0)
    lengths = []
    for g in range(n):
        res, data, ply = oracle_pair(first + g, sims, oracle_eval=evaluator, K=K)
        lengths.append(ply)
        assert (dev[g] == data).all(), "slot %d: final position differs from the oracle replay" % g
        for k in ("count", "draw", "az_moves", "opp_turns"):
            tot[k] += res[k]
        for i in range(2):
            tot["win"][i] += res["win"][i]; tot["was"][i] += res["was"][i]
    assert (r["count"], r["draw"], r["win"], r["win_and_started"]) == (tot["count"], tot["draw"], tot["win"], tot["was"])
    assert r["az_moves"] == tot["az_moves"] and r["opponent_turns"] == tot["opp_turns"] and r["az_sims"] == tot["az_moves"] * sims
    assert len(set(lengths)) > 1            # slots finished at different times: the compacted batch really shrank
    arena.close(); mc.close(); env.close(); net.close()


def test_arena_claims_pairs_like_the_counter(api):
    """more games than slots, odd request: 2 * floor(n / 2) games are played (Counter::hasNext(2)), tallies are consistent"""
    n, sims = 6, 4
    env = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1), first_game_id=0)
    mc = api.Mcts(env, evaluator=api.EVAL_UNIFORM)
    arena = api.Arena(mc, api.OPPONENT_SCRIPT, mirror_games=True)
    r = arena.play(27, SEED)
    assert r["errors"] == 0 and r["count"] == 26
    assert r["draw"] + r["win"][0] + r["win"][1] == 26
    assert r["win_and_started"][0] <= r["win"][0] and r["win_and_started"][1] <= r["win"][1]
    assert r["az_sims"] == r["az_moves"] * sims and r["opponent_turns"] > 26 * 13
    r2 = arena.play(4, SEED + 1)          # the handle can be reused for another match
    assert r2["count"] == 4 and r2["errors"] == 0
    arena.close(); mc.close(); env.close()


@pytest.mark.parametrize("pairing", ["script_vs_script", "script_vs_random", "random_vs_random", "script_vs_random_unit_move_1"])
def test_scripted_turn_samples(api, pairing):
    """Player::addTrainingSample inside scripted / random turns, recorded on the device (az_env_record_turns): every finished game's
    records — state image, one-hot policy, value from updateValues — equal the oracle's (which is pinned byte for byte to the
    reference's own players and NNTrainDataStorage, tests/test_samples.py), and the board states stay in lockstep while recording
    (the recording path applies reinforcement / mobilisation in MIN_UNIT_MOVE steps, the plain path in one)"""
    n, first = 48, 4100
    kinds = {"script_vs_script": (api.OPPONENT_SCRIPT, api.OPPONENT_SCRIPT), "script_vs_random": (api.OPPONENT_SCRIPT, api.OPPONENT_RANDOM),
             "random_vs_random": (api.OPPONENT_RANDOM, api.OPPONENT_RANDOM),
             "script_vs_random_unit_move_1": (api.OPPONENT_SCRIPT, api.OPPONENT_RANDOM)}[pairing]
    # MIN_UNIT_MOVE sets the step of the recorded reinforcement / mobilisation moves: also with a non-default value
    rule_kw = {"min_unit_move": 1, "allow_yield": 0} if pairing.endswith("unit_move_1") else {}
    o_rules = po.default_rules(**rule_kw)
    env = api.Env(n, rules=api.default_rules(**rule_kw), first_game_id=first)
    env.reset(SEED)
    env.record_turns(capacity_samples=n * 6000, max_samples_per_game=8192)
    script = np.full((n, 2), api.SCRIPT_INIT, np.uint32)
    games = [po.OracleGame(o_rules) for _ in range(n)]
    sps = [[po.new_script(), po.new_script()] for _ in range(n)]
    sinks = [po.TurnSink(8192) for _ in range(n)]
    for g, o in enumerate(games):
        o.new_game(SEED, first + g, 0)
    expect = {}
    for ply in range(130):
        st = env.play_turn(kinds[0], kinds[1], script)
        dev = env.export_aos()
        for g, o in enumerate(games):
            if o.status() != -1:
                assert st[g] == -4
                continue
            cur = o.s.cur
            if kinds[cur] == api.OPPONENT_SCRIPT:
                assert o.script_turn_rec(sps[g][cur], SEED, first + g, ply, sinks[g]) == 0
            else:
                assert o.random_turn_rec(SEED, first + g, ply, sinks[g]) == 0
            assert st[g] == o.status()
            assert (dev[g] == o.data()).all(), "game %d differs after turn %d" % (g, ply)
            if o.status() != -1:
                expect[g] = sinks[g].records(o.status(), o)
    assert len(expect) > n // 2
    recs, dropped = env.turn_samples()
    assert dropped == 0 and len(recs) == sum(len(v) for v in expect.values()) and len(recs) > 5000
    # the queue holds whole games in the order their warps reserved space: match each game's block by content
    want = sorted(v.tobytes() for v in expect.values())
    got, i = [], 0
    blob = recs.tobytes()
    remaining = dict((k, v.tobytes()) for k, v in expect.items())
    while i < len(recs):
        hit = None
        for g, b in remaining.items():
            if blob.startswith(b, i * 265):
                hit = g
                break
        assert hit is not None, "record %d starts no expected game block" % i
        got.append(remaining.pop(hit))
        i += len(got[-1]) // 265
    assert sorted(got) == want and not remaining
    assert env.turn_samples()[0].shape[0] == 0              # drained
    env.close()
