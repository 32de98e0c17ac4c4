"""CPU suite: the C oracle against the compiled reference run LIVE (oracle/_ref, built from
/root/reference by oracle/ref/build_ref.sh).  Skipped where the library is absent."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref/libref_oracle.so not built")
SEED = 0x5EED0001


def test_reference_struct_sizes_and_defaults():
    L = po.ref_lib()
    assert (L.ref_sizeof_data(), L.ref_sizeof_player_status(), L.ref_sizeof_nn_input()) == (160, 48, 88)
    d = np.zeros(11, np.float32)
    L.ref_get_default_settings(d)
    r = po.default_rules()
    assert list(d) == [r.mcts_simulations, r.threads_per_mcts, np.float32(r.cpuct), np.float32(r.dir_noise_value),
                       np.float32(r.dir_noise_epsi), r.allow_yield, r.limit_reinforcement, r.limit_attack,
                       r.temperature_threshold, r.max_game_rounds, r.min_unit_move]


def test_map_tables_match_reference():
    import ctypes as C
    nm, nl = np.zeros(42, np.uint64), np.zeros(42 * 6, np.int8)
    cm, cb = np.zeros(6, np.uint64), np.zeros(6, np.int32)
    po.ref_lib().ref_get_map(nm, nl, cm, cb)
    L = po.oracle_lib()
    assert list((C.c_uint64 * 42).in_dll(L, "RO_NBR_MASK")) == list(nm)
    assert list((C.c_int8 * 252).in_dll(L, "RO_NBR_LIST")) == list(nl)
    assert list((C.c_uint64 * 6).in_dll(L, "RO_CONTINENT_MASK")) == list(cm)
    assert list((C.c_int * 6).in_dll(L, "RO_CONTINENT_BONUS")) == list(cb)


RULE_SETS = [dict(), dict(allow_yield=0), dict(limit_reinforcement=0), dict(limit_attack=1),
             dict(allow_yield=0, max_game_rounds=40), dict(min_unit_move=1),
             dict(limit_attack=1, limit_reinforcement=0, allow_yield=0, max_game_rounds=45, min_unit_move=5)]


@pytest.mark.parametrize("kw", RULE_SETS)
def test_random_games_lockstep(kw):
    seed = 0xC0FFEE
    rules = po.default_rules(**kw)
    po.ref_apply_rules(rules)
    mask = po.data_byte_mask()
    o, r = po.OracleGame(rules), po.RefGame()
    steps = 0
    for g in range(40):
        o.new_game(seed, g)
        r.new_game(seed, g)
        ply = 0
        while True:
            assert (o.data()[mask] == r.data()[mask]).all()
            assert o.valid() == r.valid() and o.status() == r.status()
            if ply % 11 == 0:
                assert (o.encode().view(np.uint32) == r.encode().view(np.uint32)).all()
            if o.status() != -1:
                break
            a = o.random_action(seed, g, ply)
            assert a == r.random_action(seed, g, ply)
            assert o.move(a, seed, g, ply) == 0 and r.move(a, seed, g, ply) == 0
            assert r.violations() == 0
            ply += 1
        steps += ply
    po.ref_apply_rules(po.default_rules())
    assert steps > 5000


# non-default search hyper-parameters (settings.h:40-62: CPUCT, DIR_NOISE_VALUE, DIR_NOISE_EPSI, TEMPERATURE_THRESHOLD) and
# game rules inside the search; tests/test_mcts_gpu.py runs the SAME sets CUDA-vs-oracle
MCTS_RULE_SETS = [dict(), dict(cpuct=2.5, dir_noise_value=0.05, dir_noise_epsi=0.5, temperature_threshold=6),
                  dict(cpuct=0.4, dir_noise_epsi=0.0, temperature_threshold=0, limit_attack=1, limit_reinforcement=0),
                  dict(dir_noise_value=1.0, dir_noise_epsi=1.0, allow_yield=0, max_game_rounds=40, min_unit_move=1)]


@pytest.mark.parametrize("sims,T,play_mode,kw", [(16, 1, False, 0), (48, 1, False, 0), (33, 2, True, 0),
                                                 (24, 1, False, 1), (20, 1, False, 2), (16, 1, True, 3), (32, 1, False, 3)])
def test_mcts_lockstep(sims, T, play_mode, kw):
    seed = 0xFACADE
    rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=T, **MCTS_RULE_SETS[kw])
    po.ref_apply_rules(rules)
    mask = po.data_byte_mask()
    o, r, om, rm = po.OracleGame(rules), po.RefGame(), po.OracleMcts(rules), po.RefMcts()
    o.new_game(seed, 3)
    r.new_game(seed, 3)
    ply, last = 0, None
    while o.status() == -1:
        if play_mode and o.s.cur != last:
            om.trim(); rm.trim(); last = o.s.cur
        a, b = om.search(o, seed, 3, ply), rm.search(r, seed, 3, ply)
        for k in ("N", "Q", "P", "pi"):
            assert (a[k].view(np.uint32) == b[k].view(np.uint32)).all(), (ply, k)
        assert a["sumN"] == b["sumN"] and a["value"] == b["value"] and om.table_size() == rm.table_size()
        sample = (not play_mode) and o.s.round <= rules.temperature_threshold
        mv = om.pick(a["pi"], sample, seed, 3, ply)
        assert mv == rm.pick(b["pi"], sample, seed, 3, ply)
        assert o.move(mv, seed, 3, ply) == 0 and r.move(mv, seed, 3, ply) == 0
        assert (o.data()[mask] == r.data()[mask]).all()
        ply += 1
    po.ref_apply_rules(po.default_rules())
    assert ply > 100


@pytest.mark.parametrize("sims,K,play_mode,evaluator", [(16, 2, True, "pseudo"), (48, 4, False, "pseudo"), (32, 8, False, "uniform"),
                                                        (24, 3, True, "uniform")])
def test_mcts_virtual_loss_lockstep(sims, K, play_mode, evaluator):
    """THREADS_PER_MCTS = K search threads of the UNMODIFIED reference (AlphaZeroMCTS::search, active_N rule of
    getNextBestMoveAndSetVisited, alphazero_mcts.cpp:67-119) made to take turns in the lockstep schedule ==
    ro_mcts_search_lockstep, bit for bit, over whole games; the duplicate-request branch must have fired"""
    seed = 0xFACADE
    rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=K)
    po.ref_apply_rules(rules)
    mask = po.data_byte_mask()
    o, r, om, rm = po.OracleGame(rules), po.RefGame(), po.OracleMcts(rules, evaluator), po.RefMcts(evaluator)
    o.new_game(seed, 5)
    r.new_game(seed, 5)
    ply, last = 0, None
    while o.status() == -1:
        if play_mode and o.s.cur != last:
            om.trim(); rm.trim(); last = o.s.cur
        a, b = om.search(o, seed, 5, ply, lockstep=K), rm.search(r, seed, 5, ply, lockstep=K)
        for k in ("N", "Q", "P", "pi"):
            assert (a[k].view(np.uint32) == b[k].view(np.uint32)).all(), (ply, k)
        assert a["sumN"] == b["sumN"] and a["value"] == b["value"] and om.table_size() == rm.table_size()
        sample = (not play_mode) and o.s.round <= rules.temperature_threshold
        mv = om.pick(a["pi"], sample, seed, 5, ply)
        assert mv == rm.pick(b["pi"], sample, seed, 5, ply)
        assert o.move(mv, seed, 5, ply) == 0 and r.move(mv, seed, 5, ply) == 0
        assert (o.data()[mask] == r.data()[mask]).all()
        ply += 1
    po.ref_apply_rules(po.default_rules())
    skips, dups = om.vl_counts()
    assert ply > 100 and skips > 0
    print("active_N rule: %d moves passed over, %d duplicate requests" % (skips, dups))


@pytest.mark.parametrize("mode", ["script_vs_script", "script_vs_random_mover", "mirror_pair"])
def test_script_player_lockstep(mode):
    """ro_script_turn == the UNMODIFIED ScriptPlayer::takeTurn (player/script/script_player.cpp:162-227): every Data field
    after every turn, with the scripted-opponent RNG streams of include/az_philox.h; also State::invertPlayers for the mirror game"""
    L = po.ref_lib()
    po.ref_apply_rules(po.default_rules())
    mask = po.data_byte_mask().astype(bool)
    turns = 0
    for g in range(18):
        ref, orc = po.RefGame(), po.OracleGame()
        ref.new_game(SEED, 100 + g, 0); orc.new_game(SEED, 100 + g, 0)
        if mode == "mirror_pair":            # Game::newGame second game of a pair: same deal, sides swapped, player 1 starts
            ref.invert_players(); ref.set_current_player(1)
            orc.invert_players(); orc.s.cur = 1
            assert (ref.data()[mask] == orc.data()[mask]).all()
        rs, os_ = [L.ref_script_new(), L.ref_script_new()], [po.new_script(), po.new_script()]
        ply = 0
        while ref.status() == -1 and ply < 450:
            cur = orc.s.cur
            if mode != "script_vs_random_mover" or cur == 1:
                assert ref.script_turn(rs[cur], SEED, 100 + g, ply) == 0, L.ref_last_error()
                assert orc.script_turn(os_[cur], SEED, 100 + g, ply) == 0
                turns += 1
            else:
                a = orc.random_action(SEED, 100 + g, ply)
                assert ref.move(a, SEED, 100 + g, ply) == 0 and orc.move(a, SEED, 100 + g, ply) == 0
            ply += 1
            assert (ref.data()[mask] == orc.data()[mask]).all(), "game %d ply %d" % (g, ply)
            assert ref.status() == orc.status() and ref.violations() == 0
        for h in rs:
            L.ref_script_free(h)
    assert turns > 300


@pytest.mark.parametrize("opponent", ["random_vs_random", "random_vs_script"])
def test_random_player_lockstep(opponent):
    """ro_random_turn == the UNMODIFIED RandomPlayer::takeTurn (player/random/random_player.cpp:22-111), every Data field after
    every turn, with the opponent RNG streams of include/az_philox.h"""
    L = po.ref_lib()
    po.ref_apply_rules(po.default_rules())
    mask = po.data_byte_mask().astype(bool)
    turns = 0
    for g in range(14):
        ref, orc = po.RefGame(), po.OracleGame()
        ref.new_game(SEED, 300 + g, 0); orc.new_game(SEED, 300 + g, 0)
        rp = [L.ref_random_new(0), L.ref_random_new(1)]
        rs, os_ = L.ref_script_new(), po.new_script()
        ply = 0
        while ref.status() == -1 and ply < 500:
            cur = orc.s.cur
            if opponent == "random_vs_random" or cur == 0:
                assert ref.random_turn(rp[cur], SEED, 300 + g, ply) == 0, L.ref_last_error()
                assert orc.random_turn(SEED, 300 + g, ply) == 0
                turns += 1
            else:
                assert ref.script_turn(rs, SEED, 300 + g, ply) == 0 and orc.script_turn(os_, SEED, 300 + g, ply) == 0
            ply += 1
            assert (ref.data()[mask] == orc.data()[mask]).all(), "game %d ply %d" % (g, ply)
            assert ref.status() == orc.status() and ref.violations() == 0
        for h in rp:
            L.ref_random_free(h)
        L.ref_script_free(rs)
    assert turns > 250
