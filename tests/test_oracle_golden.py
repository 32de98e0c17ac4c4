"""CPU suite: the C oracle (oracle/risk_oracle.c) against the golden fixtures that the
COMPILED REFERENCE produced (tests/golden/gen_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

SEED = 0x5EED0001


@pytest.fixture(scope="module")
def env_trace(golden_dir):
    return np.load(os.path.join(golden_dir, "env_trace.npz"))


def test_data_layout_sizes():
    m = po.data_byte_mask()
    assert m.shape == (160,) and int(m.sum()) == 42 + 2 * (5 * 6 + 2 + 1) + 12


def test_deal_matches_reference(env_trace):
    mask = po.data_byte_mask()
    seed = int(env_trace["seed"])
    for g, (draws, dealt) in enumerate(zip(env_trace["deal_draws"], env_trace["dealt"])):
        o = po.OracleGame()
        o.new_game_tape(draws)
        assert (o.data()[mask] == dealt[mask]).all()
        o2 = po.OracleGame()
        o2.new_game(seed, g, 0)           # the Philox contract reproduces the same draws
        assert (o2.data() == o.data()).all()
        owners = o.data()[:42] >> 6
        assert [(owners == p).sum() for p in range(3)] == [14, 14, 14]
        assert (o.data()[:42] & 63 == 1).all() and o.s.reinf == 52 and o.s.phase == 0 and o.s.round == 1


def test_env_trace_bit_exact(env_trace):
    """every transition of the reference trace: state -> mask, (action, dice) -> next state, status"""
    mask = po.data_byte_mask()
    o = po.OracleGame()
    t = env_trace
    phases = np.zeros(6, int)
    for i in range(len(t["action"])):
        assert o.set_data(t["before"][i]) == 0
        assert (o.data() == t["before"][i]).all()      # export(import(x)) == x incl. all derived masks
        assert o.valid() == int(t["valid"][i])
        phases[o.s.phase] += 1
        rc, used = o.move_tape(int(t["action"][i]), t["dice"][i].astype(np.int32))
        assert rc == 0 and used == int(t["n_dice"][i])
        assert (o.data()[mask] == t["after"][i][mask]).all(), i
        assert o.status() == int(t["status"][i])
    assert (phases > 0).all()                         # all six phases are covered
    assert (t["status"] == -2).sum() + (t["status"] == 0).sum() + (t["status"] == 1).sum() == len(t["deal_draws"])


def test_env_trace_philox_contract(env_trace):
    """the same trace replayed with dice and actions taken from the Philox contract"""
    t = env_trace
    seed = int(t["seed"])
    o = po.OracleGame()
    for i in range(len(t["action"])):
        g, ply = int(t["game"][i]), int(t["ply"][i])
        if ply == 0:
            o.new_game(seed, g, 0)
        assert o.random_action(seed, g, ply) == int(t["action"][i])
        assert o.move(int(t["action"][i]), seed, g, ply) == 0
        assert (o.data() == t["after"][i]).all()


def test_illegal_actions_rejected(env_trace):
    t = env_trace
    o = po.OracleGame()
    for i in range(0, len(t["action"]), 37):
        o.set_data(t["before"][i])
        valid = int(t["valid"][i])
        for a in range(44):
            if a < 43 and (valid >> a) & 1:
                continue
            before = o.data()
            rc, _ = o.move_tape(a, np.array([6, 6, 6, 1, 1], np.int32))
            assert rc == -1
            assert (o.data() == before).all()
    # a finished game accepts nothing
    last = int(np.nonzero(t["status"] != -1)[0][0])
    o.set_data(t["after"][last])
    rc, _ = o.move_tape(42, np.array([1, 1, 1, 1, 1], np.int32))
    assert rc == -2


def test_encoding_exact(golden_dir):
    e = np.load(os.path.join(golden_dir, "encode.npz"))
    o = po.OracleGame()
    for d, x in zip(e["data"], e["x"]):
        assert o.set_data(d) == 0
        got = o.encode().reshape(7, 6, 13)
        assert (got.view(np.uint32) == x.view(np.uint32)).all()


def test_normalize_policy():
    rng = np.random.default_rng(0)
    L = po.oracle_lib()
    for _ in range(50):
        p = rng.random(43).astype(np.float32)
        valid = int(rng.integers(1, 1 << 43))
        q = p.copy()
        L.ro_normalize_policy(q, valid)
        s = np.float32(0)
        for i in range(43):
            if (valid >> i) & 1:
                s = np.float32(s + p[i])
        for i in range(43):
            exp = np.float32(p[i] / s) if (valid >> i) & 1 and p[i] > 0 else np.float32(0)
            assert q[i] == exp


@pytest.mark.parametrize("name", ["selfplay16", "selfplay64", "play32t2", "selfplay200", "selfplay800"])
def test_mcts_trace_bit_exact(golden_dir, name):
    t = np.load(os.path.join(golden_dir, "mcts_trace_deep.npz" if name in ("selfplay200", "selfplay800") else "mcts_trace.npz"))
    sims, T, play_mode, g = [int(v) for v in t[name + "_cfg"]]
    seed = int(t["seed"])
    rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=T)
    o, m = po.OracleGame(rules), po.OracleMcts(rules, "pseudo")
    o.new_game(seed, g, 0)
    n = len(t[name + "_move"])
    for ply in range(n):
        assert (o.data() == t[name + "_root"][ply]).all()
        if t[name + "_trimmed"][ply]:
            m.trim()
        res = m.search(o, seed, g, ply)
        assert (res["N"] == t[name + "_N"][ply]).all(), ply
        for k in ("Q", "P", "pi"):
            assert (res[k].view(np.uint32) == t[name + "_" + k][ply]).all(), (ply, k)
        assert res["sumN"] == int(t[name + "_sumN"][ply])
        assert np.float32(res["value"]).view(np.uint32) == t[name + "_value"][ply]
        assert m.table_size() == int(t[name + "_table"][ply])
        sample = (not play_mode) and o.s.round <= rules.temperature_threshold
        mv = m.pick(res["pi"], sample, seed, g, ply)
        assert mv == int(t[name + "_move"][ply])
        assert o.move(mv, seed, g, ply) == 0
    assert (o.data() == t[name + "_final"]).all()
    assert o.status() == int(t[name + "_status"])


@pytest.mark.parametrize("pairing", ["script_vs_script", "script_vs_random"])
def test_turn_samples_golden(golden_dir, pairing):
    """Player::addTrainingSample inside scripted / random turns: the oracle's recorder (ro_script_turn_rec / ro_random_turn_rec)
    against the bytes the REFERENCE's own players and NNTrainDataStorage::saveTrainingSamples wrote for the same games
    (tests/golden/turn_samples.npz, generator tests/golden/gen_turn_samples.py; runs without /root/reference)."""
    z = np.load(os.path.join(golden_dir, "turn_samples.npz"))
    games = sorted(int(k.split("_")[-2]) for k in z.files if k.startswith(pairing) and k.endswith("_records"))
    assert len(games) == 2
    for game in games:
        want = z["%s_%d_records" % (pairing, game)]
        status, plies, n = (int(v) for v in z["%s_%d_meta" % (pairing, game)])
        orc = po.OracleGame()
        orc.new_game(SEED, game, 0)
        sps, sink, ply = [po.new_script(), po.new_script()], po.TurnSink(4096), 0
        while orc.status() == -1:
            cur = orc.s.cur
            if pairing == "script_vs_script" or cur == 0:
                assert orc.script_turn_rec(sps[cur], SEED, game, ply, sink) == 0
            else:
                assert orc.random_turn_rec(SEED, game, ply, sink) == 0
            ply += 1
        assert (orc.status(), ply, len(sink)) == (status, plies, n)
        assert (sink.records(status, orc) == want).all()
