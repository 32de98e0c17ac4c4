"""GPU suite (pytest -m gpu): the CUDA environment kernels, called through the C ABI, against
the CPU oracle and the golden fixtures of the compiled reference.  Bit-exact everywhere."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

SEED = 0x5EED0001
# the rule sets tests/test_oracle_vs_ref.py pins oracle-vs-reference, here CUDA-vs-oracle: LIMIT_ATTACK_MOVES,
# LIMIT_REINFORCEMENT_MOVES off, ALLOW_YIELD off, MAX_GAME_ROUNDS, MIN_UNIT_MOVE (alphazero_moves.cpp:18,36,59,
# state.cpp:518-565, settings.h:40-62)
RULE_SETS = [dict(), dict(limit_attack=1), dict(limit_reinforcement=0), dict(allow_yield=0, max_game_rounds=40),
             dict(min_unit_move=1), dict(limit_attack=1, limit_reinforcement=0, allow_yield=0, max_game_rounds=45, min_unit_move=5)]
RULE_IDS = ["default", "limit_attack", "no_limit_reinf", "no_yield_40_rounds", "min_unit_move_1", "all_non_default"]


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    return a


def oracle_states(n, seed, first=0):
    out = np.zeros((n, 160), np.uint8)
    o = po.OracleGame()
    for g in range(n):
        o.new_game(seed, first + g, 0)
        out[g] = o.data()
    return out


@pytest.mark.parametrize("n,first", [(1, 0), (129, 7), (1000, 123456)])
def test_reset_matches_oracle_deal(api, n, first):
    env = api.Env(n, first_game_id=first)
    env.reset(SEED)
    assert (env.export_aos() == oracle_states(n, SEED, first)).all()
    assert (env.status() == -1).all()
    env.close()


def test_golden_trace_one_step_per_game(api, golden_dir):
    """every transition of the reference trace as one game of a batch: import -> mask -> step(dice tape) -> export"""
    t = np.load(os.path.join(golden_dir, "env_trace.npz"))
    n = len(t["action"])
    env = api.Env(n)
    env.import_aos(t["before"])
    assert (env.export_aos() == t["before"]).all()
    assert (env.valid_moves() == t["valid"]).all()
    st = env.step(t["action"], dice=t["dice"])
    assert (st == t["status"]).all()
    assert (env.export_aos() == t["after"]).all()
    assert (env.status() == t["status"]).all()
    env.close()


def test_encode_matches_reference_tensor(api, golden_dir):
    e = np.load(os.path.join(golden_dir, "encode.npz"))
    env = api.Env(len(e["data"]))
    env.import_aos(e["data"])
    x = env.encode()
    assert (x.view(np.uint32) == e["x"].view(np.uint32)).all()
    env.close()


@pytest.mark.parametrize("kw", RULE_SETS, ids=RULE_IDS)
def test_lockstep_random_play_vs_oracle_philox(api, kw):
    """host-chosen legal actions, dice from the Philox contract on the device; legal masks (az_valid_moves), status and
    states compared as we go, under the default and every non-default rule set"""
    n, steps = 192, 420
    env = api.Env(n, rules=api.default_rules(**kw), first_game_id=1000)
    env.reset(SEED)
    games = [po.OracleGame(po.default_rules(**kw)) for _ in range(n)]
    for g, o in enumerate(games):
        o.new_game(SEED, 1000 + g, 0)
    ply = np.zeros(n, int)
    finished = 0
    for s in range(steps):
        valid = env.valid_moves()
        act = np.zeros(n, np.uint8)
        exp = np.zeros(n, np.int8)
        for g, o in enumerate(games):
            assert int(valid[g]) == o.valid()
            if o.status() != -1:
                act[g], exp[g] = 42, -4
                continue
            a = o.random_action(SEED, 1000 + g, int(ply[g]))
            assert o.move(a, SEED, 1000 + g, int(ply[g])) == 0
            ply[g] += 1
            act[g], exp[g] = a, o.status()
        st = env.step(act)
        assert (st == exp).all(), s
        if s % 30 == 0 or s == steps - 1:
            dev = env.export_aos()
            for g, o in enumerate(games):
                assert (dev[g] == o.data()).all(), (s, g)
    finished = sum(o.status() != -1 for o in games)
    assert finished > n // 4
    env.close()


def test_illegal_and_finished_games_are_flagged_and_untouched(api, golden_dir):
    t = np.load(os.path.join(golden_dir, "env_trace.npz"))
    idx = np.arange(0, len(t["action"]), 9)
    before, valid = t["before"][idx], t["valid"][idx]
    n = len(idx)
    env = api.Env(n)
    for a in (0, 13, 41, 42, 43, 200):
        env.import_aos(before)
        st = env.step(np.full(n, a, np.uint8), dice=np.full((n, 5), 3, np.uint8))
        legal = np.array([a < 43 and (int(v) >> a) & 1 for v in valid], bool)
        assert (st[~legal] == -3).all()
        assert (st[legal] >= -2).all()
        assert (env.export_aos()[~legal] == before[~legal]).all()
    over = t["after"][t["status"] != -1]
    env2 = api.Env(len(over))
    env2.import_aos(over)
    st = env2.step(np.full(len(over), 42, np.uint8))
    assert (st == -4).all() and (env2.export_aos() == over).all()
    env.close(); env2.close()


def test_import_rejects_inconsistent_image(api, golden_dir):
    t = np.load(os.path.join(golden_dir, "env_trace.npz"))
    img = t["before"][:8].copy()
    img[3, 48] ^= 1            # flip a bit of player 0's ownedLands mask
    env = api.Env(8)
    with pytest.raises(api.AzError):
        env.import_aos(img)
    env.close()


@pytest.mark.parametrize("kw", RULE_SETS, ids=RULE_IDS)
def test_rollout_matches_oracle_including_redeals(api, kw):
    """k_env_rollout (az_valid_moves_flat + az_move_flat, the headline kernel) against the oracle, all rule sets"""
    n, steps, first = 160, 700, 77
    env = api.Env(n, rules=api.default_rules(**kw), first_game_id=first)
    env.reset(SEED)
    env.rollout(steps // 2)
    env.rollout(steps - steps // 2)          # two launches == one: state and ply persist in HBM
    dev = env.export_aos()
    cnt = env.counters()
    games = wins0 = wins1 = draws = 0
    o = po.OracleGame(po.default_rules(**kw))
    for g in range(n):
        o.new_game(SEED, first + g, 0)
        ply = 0
        for _ in range(steps):
            st = o.status()
            if st != -1:
                games += 1; wins0 += st == 0; wins1 += st == 1; draws += st == -2
                o.new_game(SEED, first + g, ply)
            a = o.random_action(SEED, first + g, ply)
            assert o.move(a, SEED, first + g, ply) == 0
            ply += 1
        assert (dev[g] == o.data()).all(), g
    assert cnt["steps"] == n * steps
    assert (cnt["games"], cnt["wins"][0], cnt["wins"][1], cnt["draws"]) == (games, wins0, wins1, draws)
    assert games > n
    env.close()


@pytest.mark.parametrize("n,launches", [(1, (0, 1, 250)), (7, (3, 200)), (33, (150, 1, 0, 60)), (449, (90, 40)), (1000, (64, 65))])
def test_rollout_pool_sizes_and_launch_lengths(api, n, launches):
    """the pooled rollout (games queued per category inside a block) with block fills it is not tuned for — one game, fewer games than a
    warp, one game over a block (448 slots), a partly filled last block — and with launches of 0 / 1 / few moves: every game makes
    exactly the requested number of moves, in its own (game, ply) stream"""
    first = 4000
    env = api.Env(n, first_game_id=first)
    env.reset(SEED)
    for k in launches:
        env.rollout(k)
    steps = sum(launches)
    dev = env.export_aos()
    cnt = env.counters()
    games = 0
    o = po.OracleGame(po.default_rules())
    for g in range(n):
        o.new_game(SEED, first + g, 0)
        ply = 0
        for _ in range(steps):
            if o.status() != -1:
                games += 1
                o.new_game(SEED, first + g, ply)
            assert o.move(o.random_action(SEED, first + g, ply), SEED, first + g, ply) == 0
            ply += 1
        assert (dev[g] == o.data()).all(), g
    assert cnt["steps"] == n * steps and cnt["games"] == games
    env.close()


def test_full_size_rollout_properties(api):
    """BASELINE config 2 size: 65536 games; size-independent properties instead of a CPU replay"""
    n, steps = 65536, 1000
    env = api.Env(n)
    env.reset(SEED)
    env.rollout(steps)
    cnt = env.counters()
    assert cnt["steps"] == n * steps
    assert cnt["games"] == cnt["wins"][0] + cnt["wins"][1] + cnt["draws"]
    # ~320 steps per game in the reference probe (327.7)
    assert 2.3 * n < cnt["games"] < 3.9 * n
    assert abs(cnt["wins"][0] - cnt["wins"][1]) < 0.05 * cnt["games"]
    img = env.export_aos()
    # 64 sampled global game ids replayed on the oracle (re-deals included): bit-exact at full size
    o = po.OracleGame()
    for g in np.random.default_rng(7).choice(n, 64, replace=False):
        g = int(g)
        o.new_game(SEED, g, 0)
        for ply in range(steps):
            if o.status() != -1:
                o.new_game(SEED, g, ply)
            assert o.move(o.random_action(SEED, g, ply), SEED, g, ply) == 0
        assert (img[g] == o.data()).all(), g
    # every exported image is self-consistent: re-import runs the reference's consistencyCheck invariant
    env2 = api.Env(n)
    env2.import_aos(img)
    assert (env2.export_aos() == img).all()
    owners = img[:, :42] >> 6
    assert (owners < 3).all() and ((img[:, :42] & 63) <= 32).all() and ((img[:, :42] & 63) >= 1).all()
    # determinism: same seed, same result
    env3 = api.Env(n)
    env3.reset(SEED)
    env3.rollout(steps)
    assert (env3.export_aos() == img).all()
    # sharding independence: the second half of the games computed as its own shard
    half = api.Env(n // 2, first_game_id=n // 2)
    half.reset(SEED)
    half.rollout(steps)
    assert (half.export_aos() == img[n // 2:]).all()
    for e in (env, env2, env3, half):
        e.close()
