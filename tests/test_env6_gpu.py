"""GPU suite of the SIX-PLAYER extension (BASELINE.json configs[3], rules: SIXPLAYER.md).  There is NO reference parity here — the
reference has no six-player game — so the checker is oracle/risk6_oracle.c, the plain-C statement of the same rule text; the CUDA
environment (az_env6_*) must agree with it bit for bit.  The CPU part checks the oracle's own invariants."""
import numpy as np
import pytest

from oracle import pyoracle as po

SEED = 0x5EED0001
RULE_SETS = [dict(), dict(limit_attack=1), dict(limit_reinforcement=0), dict(allow_yield=0, max_game_rounds=30), dict(min_unit_move=1)]
RULE_IDS = ["default", "limit_attack", "no_limit_reinf", "no_yield_30_rounds", "min_unit_move_1"]


def test_oracle6_invariants():
    """the extension's rule text, checked on the oracle alone: deal 7 lands x 6 seats, 78 set-up plies, armies conserved by
    everything but battles / reinforcements, owners 0..5, eliminated seats never move again, cards handed over on elimination"""
    o = po.Oracle6Game()
    eliminations = trades = 0
    for g in range(12):
        o.new_game(SEED, g, 0)
        assert sorted(np.bincount(np.array(o.s.owner[:]), minlength=6)) == [7] * 6 and sum(o.s.army[:]) == 42 and list(o.s.pool[:]) == [13] * 6
        ply, dead = 0, set()
        while o.status() == -1:
            assert o.s.cur not in dead
            if ply < 78:
                assert o.s.phase == 0 and o.s.cur == ply % 6
            elif ply == 78:
                assert o.s.phase == 2 and o.s.cur == 0 and sum(o.s.army[:]) == 120 and o.s.round == 14
            before_sets, before_cards = o.s.card_sets, sum(o.s.cards[:])
            a = o.random_action(SEED, g, ply)
            assert (o.valid() >> a) & 1
            assert o.move(a, SEED, g, ply) == 0
            ply += 1
            trades += o.s.card_sets > before_sets
            owners = set(o.s.owner[:])
            for p in range(6):
                if p not in owners and p not in dead:
                    dead.add(p); eliminations += 1
                    assert o.s.cards[p] == 0
            assert all(1 <= a <= 32 for a in o.s.army[:]) and all(x < 6 for x in o.s.owner[:])
            assert sum(o.s.cards[:]) in (before_cards, before_cards + 1, before_cards - 3, before_cards - 2)
        assert o.status() in (-2, 0, 1, 2, 3, 4, 5)
        assert o.move(42, SEED, g, ply) == -2                  # a finished game refuses moves
    assert eliminations > 5 and trades > 50


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("n,first", [(1, 0), (300, 5000)])
def test_deal_matches_oracle(api, n, first):
    env = api.Env6(n, first_game_id=first)
    env.reset(SEED)
    img = env.export()
    o = po.Oracle6Game()
    for g in range(n):
        o.new_game(SEED, first + g, 0)
        assert (img[g] == o.image()).all(), g
    v, st = env.query()
    assert (st == -1).all()
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", RULE_SETS, ids=RULE_IDS)
def test_lockstep_play_vs_oracle(api, kw):
    """host-chosen legal actions (the oracle's random rule), dice from the Philox contract on the device: legal masks, status bytes
    and every byte of the state compared as the games go, through eliminations, trade-ins and game ends"""
    n, steps, first = 64, 2900, 700
    env = api.Env6(n, rules=api.default_rules(**kw), first_game_id=first)
    env.reset(SEED)
    games = [po.Oracle6Game(po.default_rules(**kw)) for _ in range(n)]
    for g, o in enumerate(games):
        o.new_game(SEED, first + g, 0)
    ply = np.zeros(n, int)
    for s in range(steps):
        check_masks = s % 7 == 0
        if check_masks:
            valid, _ = env.query()
        act = np.zeros(n, np.uint8)
        exp = np.zeros(n, np.int8)
        for g, o in enumerate(games):
            if check_masks:
                assert int(valid[g]) == o.valid(), (s, g)
            if o.status() != -1:
                act[g], exp[g] = 42, -4
                continue
            a = o.random_action(SEED, first + g, int(ply[g]))
            assert o.move(a, SEED, first + g, int(ply[g])) == 0
            ply[g] += 1
            act[g], exp[g] = a, o.status()
        st = env.step(act)
        assert (st == exp).all(), s
        if s % 50 == 0 or s == steps - 1:
            img = env.export()
            for g, o in enumerate(games):
                assert (img[g] == o.image()).all(), (s, g)
    assert sum(o.status() != -1 for o in games) >= (n // 8 if not kw.get("max_game_rounds") else n)
    env.close()


@pytest.mark.gpu
def test_illegal_actions_and_import(api):
    n = 64
    env = api.Env6(n, first_game_id=3)
    env.reset(SEED)
    env.rollout(400)
    img = env.export()
    valid, _ = env.query()
    for a in (0, 17, 41, 42, 43, 255):
        st = env.step(np.full(n, a, np.uint8))
        legal = np.array([a < 43 and (int(v) >> a) & 1 for v in valid], bool)
        assert (st[~legal] == -3).all() and (st[legal] >= -2).all()
        assert (env.export()[~legal] == img[~legal]).all()
        env.import_images(img)                                # back to the saved position
        assert (env.export() == img).all()
    bad = img.copy(); bad[5, 42 + 7] = 6                       # owner 6 does not exist
    with pytest.raises(api.AzError):
        env.import_images(bad)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(allow_yield=0, max_game_rounds=30)], ids=["default", "no_yield_30_rounds"])
def test_rollout_matches_oracle_including_redeals(api, kw):
    n, steps, first = 128, 3500, 40
    env = api.Env6(n, rules=api.default_rules(**kw), first_game_id=first)
    env.reset(SEED)
    env.rollout(steps // 2)
    env.rollout(steps - steps // 2)
    img = env.export()
    cnt = env.counters()
    games, draws, wins = 0, 0, [0] * 6
    o = po.Oracle6Game(po.default_rules(**kw))
    for g in range(n):
        o.new_game(SEED, first + g, 0)
        for ply in range(steps):
            st = o.status()
            if st != -1:
                games += 1
                if st == -2:
                    draws += 1
                else:
                    wins[st] += 1
                o.new_game(SEED, first + g, ply)
            assert o.move(o.random_action(SEED, first + g, ply), SEED, first + g, ply) == 0
        assert (img[g] == o.image()).all(), g
    assert cnt["steps"] == n * steps and (cnt["games"], cnt["draws"], cnt["wins"]) == (games, draws, wins)
    assert games > n // 2
    env.close()


@pytest.mark.gpu
def test_full_size_properties(api):
    """configs[3]'s game count: 16384 games; determinism, shard invariance, a sampled oracle replay"""
    n, steps = 16384, 1200
    env = api.Env6(n)
    env.reset(SEED)
    env.rollout(steps)
    img = env.export()
    cnt = env.counters()
    assert cnt["steps"] == n * steps and cnt["games"] == cnt["draws"] + sum(cnt["wins"])
    assert ((img[:, :42] >= 1) & (img[:, :42] <= 32)).all() and (img[:, 42:84] < 6).all()
    half = api.Env6(n // 2, first_game_id=n // 2)
    half.reset(SEED)
    half.rollout(steps)
    assert (half.export() == img[n // 2:]).all()
    o = po.Oracle6Game()
    for g in np.random.default_rng(3).choice(n, 32, replace=False):
        g = int(g)
        o.new_game(SEED, g, 0)
        for ply in range(steps):
            if o.status() != -1:
                o.new_game(SEED, g, ply)
            assert o.move(o.random_action(SEED, g, ply), SEED, g, ply) == 0
        assert (img[g] == o.image()).all(), g
    env.close(); half.close()
