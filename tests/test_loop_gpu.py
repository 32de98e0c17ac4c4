"""GPU suite: the pieces AlphaZeroTrainer::train strings together (alphazero_trainer.cpp:12-35, 147-166) run as one loop on the device:
self-play with sample emission -> AlphaZeroNN::train on the collected records -> hand-off to the generating copy -> comparison match
new vs old -> benchmark against RandomPlayer.  Every piece has its own parity test; this one checks that they compose: a searcher
sees the weights a training call has just produced, the trained network fits its samples better than the untrained one, and the
match bookkeeping stays consistent."""
import numpy as np
import pytest

from oracle import nn_oracle as no

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    return a


def unpack(api, recs):
    """265-byte records -> (x [n,7,6,13], pi [n,43], z [n]) through the network's own input encoding of the NNInputData image"""
    n = len(recs)
    z = recs[:, 89:93].copy().view(np.float32).reshape(n)
    pi = recs[:, 93:].copy().view(np.float32).reshape(n, 43)
    return pi, z


def test_generate_train_compare_benchmark(api):
    blocks, games, sims = 2, 96, 8
    rules = api.default_rules(mcts_simulations=sims, threads_per_mcts=1)
    train_net, gen_net = api.Net(blocks=blocks, seed=11), api.Net(blocks=blocks, seed=12)
    gen_net.copy_state_from(train_net)                       # both groups start from the same checkpoint (executeTrain)
    env = api.Env(games, rules=rules)
    env.reset(SEED)
    gen = api.Mcts(env, net=gen_net, evaluator=api.EVAL_NN, precision=api.BF16)
    gen.record(capacity_samples=200000, max_moves_per_game=1024)
    # ---- generateTrainData: self-play until enough finished games
    recs = np.zeros((0, 265), np.uint8)
    for _ in range(12):
        gen.selfplay(150)
        r, dropped = gen.samples()
        assert dropped == 0
        recs = np.concatenate([recs, r])
        if len(recs) >= 4096:
            break
    assert len(recs) >= 4096 and gen.counters()["errors"] == 0
    pi, z = unpack(api, recs)
    assert np.allclose(pi.sum(1), 1.0, atol=1e-4) and set(np.unique(z)) <= {-1.0, 0.0, 1.0}
    # ---- trainGroup->train(data, epochs): the tensor-core training mode, 6 epochs of whole batches of 512
    train_net.train_precision(api.BF16)
    lp, lv = train_net.train(recs, epochs=6, batch_size=512, seed=1)
    assert np.isfinite(lp).all() and np.isfinite(lv).all()
    assert lp[-1] < lp[0] and lv[-1] < lv[0], (lp, lv)        # it learns its own samples
    # the fp32 parity mode continues from the same optimizer state and agrees on the next epoch's losses
    probe = api.Net(blocks=blocks, seed=13)
    probe.copy_state_from(train_net)
    l16 = train_net.train(recs[:2048], epochs=1, batch_size=512, seed=9)
    l32 = probe.train(recs[:2048], epochs=1, batch_size=512, seed=9)
    assert abs(l16[0][0] - l32[0][0]) < 2e-2 * abs(l32[0][0]) and abs(l16[1][0] - l32[1][0]) < 5e-2 * max(abs(l32[1][0]), 0.05), (l16, l32)
    # ---- a searcher built BEFORE the training call must see the new weights (finalize on demand)
    env2 = api.Env(8, rules=rules, first_game_id=500)
    env2.reset(SEED)
    mc_new = api.Mcts(env2, net=train_net, evaluator=api.EVAL_NN, precision=api.FP32)
    x = env2.encode()
    _, v_before = train_net.forward(x, api.FP32)
    train_net.train(recs[:1024], epochs=1, batch_size=512, seed=2)
    _, v_after = train_net.forward(x, api.FP32)
    assert np.abs(v_after - v_before).max() > 0
    res = mc_new.search(pick_mode=api.PICK_ARGMAX, apply_move=False)
    root_value = mc_new.root_stats()["value"]
    assert np.allclose(root_value, v_after, atol=1e-5), "the search still evaluates with the weights from before the training call"
    # ---- updateIfImprovement: comparison match new vs old on the device, both collecting samples
    mc_old = api.Mcts(env2, net=gen_net, evaluator=api.EVAL_NN, precision=api.FP32)
    arena = api.Arena(mc_new, mirror_games=True, opponent_mcts=mc_old)
    r = arena.play(16, SEED + 1)
    assert r["errors"] == 0 and r["count"] == 16 and r["draw"] + r["win"][0] + r["win"][1] == 16
    assert r["az_moves"] > 0 and r["opponent_turns"] > 0
    # ---- generateGroup->loadCheckpoint(best): hand-off, then the benchmark against RandomPlayer
    gen_net.copy_state_from(train_net)
    pg, vg = gen_net.forward(x, api.FP32)
    pt, vt = train_net.forward(x, api.FP32)
    assert (pg == pt).all() and (vg == vt).all()
    bench = api.Arena(mc_old, api.OPPONENT_RANDOM, mirror_games=True)
    rb = bench.play(16, SEED + 2)
    assert rb["errors"] == 0 and rb["count"] == 16
    print("loop: %d samples, policy loss %.3f -> %.3f, value loss %.3f -> %.3f; new vs old %s draw %d; vs random %s"
          % (len(recs), lp[0], lp[-1], lv[0], lv[-1], r["win"], r["draw"], rb["win"]))
    for h in (bench, arena, mc_old, mc_new, gen, env2, env, probe, train_net, gen_net):
        h.close()
