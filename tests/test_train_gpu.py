"""GPU suite: the training step (az_nn_train_step / az_nn_train) and checkpoint exchange (az_nn_save/load_checkpoint) through the
C ABI against oracle/nn_oracle.py's Trainer (PyTorch autograd in fp64 of python/src/build_graph.py:54-106 with the constants of the
shipped GraphDef) and oracle/ckpt_oracle.py.  PARITY UNPINNED at the TensorFlow boundary (see both oracle headers); tolerances are
for fp32 kernels against an fp64 restatement."""
import json
import os

import numpy as np
import pytest

from oracle import ckpt_oracle as co
from oracle import nn_oracle as no
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

GRAD_TOL = 2e-4        # max |g - g_ref| <= GRAD_TOL * max |g_ref| per variable (fp32 accumulation over up to 16 x 42 x 256 x 9 terms)
LOSS_TOL = 2e-5


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    return a


def positions(n, seed=0xC0FFEE):
    """(x, target pi, target z) from seeded random play: real inputs, random policy targets on the legal moves, random outcomes"""
    xs, tps, o, g, rng = [], [], po.OracleGame(), 0, np.random.default_rng(seed)
    while len(xs) < n:
        o.new_game(seed, g, 0)
        ply = 0
        while o.status() == -1 and len(xs) < n:
            if ply % 7 == 3:
                xs.append(o.encode())
                valid = o.valid()
                t = np.array([rng.random() if (valid >> i) & 1 else 0.0 for i in range(43)])
                tps.append(t / t.sum())
            o.move(o.random_action(seed, g, ply), seed, g, ply)
            ply += 1
        g += 1
    tv = rng.choice([-1.0, 0.0, 1.0], n).astype(np.float32)
    return np.array(xs, np.float32).reshape(n, 7, 6, 13), np.array(tps, np.float32), tv


def perturbed_net(api, blocks, seed):
    net = api.Net(blocks=blocks, seed=seed)
    rng = np.random.default_rng(seed)
    for name, shape in net.variables():       # make BN parameters non-trivial
        if name.endswith("/gamma"):
            net.load(name, rng.uniform(0.8, 1.2, shape))
        elif name.endswith("/beta") or name.endswith("/bias"):
            net.load(name, rng.uniform(-0.1, 0.1, shape))
    return net


@pytest.mark.parametrize("blocks,n", [(2, 16), (1, 5), (3, 70)])
def test_one_step_losses_gradients_moving_statistics(api, blocks, n):
    net = perturbed_net(api, blocks, 77)
    ref = no.Trainer(net.weights(), blocks)
    x, tp, tv = positions(n)
    lp, lv = net.train_step(x, tp, tv)
    rp, rv = ref.step(x, tp, tv)
    assert abs(lp - rp) <= LOSS_TOL * max(1.0, abs(rp)) and abs(lv - rv) <= LOSS_TOL * max(1.0, abs(rv)), (lp, rp, lv, rv)
    shapes = dict(net.variables())
    worst = 0.0
    for name in no.trainable_names(blocks):
        g, gr = net.grad(name, shapes[name]), ref.grads[name].numpy()
        scale = np.abs(gr).max()
        err = np.abs(g - gr).max() / max(scale, 1e-12)
        worst = max(worst, err)
        assert err <= GRAD_TOL, (name, err, scale)
    w, wr = net.weights(), ref.weights()
    for name in shapes:
        if "moving_" in name:
            assert np.abs(w[name] - wr[name]).max() <= 1e-5 * max(1.0, np.abs(wr[name]).max()), name
    # Adam: the slots are linear / quadratic in the gradient; after the first step every weight moved by ~lr * sign(g)
    for name in no.trainable_names(blocks):
        m, v = net.optimizer_slot(name, 0, shapes[name]), net.optimizer_slot(name, 1, shapes[name])
        assert np.abs(m - ref.m[name].numpy()).max() <= GRAD_TOL * 0.1 * max(np.abs(ref.grads[name].numpy()).max(), 1e-12) + 1e-12
        assert np.abs(v - ref.v[name].numpy()).max() <= 3 * GRAD_TOL * 0.001 * max((ref.grads[name].numpy() ** 2).max(), 1e-24) + 1e-20
        clear = np.abs(ref.grads[name].numpy()) > 1e-6          # sign of a gradient at fp32 noise level is not defined
        assert np.abs(w[name] - wr[name])[clear].max(initial=0.0) <= 2e-5, name
    b1, b2, steps = net.optimizer_powers()
    assert steps == 1 and abs(b1 - 0.81) < 1e-6 and abs(b2 - 0.998001) < 1e-6
    print("blocks %d n %d: worst relative gradient error %.2e" % (blocks, n, worst))
    net.close()


def _layer_names(blocks):
    return ["conv_bn"] + ["bn%d%s_branch%s" % (i, chr(97 + i), br) for i in range(blocks) for br in ("2a", "2b")]


def _rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("blocks,n", [(1, 5), (1, 64), (2, 16), (3, 70), (1, 131)])
def test_tensor_core_training_mode_against_the_oracle(api, blocks, n):
    """train_precision(BF16): the 3x3 convolutions' three contractions (forward, data gradient, weight gradient) as bf16 tcgen05
    GEMMs with fp32 accumulation; BatchNorm, heads, losses and Adam stay fp32.  Two checkers:
      * Trainer(bf16=True) rounds exactly those operands to bf16 (nn_oracle._ConvBf16) and is otherwise exact.  Rounding is a
        discontinuous map, so an fp32-level difference e between two runs flips a fraction e / ulp of the next layer's operands and
        comes out as sqrt(e * ulp): agreement is 1e-7 at the stem, <= 1e-4 at the first tower layer, and decays towards the bf16 ulp
        (4e-3) with depth.  The stated bounds: stem output 1e-6, first tower layer 1e-4 (relative L2); for a one-block tower every
        gradient within 2e-2 relative L2 of the emulation (measured <= 9e-3; the exact fp64 oracle is 3e-2 .. 1e-1 away there);
      * the exact fp64 Trainer: losses within 1e-2 relative, every gradient's cosine >= 0.98 (measured values are printed)."""
    net = perturbed_net(api, blocks, 77)
    net.train_precision(api.BF16)
    w0 = net.weights()
    exact, emu = no.Trainer(w0, blocks), no.Trainer(w0, blocks, bf16=True)
    x, tp, tv = positions(n)
    lp, lv = net.train_step(x, tp, tv)
    rp, rv = exact.step(x, tp, tv)
    no.TRACE = {}
    try:
        ep, ev = emu.step(x, tp, tv)
        trace = no.TRACE
    finally:
        no.TRACE = None
    assert abs(lp - rp) <= 1e-2 * max(1.0, abs(rp)) and abs(lv - rv) <= 1e-2 * max(1.0, abs(rv)), (lp, rp, lv, rv)
    assert abs(lp - ep) <= 2e-3 * max(1.0, abs(ep)) and abs(lv - ev) <= 2e-3 * max(1.0, abs(ev)), (lp, ep, lv, ev)
    names = _layer_names(blocks)
    z0, z1 = _rel(net.layer(0, 0, n), trace[names[0]]), _rel(net.layer(1, 0, n), trace[names[1]])
    shapes = dict(net.variables())
    worst_emu, worst_cos, bad = ("", 0.0), ("", 1.0), []
    for name in no.trainable_names(blocks):
        g = net.grad(name, shapes[name]).ravel().astype(np.float64)
        ge, gx = emu.grads[name].numpy().ravel(), exact.grads[name].numpy().ravel()
        r = _rel(g, ge)
        cos = float(g @ gx / max(np.linalg.norm(g) * np.linalg.norm(gx), 1e-30))
        if r > worst_emu[1]:
            worst_emu = (name, r)
        if cos < worst_cos[1]:
            worst_cos = (name, cos)
        if cos < 0.98 and ge.size > 64:                        # the few-element BatchNorm vectors are covered by the emulation bound
            bad.append((name, "cosine", cos))
        if blocks == 1 and r > 2e-2:
            bad.append((name, "rel l2 vs emulation", r))
    w, wr = net.weights(), emu.weights()
    for name in shapes:
        if "moving_" in name:
            assert np.abs(w[name] - wr[name]).max() <= 2e-3 * max(1.0, np.abs(wr[name]).max()), name
    print("bf16 tcgen05 training, blocks %d n %d: losses %.5f / %.5f (emulation %.5f / %.5f, exact %.5f / %.5f); layer z rel l2 vs emulation "
          "%.1e / %.1e; worst gradient rel l2 vs emulation %.2e (%s), worst cosine vs exact %.5f (%s)"
          % (blocks, n, lp, lv, ep, ev, rp, rv, z0, z1, worst_emu[1], worst_emu[0], worst_cos[1], worst_cos[0]))
    net.close()
    assert z0 <= 1e-6 and z1 <= 1e-4, (z0, z1)
    assert not bad, bad


def test_several_steps_track_the_oracle_and_inference_uses_the_trained_weights(api):
    blocks, n = 2, 24
    net = perturbed_net(api, blocks, 5)
    ref = no.Trainer(net.weights(), blocks)
    x, tp, tv = positions(n, seed=99)
    first = None
    for step in range(4):
        lp, lv = net.train_step(x, tp, tv)
        rp, rv = ref.step(x, tp, tv)
        assert abs(lp - rp) <= 2e-3 * max(1.0, abs(rp)) and abs(lv - rv) <= 2e-3 * max(1.0, abs(rv)), (step, lp, rp, lv, rv)
        first = first if first is not None else (lp, lv)
    assert lp < first[0] and lv < first[1]      # four steps on the same batch: both losses went down
    w, wr = net.weights(), ref.weights()
    for name in w:
        # Adam divides by sqrt(v): an element whose gradient is at fp32 noise level (inputs that are constant over the batch) moves by
        # ~lr per step in a direction the noise picks, so a small share of elements may differ by up to steps * 2 * lr
        # ... and everything downstream of such an element drifts a little with it: the bulk must agree closely, nothing may run away
        d = np.abs(w[name] - wr[name])
        assert d.max() <= 8.1e-3 and (d.size < 256 or np.median(d) <= 2e-5), (name, np.median(d), d.max())
    # the forward paths pick the trained weights up (fp32 path against the restatement, bf16 path against fp32)
    p, v = net.forward(x, api.FP32)
    pr, vr = no.forward(w, x, blocks)
    assert np.abs(p - pr).max() <= 1e-5 and np.abs(v - vr).max() <= 1e-5
    pb, vb = net.forward(x, api.BF16)
    assert np.abs(pb - p).max() <= 2e-2 and np.abs(vb - v).max() <= 5e-2
    net.close()


def test_checkpoint_round_trip_and_inventory(api, golden_dir, tmp_path):
    """az_nn_save_checkpoint writes exactly the tensors the reference graph's Saver writes (names, order, shapes from the shipped
    GraphDef), readable by the independent Python reader; az_nn_load_checkpoint restores weights, Adam slots and beta powers"""
    net = perturbed_net(api, 5, 11)
    x, tp, tv = positions(8)
    net.train_step(x, tp, tv); net.train_step(x, tp, tv)
    prefix = str(tmp_path / "az_train")
    net.save_checkpoint(prefix)
    spec = json.load(open(os.path.join(golden_dir, "ckpt_tensors_V2_5.json")))["tensors"]
    ck = api.Checkpoint(prefix)
    assert [(t[0], list(t[2])) for t in ck.tensors()] == [(n, s) for n, s in spec]
    ck.close()
    back = co.read_bundle(prefix)
    shapes = dict(net.variables())
    w = net.weights()
    for name in shapes:
        assert (back[name].view(np.uint32) == w[name].view(np.uint32)).all(), name
        if "moving_" not in name:
            assert (back[name + "/optimize"] == net.optimizer_slot(name, 0, shapes[name])).all()
            assert (back[name + "/optimize_1"] == net.optimizer_slot(name, 1, shapes[name])).all()
    b1, b2, _ = net.optimizer_powers()
    assert back["beta1_power"].shape == () and float(back["beta1_power"]) == np.float32(b1) and float(back["beta2_power"]) == np.float32(b2)
    other = api.Net(blocks=5, seed=1)
    other.load_checkpoint(prefix)
    w2 = other.weights()
    for name in shapes:
        assert (w2[name].view(np.uint32) == w[name].view(np.uint32)).all(), name
    assert other.optimizer_powers()[:2] == (b1, b2)
    # both continue identically: the optimizer state came along
    a = net.train_step(x, tp, tv); b = other.train_step(x, tp, tv)
    assert a == b
    wa, wb = net.weights(), other.weights()
    for name in shapes:
        assert (wa[name].view(np.uint32) == wb[name].view(np.uint32)).all(), name
    # a checkpoint of another architecture is refused, like restore_all would
    small = api.Net(blocks=2, seed=1)
    small.save_checkpoint(str(tmp_path / "small"))
    with pytest.raises(api.AzError):
        other.load_checkpoint(str(tmp_path / "small"))
    with pytest.raises(api.AzError):
        other.load_checkpoint(str(tmp_path / "does_not_exist"))
    net.close(); other.close(); small.close()


def test_train_on_sample_records_equals_manual_batches(api):
    """az_nn_train (AlphaZeroNN::train, alphazero_nn.cpp:351-410): records -> input planes / targets on the device, shuffled whole
    batches per epoch == az_nn_train_step on the same batches built on the host from the oracle's encoding"""
    rules = po.default_rules()
    o, rng = po.OracleGame(rules), np.random.default_rng(3)
    recs, xs, tps, tvs = [], [], [], []
    g = 0
    while len(recs) < 150:
        o.new_game(0xABCD, g, 0)
        ply = 0
        while o.status() == -1 and len(recs) < 150:
            if ply % 5 == 0:
                valid = o.valid()
                t = np.array([rng.random() if (valid >> i) & 1 else 0.0 for i in range(43)], np.float32)
                t /= t.sum()
                status = int(rng.choice([0, 1, -2]))
                recs.append(o.sample_record(t, status)); xs.append(o.encode()); tps.append(t)
                tvs.append(0.0 if status == -2 else (1.0 if status == o.s.cur else -1.0))
            o.move(o.random_action(0xABCD, g, ply), 0xABCD, g, ply)
            ply += 1
        g += 1
    recs = np.stack(recs); xs = np.array(xs, np.float32); tps = np.array(tps, np.float32); tvs = np.array(tvs, np.float32)
    epochs, batch, seed = 2, 32, 42
    a, b = perturbed_net(api, 1, 8), perturbed_net(api, 1, 8)
    lp, lv = a.train(recs, epochs, batch, seed)
    order, st, M = np.arange(len(recs)), seed, (1 << 64) - 1
    for e in range(epochs):
        for i in range(len(recs) - 1, 0, -1):                       # the library's Fisher-Yates on splitmix64
            st = (st + 0x9E3779B97F4A7C15) & M
            z = st; z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M; z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M; z ^= z >> 31
            j = z % (i + 1)
            order[i], order[j] = order[j], order[i]
        losses = []
        for k in range(len(recs) // batch):
            idx = order[k * batch:(k + 1) * batch]
            losses.append(b.train_step(xs[idx], tps[idx], tvs[idx]))
        assert abs(np.mean([l[0] for l in losses]) - lp[e]) < 1e-5 and abs(np.mean([l[1] for l in losses]) - lv[e]) < 1e-5
    wa, wb = a.weights(), b.weights()
    for name in wa:
        assert (wa[name].view(np.uint32) == wb[name].view(np.uint32)).all(), name
    assert lp[1] < lp[0]
    a.close(); b.close()


def test_replica_hand_off_copies_variables_and_optimizer_state(api):
    """az_nn_copy_state = AlphaZeroNNGroup::train's save-temp-checkpoint + reload into the other copies (alphazero_gpu_cluster.cpp:221-231)
    without the file: after the copy both networks predict the same, and one more identical step leaves weights and Adam slots equal bit
    for bit (so the slots and beta powers travelled too).  With two GPUs the copy crosses devices."""
    blocks, n = 2, 24
    x, tp, tv = positions(n)
    a = perturbed_net(api, blocks, 3)
    for _ in range(2):
        a.train_step(x, tp, tv)
    devices = [0] + ([1] if api.lib().az_device_count() >= 2 else [])
    for dev in devices:
        b = api.Net(blocks=blocks, seed=99, device=dev)
        b.train_step(x[:8], tp[:8], tv[:8])                      # b has its own (different) optimizer history to be overwritten
        b.copy_state_from(a)
        pa, va = a.forward(x, api.FP32)
        pb, vb = b.forward(x, api.FP32)
        assert (pa == pb).all() and (va == vb).all()
        a2 = api.Net(blocks=blocks, seed=5)
        a2.copy_state_from(a)                                    # a stays untouched for the next device: step a copy of it
        la, lb = a2.train_step(x, tp, tv), b.train_step(x, tp, tv)
        assert la == lb
        wa, wb = a2.weights(), b.weights()
        shapes = dict(a2.variables())
        for name in shapes:
            assert (wa[name] == wb[name]).all(), name
        for name in no.trainable_names(blocks):
            for which in (0, 1):
                assert (a2.optimizer_slot(name, which, shapes[name]) == b.optimizer_slot(name, which, shapes[name])).all(), (name, which)
        assert a2.optimizer_powers() == b.optimizer_powers()
        a2.close(); b.close()
    fresh, target = api.Net(blocks=blocks, seed=8), api.Net(blocks=blocks, seed=9)
    target.train_step(x[:8], tp[:8], tv[:8])
    target.copy_state_from(fresh)                                # a source that never trained resets the target's optimizer
    assert target.optimizer_powers()[2] == 0 and (target.weights()["conv/kernel"] == fresh.weights()["conv/kernel"]).all()
    other = api.Net(blocks=3, seed=1)
    with pytest.raises(api.AzError):
        other.copy_state_from(a)
    a.close(); fresh.close(); target.close(); other.close()


def test_bad_arguments(api):
    net = api.Net(blocks=1, seed=3)
    x, tp, tv = positions(4)
    with pytest.raises(api.AzError):
        net.train_step(x[:1], tp[:1], tv[:1])             # batch statistics need >= 2 samples
    with pytest.raises(api.AzError):
        net.grad("conv/kernel", (3, 3, 13, 256))          # no step has run yet
    with pytest.raises(api.AzError):
        net.train(np.zeros((3, 265), np.uint8), 1, 8, 0)  # fewer samples than one batch
    net.close()
