"""CPU suite: self-consistency of oracle/nn_oracle.py's training restatement (the checker of the GPU training tests).
  * the bf16-operand emulation (Trainer(bf16=True), nn_oracle._ConvBf16) changes nothing when every operand of every contraction is
    already a bf16 number, and its three contractions equal plain autograd on rounded operands;
  * the analytic gradients of the exact Trainer agree with central finite differences of its own loss (fp64)."""
import numpy as np
import torch

from oracle import nn_oracle as no


def _weights(blocks, seed):
    rng = np.random.default_rng(seed)
    shapes = {"conv/kernel": (3, 3, 13, 256), "pi/kernel": (1, 1, 256, 2), "dense/kernel": (84, 43), "dense/bias": (43,), "v/kernel": (1, 1, 256, 1),
              "dense_1/kernel": (42, 256), "dense_1/bias": (256,), "dense_2/kernel": (256, 1), "dense_2/bias": (1,)}
    for i in range(blocks):
        sfx = "%d%s" % (i, chr(97 + i))
        for br in ("2a", "2b"):
            shapes["res%s_branch%s/kernel" % (sfx, br)] = (3, 3, 256, 256)
    w = {}
    for name in no.variable_names(blocks):
        if name in shapes:
            w[name] = (rng.standard_normal(shapes[name]) * 0.05).astype(np.float32)
        else:
            c = 7 if name.startswith("conv_bn") else (2 if name.startswith("bn_pi") else (1 if name.startswith("bn_v") else 256))
            base = {"gamma": 1.0, "beta": 0.0, "moving_mean": 0.0, "moving_variance": 1.0}[name.split("/")[1]]
            w[name] = (base + rng.uniform(-0.1, 0.1, c)).astype(np.float32)
    return w


def _batch(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 7, 6, 13)).astype(np.float32)
    tp = rng.random((n, 43)).astype(np.float32)
    tp /= tp.sum(1, keepdims=True)
    return x, tp, rng.choice([-1.0, 0.0, 1.0], n).astype(np.float32)


def test_bf16_convolution_equals_autograd_on_rounded_operands():
    torch.manual_seed(0)
    x = torch.randn(3, 13, 7, 6, dtype=torch.float64, requires_grad=True)
    k = torch.randn(32, 13, 3, 3, dtype=torch.float64, requires_grad=True)
    g = torch.randn(3, 32, 7, 6, dtype=torch.float64)
    y = no._ConvBf16.apply(x, k)
    gx, gk = torch.autograd.grad(y, (x, k), g)
    xr, kr, gr = no._rb(x.detach()).requires_grad_(True), no._rb(k.detach()).requires_grad_(True), no._rb(g)
    yr = torch.nn.functional.conv2d(xr, kr, padding=1)
    assert torch.equal(y, yr)
    gxr, gkr = torch.autograd.grad(yr, (xr, kr), gr)
    assert torch.allclose(gx, gxr, rtol=0, atol=1e-12) and torch.allclose(gk, gkr, rtol=0, atol=1e-12)
    # rounding is idempotent and really lands on bf16 numbers (low 16 bits of the fp32 pattern clear)
    r = no._rb(x.detach()).to(torch.float32).numpy().view(np.uint32)
    assert (r & 0xffff == 0).all() and torch.equal(no._rb(no._rb(x.detach())), no._rb(x.detach()))


def test_exact_trainer_gradients_match_finite_differences():
    blocks, n = 1, 4
    w0 = _weights(blocks, 3)
    x, tp, tv = _batch(n, 4)
    tr = no.Trainer(w0, blocks)
    tr.step(x, tp, tv)

    def loss(wd):
        t = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in wd.items()}
        lp, lv, l2 = no.train_losses(t, torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(tp, dtype=torch.float64).reshape(-1, 43),
                                     torch.as_tensor(tv, dtype=torch.float64), blocks, {})
        return float(lp + lv + l2)

    rng = np.random.default_rng(9)
    for name in ("conv/kernel", "res0a_branch2b/kernel", "bn0a_branch2a/gamma", "conv_bn/beta", "dense_1/kernel", "pi/kernel", "dense_2/bias"):
        g = tr.grads[name].numpy()
        for _ in range(3):
            idx = tuple(int(rng.integers(0, d)) for d in g.shape)
            wp = {k: np.asarray(v, np.float64).copy() for k, v in w0.items()}
            wm = {k: np.asarray(v, np.float64).copy() for k, v in w0.items()}
            h = 1e-5
            wp[name][idx] += h
            wm[name][idx] -= h
            fd = (loss(wp) - loss(wm)) / (2 * h)
            assert abs(fd - g[idx]) <= 1e-6 + 1e-4 * abs(g[idx]), (name, idx, fd, g[idx])


def test_bf16_trainer_tracks_the_exact_one():
    """same step, operands rounded: losses within 1e-2, every large gradient's direction preserved (what the GPU test demands of the device)"""
    blocks, n = 1, 48
    w0 = _weights(blocks, 5)
    x, tp, tv = _batch(n, 6)
    a, b = no.Trainer(w0, blocks), no.Trainer(w0, blocks, bf16=True)
    la, lb = a.step(x, tp, tv), b.step(x, tp, tv)
    assert abs(la[0] - lb[0]) < 1e-2 * abs(la[0]) and abs(la[1] - lb[1]) < 1e-2 * max(abs(la[1]), 1.0)
    for name in no.trainable_names(blocks):
        ga, gb = a.grads[name].numpy().ravel(), b.grads[name].numpy().ravel()
        if ga.size > 64:
            assert float(ga @ gb / (np.linalg.norm(ga) * np.linalg.norm(gb))) > 0.97, name
    assert not no.BF16_CONTRACTIONS and no.TRACE is None       # the switches are restored
