"""TEST INFRASTRUCTURE — CPU restatement (PyTorch fp32 / fp64) of the reference's policy/value
network, /root/reference/python/src/build_graph.py:37-90, cross-checked structurally against the
shipped GraphDef python/model/model_txt_V2_5.pb (variable names and shapes, SAME padding, NHWC,
epsilon 0.001, the stem's BatchNorm running over the board-ROW axis: build_graph.py:68 passes
axis=1 on an NHWC tensor, so conv_bn/* have 7 elements).

PARITY: the arithmetic the reference executes lives in TensorFlow (un-vendored, un-pinned, not installable here) and the reference
ships neither a checkpoint nor an example output.  What CAN be pinned is pinned: oracle/graphdef_oracle.py executes the inference
slice of the reference's own shipped GraphDef (tests/golden/graph_V2_5_inference.json, extracted from python/model/model_txt_V2_5.pb)
op by op, and tests/test_graphdef_cpu.py holds this restatement to it (< 1e-9 in float64; golden vectors in
tests/golden/graph_forward_V2_5.npz).  PARITY UNPINNED remains true for the TensorFlow kernels themselves (Conv2D, FusedBatchNormV3, ...:
restated from their published definitions).  Training step: the graph's TRAINING forward (input_training = true: both losses, the
minimised total add_6, every moving-average update) is evaluated from the reference's own nodes (tests/golden/graph_V2_5_training.json)
and equals train_losses / Trainer's update to < 1e-11, and the 45 ResourceApplyAdam ops' wiring and constants are read from the file;
the gradient sub-graph itself is not interpreted (it is TensorFlow's derivative of that same forward; Trainer uses autograd).
This module is the checker for the <= 1e-5 fp32 agreement north_star asks for; only tests/ and __graft_entry__.smoke() may import it.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = float(np.float32(0.001))           # the GraphDef's attr is a float32: 0.0010000000474974513


def variable_names(blocks):
    names = ["conv/kernel"] + ["conv_bn/" + s for s in ("gamma", "beta", "moving_mean", "moving_variance")]
    for i in range(blocks):
        sfx = "%d%s" % (i, chr(ord("a") + i))
        for br in ("2a", "2b"):
            names.append("res%s_branch%s/kernel" % (sfx, br))
            names += ["bn%s_branch%s/%s" % (sfx, br, s) for s in ("gamma", "beta", "moving_mean", "moving_variance")]
    names += ["pi/kernel"] + ["bn_pi/" + s for s in ("gamma", "beta", "moving_mean", "moving_variance")]
    names += ["dense/kernel", "dense/bias", "v/kernel"]
    names += ["bn_v/" + s for s in ("gamma", "beta", "moving_mean", "moving_variance")]
    names += ["dense_1/kernel", "dense_1/bias", "dense_2/kernel", "dense_2/bias"]
    return names


def _bn(x, w, prefix, axis):
    """inference FusedBatchNormV3: (x - mean) * (gamma * rsqrt(var + eps)) + beta along `axis` of an NCHW tensor"""
    g, b, m, v = (w[prefix + "/" + s] for s in ("gamma", "beta", "moving_mean", "moving_variance"))
    shape = [1, 1, 1, 1]
    shape[axis] = -1
    scale = g * torch.rsqrt(v + EPS)
    return (x - m.reshape(shape)) * scale.reshape(shape) + b.reshape(shape)


def _rb(t):
    """round to bf16 (round-to-nearest-even from fp32, what the device does to a GEMM operand) and back"""
    return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)


class _ConvBf16(torch.autograd.Function):
    """the 3x3 convolution of the tensor-core training mode (az_nn_train_precision(AZ_NN_BF16)): each of its three contractions
    (forward, data gradient, weight gradient) takes both operands rounded to bf16 and accumulates exactly (fp32 on the device)"""

    @staticmethod
    def forward(ctx, x, k):
        ctx.save_for_backward(x, k)
        return F.conv2d(_rb(x), _rb(k), padding=k.shape[2] // 2)

    @staticmethod
    def backward(ctx, g):
        x, k = ctx.saved_tensors
        pad = k.shape[2] // 2
        gx = torch.nn.grad.conv2d_input(x.shape, _rb(k), _rb(g), padding=pad)
        gk = torch.nn.grad.conv2d_weight(_rb(x), k.shape, _rb(g), padding=pad)
        return gx, gk


TRACE = None                   # a dict here collects every BatchNorm's input during train_losses (per-layer forward checks)
BF16_CONTRACTIONS = False      # set by Trainer(bf16=True) around its step: 3x3 convolutions through _ConvBf16


def _conv(x, k):
    """k: HWIO -> conv2d with SAME padding, stride 1, no bias; x NCHW"""
    kh = k.shape[0]
    kk = k.permute(3, 2, 0, 1).contiguous()
    if BF16_CONTRACTIONS and kh == 3:
        return _ConvBf16.apply(x, kk)
    return F.conv2d(x, kk, padding=kh // 2)


def forward(weights, x, blocks, dtype=torch.float32):
    w = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in weights.items()}
    t = torch.as_tensor(np.asarray(x), dtype=dtype).reshape(-1, 7, 6, 13).permute(0, 3, 1, 2)   # NCHW: [n,13,7,6]
    t = _conv(t, w["conv/kernel"])
    t = torch.relu(_bn(t, w, "conv_bn", axis=2))            # axis=1 of NHWC == board row y == dim 2 of NCHW
    for i in range(blocks):
        sfx = "%d%s" % (i, chr(ord("a") + i))
        r = torch.relu(_bn(_conv(t, w["res%s_branch2a/kernel" % sfx]), w, "bn%s_branch2a" % sfx, axis=1))
        r = _bn(_conv(r, w["res%s_branch2b/kernel" % sfx]), w, "bn%s_branch2b" % sfx, axis=1)
        t = torch.relu(r + t)
    p = torch.relu(_bn(_conv(t, w["pi/kernel"]), w, "bn_pi", axis=1))
    p = p.permute(0, 2, 3, 1).reshape(-1, 84)               # tf.layers.flatten of NHWC [n,7,6,2]
    logits = p @ w["dense/kernel"] + w["dense/bias"]
    policy = torch.softmax(logits, dim=1)
    v = torch.relu(_bn(_conv(t, w["v/kernel"]), w, "bn_v", axis=1))
    v = v.permute(0, 2, 3, 1).reshape(-1, 42)
    v = torch.relu(v @ w["dense_1/kernel"] + w["dense_1/bias"])
    v = torch.tanh(v @ w["dense_2/kernel"] + w["dense_2/bias"]).reshape(-1)
    return policy.numpy(), v.numpy()


def flops_per_position(blocks):
    """SURVEY.md §8d conventional count (padded taps included)"""
    return 2 * 42 * (9 * 13 * 256 + 2 * blocks * 9 * 256 * 256 + 256 * 2 + 256 * 1) + 2 * (84 * 43 + 42 * 256 + 256)


# --------------------------------------------------------------------------- training step
# What session->Run(TF_OP_OPTIMIZE) does for one batch (neural_network/alphazero_nn.cpp:389-390) according to
# python/src/build_graph.py:92-106 and the constants recorded in the shipped GraphDef (python/model/model_txt_V2_5.pb):
#   * every BatchNorm runs FusedBatchNormV3 with is_training = true: normalise with the BATCH mean and the biased batch
#     variance, epsilon 0.001; the moving statistics move towards the batch mean and the UNBIASED batch variance
#     (N / (N - 1)) with momentum 0.99 (conv_bn_cond_1_true: 0.9900000095, AssignMovingAvg: moving -= (moving - batch) * 0.01);
#     the stem's BN is the NCHW variant over the board-row axis (conv_bn_cond_true: data_format "NCHW")
#   * loss_pi = tf.losses.softmax_cross_entropy (mean over the batch), loss_v = tf.losses.mean_squared_error (mean),
#     l2 = 0.001 * sum(w^2) over the 16 kernels that carry a kernel_regularizer (convs and dense kernels; no bias, no BN)
#   * AdamOptimizer(0.001, beta1 0.9, beta2 0.999, epsilon 1e-8; optimize/{learning_rate,beta1,beta2,epsilon}):
#     lr_t = lr * sqrt(1 - beta2_power) / (1 - beta1_power); m += (g - m)(1 - beta1); v += (g^2 - v)(1 - beta2);
#     var -= lr_t * m / (sqrt(v) + eps); then beta1_power *= beta1, beta2_power *= beta2 (initially beta1, beta2)
# Pinned to the GraphDef's own training-forward nodes and optimizer wiring by tests/test_graphdef_cpu.py; the TensorFlow kernels
# (FusedBatchNormV3 training mode, ResourceApplyAdam) are restated from their published definitions.
# float32 constants of the GraphDef, as float32 values (tests/test_graphdef_cpu.py evaluates the graph's own nodes against these)
BN_MOMENTUM = float(np.float32(0.99))
L2_C = float(np.float32(0.001))
ADAM = dict(lr=float(np.float32(0.001)), beta1=float(np.float32(0.9)), beta2=float(np.float32(0.999)), eps=float(np.float32(1e-8)))


def trainable_names(blocks):
    return [n for n in variable_names(blocks) if not n.endswith(("moving_mean", "moving_variance"))]


def _bn_train(x, w, prefix, axis, batch_stats):
    g, b = w[prefix + "/gamma"], w[prefix + "/beta"]
    dims = [d for d in range(4) if d != axis]
    mean = x.mean(dim=dims)
    var = ((x - mean.reshape([-1 if d == axis else 1 for d in range(4)])) ** 2).mean(dim=dims)
    n = x.numel() // x.shape[axis]
    batch_stats[prefix] = (mean.detach(), (var * n / (n - 1)).detach())
    if TRACE is not None:
        TRACE[prefix] = x.detach().permute(0, 2, 3, 1).numpy().copy()       # the convolution output entering this BatchNorm, NHWC
    shape = [1, 1, 1, 1]
    shape[axis] = -1
    return (x - mean.reshape(shape)) * torch.rsqrt(var.reshape(shape) + EPS) * g.reshape(shape) + b.reshape(shape)


def train_losses(w, x, target_policy, target_value, blocks, batch_stats):
    """training-mode forward; w: dict of torch tensors (leaf tensors requiring grad for the trainable ones)"""
    t = x.reshape(-1, 7, 6, 13).permute(0, 3, 1, 2)
    t = torch.relu(_bn_train(_conv(t, w["conv/kernel"]), w, "conv_bn", 2, batch_stats))
    for i in range(blocks):
        sfx = "%d%s" % (i, chr(ord("a") + i))
        r = torch.relu(_bn_train(_conv(t, w["res%s_branch2a/kernel" % sfx]), w, "bn%s_branch2a" % sfx, 1, batch_stats))
        r = _bn_train(_conv(r, w["res%s_branch2b/kernel" % sfx]), w, "bn%s_branch2b" % sfx, 1, batch_stats)
        t = torch.relu(r + t)
    p = torch.relu(_bn_train(_conv(t, w["pi/kernel"]), w, "bn_pi", 1, batch_stats))
    logits = p.permute(0, 2, 3, 1).reshape(-1, 84) @ w["dense/kernel"] + w["dense/bias"]
    loss_pi = -(target_policy * torch.log_softmax(logits, dim=1)).sum(dim=1).mean()
    v = torch.relu(_bn_train(_conv(t, w["v/kernel"]), w, "bn_v", 1, batch_stats))
    v = torch.relu(v.permute(0, 2, 3, 1).reshape(-1, 42) @ w["dense_1/kernel"] + w["dense_1/bias"])
    v = torch.tanh(v @ w["dense_2/kernel"] + w["dense_2/bias"]).reshape(-1)
    loss_v = ((v - target_value) ** 2).mean()
    l2 = sum((w[n] ** 2).sum() for n in w if n.endswith("/kernel")) * L2_C
    return loss_pi, loss_v, l2


class Trainer:
    """the graph's variables + Adam slots; step() = one TF_OP_OPTIMIZE run"""

    def __init__(self, weights, blocks, dtype=torch.float64, bf16=False):
        self.blocks, self.dtype, self.bf16 = blocks, dtype, bf16
        self.w = {k: torch.as_tensor(np.asarray(v), dtype=dtype).clone() for k, v in weights.items()}
        self.m = {k: torch.zeros_like(self.w[k]) for k in trainable_names(blocks)}
        self.v = {k: torch.zeros_like(self.w[k]) for k in trainable_names(blocks)}
        self.beta1_power, self.beta2_power = ADAM["beta1"], ADAM["beta2"]
        self.grads = {}

    def step(self, x, target_policy, target_value):
        names = trainable_names(self.blocks)
        for k in names:
            self.w[k].requires_grad_(True)
            self.w[k].grad = None
        stats = {}
        x = torch.as_tensor(np.asarray(x), dtype=self.dtype)
        tp = torch.as_tensor(np.asarray(target_policy), dtype=self.dtype).reshape(-1, 43)
        tv = torch.as_tensor(np.asarray(target_value), dtype=self.dtype).reshape(-1)
        global BF16_CONTRACTIONS
        BF16_CONTRACTIONS = self.bf16
        try:
            loss_pi, loss_v, l2 = train_losses(self.w, x, tp, tv, self.blocks, stats)
            (loss_pi + loss_v + l2).backward()
        finally:
            BF16_CONTRACTIONS = False
        lr_t = ADAM["lr"] * np.sqrt(1.0 - self.beta2_power) / (1.0 - self.beta1_power)
        with torch.no_grad():
            for k in names:
                g = self.w[k].grad
                self.grads[k] = g.detach().clone()
                self.m[k] += (g - self.m[k]) * (1.0 - ADAM["beta1"])
                self.v[k] += (g * g - self.v[k]) * (1.0 - ADAM["beta2"])
                self.w[k] -= lr_t * self.m[k] / (torch.sqrt(self.v[k]) + ADAM["eps"])
                self.w[k].requires_grad_(False)
            for prefix, (mean, var_unbiased) in stats.items():
                self.w[prefix + "/moving_mean"] -= (self.w[prefix + "/moving_mean"] - mean) * (1.0 - BN_MOMENTUM)
                self.w[prefix + "/moving_variance"] -= (self.w[prefix + "/moving_variance"] - var_unbiased) * (1.0 - BN_MOMENTUM)
        self.beta1_power *= ADAM["beta1"]
        self.beta2_power *= ADAM["beta2"]
        return float(loss_pi.detach()), float(loss_v.detach())

    def weights(self):
        return {k: v.detach().numpy().astype(np.float32) for k, v in self.w.items()}
