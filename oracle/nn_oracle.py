"""TEST INFRASTRUCTURE — CPU restatement (PyTorch fp32 / fp64) of the reference's policy/value
network, /root/reference/python/src/build_graph.py:37-90, cross-checked structurally against the
shipped GraphDef python/model/model_txt_V2_5.pb (variable names and shapes, SAME padding, NHWC,
epsilon 0.001, the stem's BatchNorm running over the board-ROW axis: build_graph.py:68 passes
axis=1 on an NHWC tensor, so conv_bn/* have 7 elements).

PARITY UNPINNED: the arithmetic the reference executes lives in TensorFlow (un-vendored, un-pinned,
not installable here) and the reference ships neither a checkpoint nor an example output, so there
is no reference number to pin this restatement to.  It is the checker for the <= 1e-5 fp32
agreement north_star asks for; only tests/ and __graft_entry__.smoke() may import it.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 0.001


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


def _conv(x, k):
    """k: HWIO -> conv2d with SAME padding, stride 1, no bias; x NCHW"""
    kh = k.shape[0]
    return F.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), padding=kh // 2)


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
