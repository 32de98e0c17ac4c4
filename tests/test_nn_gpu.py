"""GPU suite: the policy/value network forward through the C ABI against the PyTorch CPU
restatement of the reference graph (oracle/nn_oracle.py), random-init weights (the graph's own
init op: Glorot-uniform kernels, BN identity) and perturbed BN statistics."""
import numpy as np
import pytest

from oracle import nn_oracle as no
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5       # north_star: network outputs within 1e-5 (fp32)


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    return a


def game_inputs(n, seed=0xBEEF):
    """encoded positions from seeded random play (the network's real input distribution)"""
    xs, o, g = [], po.OracleGame(), 0
    while len(xs) < n:
        o.new_game(seed, g, 0)
        ply = 0
        while o.status() == -1 and len(xs) < n:
            if ply % 9 == 0:
                xs.append(o.encode())
            o.move(o.random_action(seed, g, ply), seed, g, ply)
            ply += 1
        g += 1
    return np.array(xs, np.float32).reshape(n, 7, 6, 13)


def test_variable_inventory_matches_graph(api):
    net = api.Net(blocks=5)
    names = [n for n, _ in net.variables()]
    assert names == no.variable_names(5)
    shapes = dict(net.variables())
    assert shapes["conv/kernel"] == (3, 3, 13, 256) and shapes["conv_bn/gamma"] == (7,)
    assert shapes["res4e_branch2b/kernel"] == (3, 3, 256, 256) and shapes["dense/kernel"] == (84, 43)
    assert shapes["dense_1/kernel"] == (42, 256) and shapes["dense_2/kernel"] == (256, 1) and shapes["pi/kernel"] == (1, 1, 256, 2)
    # VarHandleOp inventory of the shipped 5-block GraphDef (python/model/model_txt_V2_5.pb):
    # 5 954 160 elements, 5 949 020 of them trainable (everything but the BN moving statistics)
    trainable = sum(int(np.prod(s)) for n, s in net.variables() if "moving_" not in n)
    assert trainable == 5949020 and net.num_params() == 5954160
    net.close()


def test_forward_matches_the_golden_vectors_of_the_shipped_graphdef(api, golden_dir):
    """tests/golden/graph_forward_V2_5.npz: the reference's own GraphDef (python/model/model_txt_V2_5.pb, inference slice) executed by
    oracle/graphdef_oracle.py in float64 on 12 real positions.  fp32 path within 1e-5, bf16 tensor-core path within its stated tolerance;
    the variable inventory and shapes of the library equal the GraphDef's VarHandleOps."""
    import os
    from oracle import graphdef_oracle as go
    g = np.load(os.path.join(golden_dir, "graph_forward_V2_5.npz"))
    sl = go.load_slice()
    net = api.Net(blocks=5)
    assert dict(net.variables()) == dict(go.variables(sl))
    for name, value in go.golden_weights(sl, int(g["seed"])).items():
        net.load(name, value)
    p32, v32 = net.forward(g["x"], api.FP32)
    assert np.abs(p32 - g["policy"]).max() <= FP32_TOL and np.abs(v32 - g["value"]).max() <= FP32_TOL
    p16, v16 = net.forward(g["x"], api.BF16)
    assert np.abs(p16 - g["policy"]).max() <= 2e-3 and np.abs(v16 - g["value"]).max() <= 1e-2
    net.close()


@pytest.mark.parametrize("blocks,n", [(5, 37), (2, 1), (5, 130)])
def test_fp32_forward_matches_torch(api, blocks, n):
    net = api.Net(blocks=blocks, seed=1234)
    rng = np.random.default_rng(7)
    # make BN non-trivial: the random-init identity statistics would hide indexing mistakes
    for name, shape in net.variables():
        if name.endswith("/gamma"):
            net.load(name, rng.uniform(0.8, 1.2, shape))
        elif name.endswith("/beta") or name.endswith("/moving_mean") or name.endswith("/bias"):
            net.load(name, rng.uniform(-0.1, 0.1, shape))
        elif name.endswith("/moving_variance"):
            net.load(name, rng.uniform(0.7, 1.3, shape))
    w = net.weights()
    x = game_inputs(n)
    pol, val = net.forward(x, api.FP32)
    rp, rv = no.forward(w, x, blocks)
    assert np.abs(pol.sum(1) - 1).max() < 1e-5
    assert np.abs(pol - rp).max() <= FP32_TOL, np.abs(pol - rp).max()
    assert np.abs(val - rv).max() <= FP32_TOL, np.abs(val - rv).max()
    # batch invariance (the reference's commented-out testBatchNormalization, alphazero_risk.cpp:114-158)
    p1, v1 = net.forward(x[:1], api.FP32)
    assert (p1[0] == pol[0]).all() and v1[0] == val[0]
    net.close()


def test_random_init_is_the_graph_init(api):
    net = api.Net(blocks=5, seed=42)
    w = net.weights()
    lim = {"conv/kernel": 0.049783, "res0a_branch2a/kernel": 0.036084, "pi/kernel": 0.152499, "dense/kernel": 0.217357,
           "v/kernel": 0.152795, "dense_1/kernel": 0.141895, "dense_2/kernel": 0.152795}
    for k, l in lim.items():
        m = np.abs(w[k]).max()
        assert 0.95 * l < m <= l * (1 + 1e-4) and abs(w[k].mean()) < 0.1 * l
    assert (w["conv_bn/gamma"] == 1).all() and (w["bn_pi/moving_variance"] == 1).all() and (w["dense/bias"] == 0).all()
    blob = net.export_blob()
    net2 = api.Net(blocks=5)
    net2.import_blob(blob)
    x = game_inputs(5)
    a, b = net.forward(x), net2.forward(x)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    net.close(); net2.close()


BF16_POLICY_TOL = 2e-3   # stated bf16 tolerance: |softmax prob| difference vs the fp32 path
BF16_VALUE_TOL = 1e-2    # |tanh value| difference vs the fp32 path


@pytest.mark.parametrize("blocks,n", [(1, 5), (5, 3), (5, 600), (3, 147)])
def test_bf16_tcgen05_forward_close_to_fp32(api, blocks, n):
    """the tensor-core tower (bf16 operands, fp32 TMEM accumulation, fp32 BN/heads) against the fp32
    validation path on the same weights; also exercises tiles that end inside / beyond the batch"""
    net = api.Net(blocks=blocks, seed=99)
    rng = np.random.default_rng(3)
    for name, shape in net.variables():
        if name.endswith("/gamma"):
            net.load(name, rng.uniform(0.8, 1.2, shape))
        elif name.endswith("/beta") or name.endswith("/moving_mean") or name.endswith("/bias"):
            net.load(name, rng.uniform(-0.1, 0.1, shape))
        elif name.endswith("/moving_variance"):
            net.load(name, rng.uniform(0.7, 1.3, shape))
    x = game_inputs(n)
    p32, v32 = net.forward(x, api.FP32)
    p16, v16 = net.forward(x, api.BF16)
    assert np.isfinite(p16).all() and np.isfinite(v16).all()
    assert np.abs(p16.sum(1) - 1).max() < 1e-5
    assert np.abs(p32 - p16).max() <= BF16_POLICY_TOL, np.abs(p32 - p16).max()
    assert np.abs(v32 - v16).max() <= BF16_VALUE_TOL, np.abs(v32 - v16).max()
    # a smaller batch after a larger one reuses the (now dirty) activation buffers
    p16b, v16b = net.forward(x[:2], api.BF16)
    assert (p16b == p16[:2]).all() and (v16b == v16[:2]).all()
    net.close()
