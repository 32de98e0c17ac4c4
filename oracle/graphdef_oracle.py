"""TEST INFRASTRUCTURE — executes the inference slice of the reference's own shipped GraphDef (python/model/model_txt_V2_5.pb, extracted
by tests/golden/gen_graph_slice.py into tests/golden/graph_V2_5_inference.json) with a small numpy interpreter of the TensorFlow ops it
contains.  This is the reference's network as the reference's C++ loads it (neural_network/alphazero_nn.cpp:160-187: ReadBinaryProto +
session->Create), not a reading of build_graph.py: wiring, variable names and shapes, padding, data formats, epsilon and the branch an
inference call takes (input_training = false -> every If runs its else_branch) all come from the file.

What remains restated are the kernels of the TensorFlow ops themselves (un-vendored, un-pinned dependency; GraphDef producer 175):
Conv2D (NHWC, HWIO filter, SAME: pad_total = k - 1, pad_before = pad_total // 2), FusedBatchNormV3 with is_training = false
(y = (x - mean) * rsqrt(variance + epsilon) * scale + offset along the channel axis of data_format), Relu, Add, Reshape, MatMul,
BiasAdd, Softmax (last axis), Tanh, Squeeze, Identity — their published definitions, evaluated in float64.  The training slice
(graph_V2_5_training.json: input_training = true, the then_branch of every If) adds FusedBatchNormV3 with is_training = true (batch
statistics; the variance output carries N / (N - 1)), SoftmaxCrossEntropyWithLogits, SquaredDifference, DivNoNan, Select, Sum, ... —
enough to evaluate the two losses, the minimised total (add_6) and the amount every moving statistic is decreased by.  The gradient
sub-graph is NOT interpreted: it is TensorFlow's derivative of exactly this forward, which is what nn_oracle.Trainer's autograd takes.
Only tests/ may import this module.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SLICE = os.path.join(os.path.dirname(HERE), "tests", "golden", "graph_V2_5_inference.json")

DTYPES = {"DT_FLOAT": np.float32, "DT_INT32": np.int32, "DT_INT64": np.int64, "DT_BOOL": np.bool_}


class _Resource:
    def __init__(self, name):
        self.name = name


TRAINING_SLICE = os.path.join(os.path.dirname(HERE), "tests", "golden", "graph_V2_5_training.json")


def load_slice(path=DEFAULT_SLICE):
    return json.load(open(path))


def variables(sl):
    """[(variable name, shape)] in graph order: the VarHandleOp nodes the inference path reads"""
    out = []
    for n in sl["nodes"]:
        if n["op"] == "VarHandleOp":
            out.append((n["attr"].get("shared_name", {}).get("s") or n["name"], tuple(n["attr"]["shape"]["shape"])))
    return out


def _const(t):
    dt = DTYPES[t["dtype"]]
    if "content" in t:
        a = np.frombuffer(bytes(t["content"]), dtype=dt)
    else:
        key = {"DT_FLOAT": "float_val", "DT_INT32": "int_val", "DT_INT64": "int64_val", "DT_BOOL": "bool_val"}[t["dtype"]]
        vals = [(v == "true") if dt is np.bool_ else float(v) if dt is np.float32 else int(v) for v in t.get(key, [])]   # absent: all zero / empty
        a = np.array(vals, dtype=dt)
        count = int(np.prod(t["shape"])) if t["shape"] else 1
        if a.size == 0 and count > 0:
            a = np.zeros(count, dtype=dt)
        if a.size == 1 and count > 1:
            a = np.repeat(a, count)
    return a.reshape(t["shape"])


def _conv2d(x, w, attr):
    assert attr["data_format"]["s"] == "NHWC" and attr["padding"]["s"] == "SAME", attr
    assert [int(v) for v in attr["strides"]["list_i"]] == [1, 1, 1, 1] and [int(v) for v in attr.get("dilations", {"list_i": [1, 1, 1, 1]})["list_i"]] == [1, 1, 1, 1]
    kh, kw = w.shape[:2]
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.pad(x, ((0, 0), (pt, kh - 1 - pt), (pl, kw - 1 - pl), (0, 0)))
    n, h, wd, _ = x.shape
    y = np.zeros((n, h, wd, w.shape[3]), np.float64)
    for ky in range(kh):
        for kx in range(kw):
            y += np.einsum("nhwc,co->nhwo", xp[:, ky:ky + h, kx:kx + wd, :], w[ky, kx])
    return y


def _fused_batch_norm(x, scale, offset, mean, var, attr):
    """outputs (y, batch_mean, batch_variance).  is_training = false: the given statistics; true: the batch mean and the biased batch
    variance normalise, and the variance OUTPUT (what AssignMovingAvg consumes) carries Bessel's correction N / (N - 1)"""
    axis = {"NHWC": 3, "NCHW": 1}[attr["data_format"]["s"]]
    shape = [1, 1, 1, 1]
    shape[axis] = -1
    eps = np.float64(np.float32(attr["epsilon"]["f"]))
    if attr["is_training"]["b"]:
        assert np.float32(attr.get("exponential_avg_factor", {"f": 1.0})["f"]) == 1.0
        red = tuple(d for d in range(4) if d != axis)
        n = x.size // x.shape[axis]
        mean = x.mean(axis=red)
        var = ((x - mean.reshape(shape)) ** 2).mean(axis=red)
        y = (x - mean.reshape(shape)) * (scale / np.sqrt(var + eps)).reshape(shape) + offset.reshape(shape)
        return y, mean, var * n / (n - 1)
    inv = 1.0 / np.sqrt(var + eps)
    return (x - mean.reshape(shape)) * (inv * scale).reshape(shape) + offset.reshape(shape), mean, var


def run(sl, weights, x, training=False, targets=None, fetch=None):
    """feeds input_state = x [n,7,6,13], input_training = training and (training slice) target_policy [n,43] / target_value [n,1] =
    `targets`; returns (output_policy [n,43], output_value [n]) in float64, or the list of tensors named in `fetch`.
    `weights`: dict variable name -> array (any float dtype; used as float64)."""
    assert training == (sl.get("branch", "else_branch") == "then_branch"), "the slice holds one branch of every If only"
    nodes = {n["name"]: n for n in sl["nodes"]}
    memo = {}
    feeds = {"input_state": np.asarray(x, np.float64).reshape(-1, 7, 6, 13), "input_training": np.array(training)}
    if targets is not None:
        feeds["target_policy"] = np.asarray(targets[0], np.float64).reshape(-1, 43)
        feeds["target_value"] = np.asarray(targets[1], np.float64).reshape(-1, 1)

    def tensor(t):
        name, _, out_arg = t.partition("#")         # inlined function bodies name tensors node:output_arg:index; the slice keeps "#output_arg"
        idx = 0
        if ":" in name:
            name, i = name.rsplit(":", 1)
            idx = int(i) + {"batch_mean": 1, "batch_variance": 2}.get(out_arg, 0)     # FusedBatchNormV3: y, batch_mean, batch_variance, ...
        v = node(name)
        return v[idx] if isinstance(v, tuple) else v

    def node(name):
        if name in memo:
            return memo[name]
        n = nodes[name]
        op, a = n["op"], n["attr"]
        ins = n["input"]
        if op == "Placeholder":
            v = feeds[name]
        elif op == "Const":
            v = _const(a["value"]["tensor"])
        elif op == "VarHandleOp":
            v = _Resource(a.get("shared_name", {}).get("s") or name)
        elif op == "ReadVariableOp":
            r = tensor(ins[0])
            assert isinstance(r, _Resource)
            v = np.asarray(weights[r.name], np.float64)
        elif op in ("Identity", "StopGradient"):
            v = tensor(ins[0])
        elif op == "Squeeze":
            v = np.squeeze(tensor(ins[0]))
        elif op == "IfElseOutput":
            assert not bool(tensor(ins[1])), "predicate is true: the then branch (training) is not in the slice"
            v = tensor(ins[0])
        elif op == "Conv2D":
            v = _conv2d(tensor(ins[0]), tensor(ins[1]), a)
        elif op == "IfThenOutput":
            assert bool(tensor(ins[-1])), "predicate is false: the else branch (inference) is not in the slice"
            v = tuple(tensor(t) for t in ins[:-1])
        elif op == "FusedBatchNormV3":
            v = _fused_batch_norm(*[tensor(t) for t in ins[:5]], a)
        elif op == "Shape":
            v = np.array(np.shape(tensor(ins[0])), np.int32)
        elif op == "Sub":
            v = tensor(ins[0]) - tensor(ins[1])
        elif op == "Mul":
            v = tensor(ins[0]) * tensor(ins[1])
        elif op == "Square":
            v = tensor(ins[0]) ** 2
        elif op == "SquaredDifference":
            v = (tensor(ins[0]) - tensor(ins[1])) ** 2
        elif op == "AddN":
            v = sum(tensor(t) for t in ins)
        elif op == "Pack":
            v = np.stack([tensor(t) for t in ins], axis=a.get("axis", {"i": 0})["i"])
        elif op == "Slice":
            t, b, sz = tensor(ins[0]), tensor(ins[1]), tensor(ins[2])
            v = t[tuple(slice(int(b[d]), None if int(sz[d]) < 0 else int(b[d]) + int(sz[d])) for d in range(t.ndim))]
        elif op == "ConcatV2":
            v = np.concatenate([np.atleast_1d(tensor(t)) for t in ins[:-1]], axis=int(tensor(ins[-1])))
        elif op == "Sum":
            axes = tuple(int(d) for d in np.atleast_1d(tensor(ins[1])))
            v = np.sum(tensor(ins[0]), axis=axes, keepdims=bool(a.get("keep_dims", {"b": False})["b"]))
        elif op == "Equal":
            v = np.equal(tensor(ins[0]), tensor(ins[1]))
        elif op == "Fill":
            v = np.full([int(d) for d in np.atleast_1d(tensor(ins[0]))], tensor(ins[1]), dtype=np.float64)
        elif op == "Select":
            v = np.where(tensor(ins[0]), tensor(ins[1]), tensor(ins[2]))
        elif op == "DivNoNan":
            p, q = np.asarray(tensor(ins[0]), np.float64), np.asarray(tensor(ins[1]), np.float64)
            v = np.where(q == 0.0, 0.0, p / np.where(q == 0.0, 1.0, q))
        elif op == "SoftmaxCrossEntropyWithLogits":
            z, lab = tensor(ins[0]), tensor(ins[1])
            ls = z - z.max(axis=1, keepdims=True)
            ls = ls - np.log(np.exp(ls).sum(axis=1, keepdims=True))
            v = (-(lab * ls).sum(axis=1), np.exp(ls) - lab)
        elif op == "Relu":
            v = np.maximum(tensor(ins[0]), 0.0)
        elif op in ("Add", "AddV2"):
            v = tensor(ins[0]) + tensor(ins[1])
        elif op == "Reshape":
            v = tensor(ins[0]).reshape([int(d) for d in tensor(ins[1])])
        elif op == "MatMul":
            p, q = tensor(ins[0]), tensor(ins[1])
            if a.get("transpose_a", {}).get("b"): p = p.T
            if a.get("transpose_b", {}).get("b"): q = q.T
            v = p @ q
        elif op == "BiasAdd":
            assert a.get("data_format", {"s": "NHWC"})["s"] == "NHWC"
            v = tensor(ins[0]) + tensor(ins[1])
        elif op == "Softmax":
            z = tensor(ins[0])
            e = np.exp(z - z.max(axis=-1, keepdims=True))
            v = e / e.sum(axis=-1, keepdims=True)
        elif op == "Tanh":
            v = np.tanh(tensor(ins[0]))
        else:
            raise NotImplementedError("op %s (node %s) is not in the interpreter" % (op, name))
        memo[name] = v
        return v

    if fetch is not None:
        return [np.asarray(tensor(t)) for t in fetch]
    policy, value = node("output_policy"), node("output_value")
    return np.asarray(policy), np.asarray(value).reshape(-1)


def golden_weights(sl, seed):
    """a reproducible, non-trivial weight set for the slice's variables (numpy PCG64 stream): Glorot-uniform kernels, small biases,
    BatchNorm parameters away from their initial values so that every term of the normalisation is exercised"""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in variables(sl):
        if name.endswith("/kernel"):
            fan_in = int(np.prod(shape[:-1])) if len(shape) == 4 else shape[0]
            fan_out = shape[-1] * (int(np.prod(shape[:2])) if len(shape) == 4 else 1)
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w[name] = rng.uniform(-lim, lim, shape).astype(np.float32)
        elif name.endswith("/moving_variance"):
            w[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif name.endswith("/gamma"):
            w[name] = rng.uniform(0.8, 1.2, shape).astype(np.float32)
        else:                                    # beta, moving_mean, bias
            w[name] = rng.uniform(-0.1, 0.1, shape).astype(np.float32)
    return w
