"""CPU suite: the network restatement (oracle/nn_oracle.py, the checker of the CUDA forward) against the reference's own shipped
GraphDef.  tests/golden/graph_V2_5_inference.json is the inference slice of python/model/model_txt_V2_5.pb (extracted by
tests/golden/gen_graph_slice.py); oracle/graphdef_oracle.py executes it op by op.  The TensorFlow kernels behind the ops are restated
(un-vendored dependency), everything else — wiring, variables, shapes, data formats, epsilon, the inference branch of every
BatchNorm — is the reference's artifact."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import graphdef_oracle as go
from oracle import nn_oracle as no

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PB = "/root/reference/python/model/model_txt_V2_5.pb"


@pytest.fixture(scope="module")
def sl():
    return go.load_slice()


def test_slice_reads_exactly_the_variables_of_the_restatement(sl, golden_dir):
    got = dict(go.variables(sl))
    assert sorted(got) == sorted(no.variable_names(5))
    inv = json.load(open(os.path.join(golden_dir, "ckpt_tensors_V2_5.json")))          # the Saver's inventory of the same GraphDef
    shapes = {name: tuple(shape) for name, shape in inv["tensors"]}
    for name, shape in got.items():
        assert shapes[name] == shape, name
    assert got["conv/kernel"] == (3, 3, 13, 256) and got["conv_bn/gamma"] == (7,) and got["dense/kernel"] == (84, 43)
    assert got["dense_1/kernel"] == (42, 256) and got["pi/kernel"] == (1, 1, 256, 2) and got["v/kernel"] == (1, 1, 256, 1)


def test_structural_facts_the_cuda_path_relies_on(sl):
    bns = [n for n in sl["nodes"] if n["op"] == "FusedBatchNormV3"]
    assert len(bns) == 13
    for n in bns:
        stem = n["name"].startswith("conv_bn/")
        assert n["attr"]["data_format"]["s"] == ("NCHW" if stem else "NHWC")        # the stem normalises over the board-row axis
        assert n["attr"]["is_training"]["b"] is False
        assert np.float32(n["attr"]["epsilon"]["f"]) == np.float32(0.001)
    convs = [n for n in sl["nodes"] if n["op"] == "Conv2D"]
    assert len(convs) == 13 and all(n["attr"]["padding"]["s"] == "SAME" and n["attr"]["data_format"]["s"] == "NHWC" for n in convs)
    ifs = [n for n in sl["nodes"] if n["op"] == "IfElseOutput"]
    assert len(ifs) == 13                                                            # every BatchNorm branches on input_training
    assert sum(n["op"] == "Add" for n in sl["nodes"]) == 5 and sum(n["op"] == "Relu" for n in sl["nodes"]) == 14


def test_restatement_equals_the_graphdef(sl):
    for seed, n in ((1, 3), (2, 9)):
        w = go.golden_weights(sl, seed)
        x = np.random.default_rng(100 + seed).random((n, 7, 6, 13)).astype(np.float32)
        p, v = go.run(sl, w, x)
        pr, vr = no.forward(w, x, 5, dtype=torch.float64)
        assert np.abs(p - pr).max() < 1e-9 and np.abs(v - vr).max() < 1e-9          # (epsilon: float32(0.001) in the graph, 0.001 in the restatement)
        p32, v32 = no.forward(w, x, 5, dtype=torch.float32)
        assert np.abs(p - p32).max() < 1e-5 and np.abs(v - v32).max() < 1e-5


def test_golden_forward_vectors(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "graph_forward_V2_5.npz"))
    w = go.golden_weights(sl, int(g["seed"]))
    p, v = go.run(sl, w, g["x"])
    assert np.abs(p - g["policy"]).max() < 1e-12 and np.abs(v - g["value"]).max() < 1e-12
    pr, vr = no.forward(w, g["x"], 5, dtype=torch.float32)
    assert np.abs(pr - g["policy"]).max() < 1e-5 and np.abs(vr - g["value"]).max() < 1e-5
    assert np.allclose(g["policy"].sum(1), 1.0, atol=1e-12) and g["policy"].std(0).max() > 1e-3


@pytest.mark.skipif(not os.path.exists(REF_PB), reason="the reference tree is not present (GPU box)")
def test_committed_slice_is_what_the_extractor_produces(tmp_path, golden_dir):
    out = str(tmp_path / "slice.json")
    subprocess.check_call([sys.executable, os.path.join(golden_dir, "gen_graph_slice.py"), REF_PB, out], stdout=subprocess.DEVNULL)
    assert json.load(open(out)) == json.load(open(os.path.join(golden_dir, "graph_V2_5_inference.json")))
