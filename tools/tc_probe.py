"""quick numerical probe of the bf16 tcgen05 path against the fp32 path (run on the GPU box)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from alphazero_risk_b200 import api
from oracle import pyoracle as po


def game_inputs(n, seed=0xBEEF):
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


for blocks, n in [(1, 5), (2, 37), (5, 600)]:
    net = api.Net(blocks=blocks, seed=1234)
    x = game_inputs(n)
    p32, v32 = net.forward(x, api.FP32)
    p16, v16 = net.forward(x, api.BF16)
    print("blocks", blocks, "n", n, "policy max|d|", float(np.abs(p32 - p16).max()), "value max|d|", float(np.abs(v32 - v16).max()),
          "finite", bool(np.isfinite(p16).all() and np.isfinite(v16).all()), "psum", float(p16.sum(1).min()), float(p16.sum(1).max()))
    print("  sample p32", p32[0, :5], "p16", p16[0, :5], "v", v32[:3], v16[:3])
    net.close()
