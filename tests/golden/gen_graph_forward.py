"""Golden vectors of the network forward: the reference's shipped GraphDef (inference slice, tests/golden/graph_V2_5_inference.json)
executed by oracle/graphdef_oracle.py in float64 on 12 real positions with the weight set graphdef_oracle.golden_weights(slice, 20261018).
Writes tests/golden/graph_forward_V2_5.npz = {x [12,7,6,13] f32, policy [12,43] f64, value [12] f64, seed}.

    python tests/golden/gen_graph_forward.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import graphdef_oracle as go
from oracle import pyoracle as po

SEED_W, SEED_G = 20261018, 0x5EED0001


def positions(n):
    xs, o, g = [], po.OracleGame(), 0
    while len(xs) < n:
        o.new_game(SEED_G, g, 0)
        ply = 0
        while o.status() == -1 and len(xs) < n:
            if ply % 29 == 11:
                xs.append(o.encode())
            o.move(o.random_action(SEED_G, g, ply), SEED_G, g, ply)
            ply += 1
        g += 1
    return np.array(xs, np.float32).reshape(n, 7, 6, 13)


if __name__ == "__main__":
    sl = go.load_slice()
    x = positions(12)
    policy, value = go.run(sl, go.golden_weights(sl, SEED_W), x)
    np.savez_compressed(os.path.join(HERE, "graph_forward_V2_5.npz"), x=x, policy=policy, value=value, seed=np.int64(SEED_W))
    print("wrote graph_forward_V2_5.npz", policy.shape, value)
