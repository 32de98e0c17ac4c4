"""small bf16 training steps + a short arena match with a network, for compute-sanitizer runs"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_risk_b200 import api

rng = np.random.default_rng(5)
n = 20
net = api.Net(blocks=1, seed=1234)
net.train_precision(api.BF16)
x = rng.random((n, 7, 6, 13), dtype=np.float32)
tp = rng.random((n, 43)).astype(np.float32); tp /= tp.sum(1, keepdims=True)
tv = rng.choice([-1.0, 0.0, 1.0], n).astype(np.float32)
for _ in range(2):
    print(net.train_step(x, tp, tv))
print(net.train_step(x[:7], tp[:7], tv[:7]))
rules = api.default_rules(mcts_simulations=4, threads_per_mcts=1, concurrent_descents=2)
env = api.Env(6, rules=rules)
mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
mc2 = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.FP32)
arena = api.Arena(mc, mirror_games=True, opponent_mcts=mc2)
print(arena.play(4, 11))
arena.close(); mc2.close(); mc.close(); env.close(); net.close()
