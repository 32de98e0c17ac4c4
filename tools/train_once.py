"""two training steps at batch 512 on the 5-block graph (for the ncu launch list of the training kernels)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_risk_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(5)
net = api.Net(blocks=5, seed=1234)
if len(sys.argv) > 2 and sys.argv[2] == "bf16":
    net.train_precision(api.BF16)
x = rng.random((n, 7, 6, 13), dtype=np.float32)
tp = rng.random((n, 43)).astype(np.float32); tp /= tp.sum(1, keepdims=True)
tv = rng.choice([-1.0, 0.0, 1.0], n).astype(np.float32)
for _ in range(3):
    print(net.train_step(x, tp, tv))
net.close()
