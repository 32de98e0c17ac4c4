"""per-variable gradient error of the two training modes against the fp64 oracle (diagnostic)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from alphazero_risk_b200 import api
from oracle import nn_oracle as no
import test_train_gpu as T

blocks, n = int(sys.argv[1]), int(sys.argv[2])
x, tp, tv = T.positions(n)
res = {}
for mode in (api.FP32, api.BF16):
    net = T.perturbed_net(api, blocks, 77)
    net.train_precision(mode)
    ref = no.Trainer(net.weights(), blocks, bf16=(mode == api.BF16 and len(sys.argv) > 3))     # third argument: bf16-emulating oracle
    lp, lv = net.train_step(x, tp, tv)
    no.TRACE = {}
    rp, rv = ref.step(x, tp, tv)
    names = ["conv_bn"] + ["bn%d%s_branch%s" % (i, chr(97 + i), br) for i in range(blocks) for br in ("2a", "2b")]
    for L, nm in enumerate(names):
        z, zr = net.layer(L, 0, n).astype(np.float64), no.TRACE[nm]
        print("  layer %d z: max err / max %.3e   rel l2 %.3e" % (L, np.abs(z - zr).max() / np.abs(zr).max(), np.linalg.norm(z - zr) / np.linalg.norm(zr)))
    no.TRACE = None
    shapes = dict(net.variables())
    print("mode", mode, "loss", lp, lv, "ref", rp, rv)
    for name in no.trainable_names(blocks):
        g, gr = net.grad(name, shapes[name]).ravel().astype(np.float64), ref.grads[name].numpy().ravel()
        err = np.abs(g - gr).max() / max(np.abs(gr).max(), 1e-12)
        cos = float(g @ gr / max(np.linalg.norm(g) * np.linalg.norm(gr), 1e-30))
        rel2 = np.linalg.norm(g - gr) / max(np.linalg.norm(gr), 1e-30)
        if name.endswith("kernel") or name.startswith("conv_bn"):
            print("  %-28s maxerr/max %.3e  rel l2 %.3e  cos %.6f" % (name, err, rel2, cos))
    net.close()
