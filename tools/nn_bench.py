"""device-resident network forward throughput (positions/s, TFLOP/s) — run on the GPU box"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from alphazero_risk_b200 import api

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flops = 2 * 42 * (9 * 13 * 256 + 2 * blocks * 9 * 256 * 256 + 256 * 2 + 256) + 2 * (84 * 43 + 42 * 256 + 256)
net = api.Net(blocks=blocks, seed=1)
for n in [int(a) for a in sys.argv[2:]] or [512, 4096, 16384]:
    x = torch.rand((n, 546), device="cuda")
    pol = torch.empty((n, 43), device="cuda"); val = torch.empty(n, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for prec, name in [(api.BF16, "bf16-tcgen05")] + ([(api.FP32, "fp32")] if n <= 4096 else []):
        for _ in range(3):
            net.forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), prec, s)
        torch.cuda.synchronize()
        reps = 10 if prec == api.BF16 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            net.forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), prec, s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("%s blocks=%d n=%d: %.3f ms/forward, %.0f positions/s, %.1f TFLOP/s (conventional count)" %
              (name, blocks, n, ms, n / ms * 1e3, n * flops / ms * 1e3 / 1e12))
