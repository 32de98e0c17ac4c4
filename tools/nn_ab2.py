"""interleaved A/B of tower variants in ONE process (the forward is power-capped, so separate runs drift):
one-CTA kernel on the 56-row layout / CTA pair on the 49-row layout"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from alphazero_risk_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nets = {}
for name, mode, layout in (("single56", "single", "56"), ("pair49", "pair", "49")):
    os.environ["AZ_TC_MODE"] = mode; os.environ["AZ_TC_LAYOUT"] = layout
    nets[name] = api.Net(blocks=5, seed=1)
    nets[name].finalize()
x = torch.rand((n, 546), device="cuda"); pol = torch.empty((n, 43), device="cuda"); val = torch.empty(n, device="cuda")
s = torch.cuda.current_stream().cuda_stream
ref = None
for m in nets:
    nets[m].forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), api.BF16, s)
    torch.cuda.synchronize()
    p, v = pol.cpu().numpy().copy(), val.cpu().numpy().copy()
    if ref is None:
        ref = (p, v)
    print(m, "max |dp| vs single56 %.3e  max |dv| %.3e" % (np.abs(p - ref[0]).max(), np.abs(v - ref[1]).max()))
def run(net, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), api.BF16, s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for m in nets: run(nets[m], 5)
tot = {m: [] for m in nets}
for r in range(5):
    for m in nets:
        tot[m].append(run(nets[m], 40))
for m in nets:
    print(m, "n=%d" % n, " ".join("%.3f" % t for t in tot[m]), "ms; median %.3f" % sorted(tot[m])[len(tot[m]) // 2])
