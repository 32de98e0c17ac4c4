"""a few device-resident bf16 forwards at batch 4096 (ncu target)"""
import sys
import torch
sys.path.insert(0, ".")
from alphazero_risk_b200 import api
n = 4096
net = api.Net(blocks=5, seed=1)
x = torch.rand((n, 546), device="cuda"); pol = torch.empty((n, 43), device="cuda"); val = torch.empty(n, device="cuda")
for _ in range(3):
    net.forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), api.BF16, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(pol.sum()))
