"""runs the bf16 forward back to back for a few seconds while sampling SM clock / power / throttle reasons (NVML)
    python tools/nn_power.py 4096 4.0            one batch size, seconds
    python tools/nn_power.py 773,1546,3866,4096 3.0   several (the L2-residency experiment: 773 boards = 2 full waves of tile pairs with
                                                  every activation buffer L2-resident, 3866 = 10 full waves out of HBM)"""
import sys, time, threading
import torch
sys.path.insert(0, ".")
from alphazero_risk_b200 import api
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ns = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096]
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
net = api.Net(blocks=5, seed=1)
for n in ns:
  samples, stop = [], False
  x = torch.rand((n, 546), device="cuda"); pol = torch.empty((n, 43), device="cuda"); val = torch.empty(n, device="cuda")
  s = torch.cuda.current_stream().cuda_stream
  def sampler():
      while not stop:
          samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                          pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
          time.sleep(0.05)
  for _ in range(3):
      net.forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), api.BF16, s)
  torch.cuda.synchronize()
  t = threading.Thread(target=sampler); t.start()
  t0 = time.time(); reps = 0
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  while time.time() - t0 < secs:
      for _ in range(20):
          net.forward_dev(x.data_ptr(), n, pol.data_ptr(), val.data_ptr(), api.BF16, s)
      reps += 20
      torch.cuda.synchronize()
  e1.record(); torch.cuda.synchronize()
  stop = True; t.join()
  ms = e0.elapsed_time(e1) / reps
  clk = sorted(c for c, _, _ in samples); pw = sorted(p for _, p, _ in samples)
  reasons = 0
  for _, _, r in samples: reasons |= r
  pairs = (n * 49 + 255) // 256
  print("n=%d (%d tile pairs = %.2f waves of 74) %.3f ms/forward %.0f pos/s | sm clock median %d min %d max %d MHz | power median %.0f max %.0f W | reasons 0x%x | %d samples" %
        (n, pairs, pairs / 74.0, ms, n / ms * 1e3, clk[len(clk) // 2], clk[0], clk[-1], pw[len(pw) // 2], pw[-1], reasons, len(samples)))
