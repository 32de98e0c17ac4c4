import sys, time, threading, torch
sys.path.insert(0, '.')
from alphazero_risk_b200 import api
n_sh = int(sys.argv[1]); n = 65536; S = 512; per = n // n_sh
shards = []
for k in range(n_sh):
    st = torch.cuda.Stream()
    e = api.Env(per, device=0, first_game_id=k * per)
    a = torch.empty((per, 160), dtype=torch.uint8).pin_memory(); b = torch.empty((per, 160), dtype=torch.uint8).pin_memory()
    e.reset(1, stream=st.cuda_stream); e.export_aos(out=a.numpy(), stream=st.cuda_stream)
    shards.append([e, st, a, b])
T = [[0.0] * 4 for _ in range(n_sh)]
def work(i, steps):
    e, st, hi, ho = shards[i]
    torch.cuda.set_device(0)
    for _ in range(steps):
        t0 = time.perf_counter(); e.import_aos(hi.numpy(), stream=st.cuda_stream)
        t1 = time.perf_counter(); e.rollout(S, stream=st.cuda_stream)
        t2 = time.perf_counter(); e.export_aos(out=ho.numpy(), stream=st.cuda_stream)
        t3 = time.perf_counter(); e.counters(stream=st.cuda_stream)
        t4 = time.perf_counter()
        T[i][0] += t1 - t0; T[i][1] += t2 - t1; T[i][2] += t3 - t2; T[i][3] += t4 - t3
        hi, ho = ho, hi
def run(steps):
    th = [threading.Thread(target=work, args=(i, steps)) for i in range(n_sh)]
    [t.start() for t in th]; [t.join() for t in th]
run(2)
for t in T: t[:] = [0.0] * 4
steps = 10
t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
print("shards %d: %.3f ms/step, %.3e steps/s; per call ms (import, rollout, export, counters): %s" % (
    n_sh, dt / steps * 1e3, n * S * steps / dt, [round(sum(T[i][c] for i in range(n_sh)) / n_sh / steps * 1e3, 3) for c in range(4)]))
