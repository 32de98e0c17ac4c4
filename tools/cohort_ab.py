import os, sys
sys.path.insert(0, ".")
import torch
from alphazero_risk_b200 import api
c = int(os.environ["AZC"])
for n, sims, moves in ((4096, 64, 3), (8192, 64, 2), (16384, 200, 1)):
    env = api.Env(n, rules=api.default_rules(mcts_simulations=sims, threads_per_mcts=1)); env.reset(0x5EED0001)
    net = api.Net(blocks=5, seed=1234); mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16); mc.set_cohorts(c)
    mc.selfplay(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mc.selfplay(moves); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(os.environ.get("TAG", ""), "cohorts", c, "games", n, "sims", sims, "%.1f ms -> %.4g sims/s" % (ms, n * sims * moves / ms * 1e3), mc.counters()["errors"], flush=True)
    mc.close(); net.close(); env.close()
