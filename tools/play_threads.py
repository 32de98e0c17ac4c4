"""configs[0] with the match split over several host threads, each with its own env / searcher / arena / stream on the same GPU
(every thread has its own copy of the network: a handle is not re-entrant): python tools/play_threads.py <threads> [<threads> ...]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_risk_b200 import api

SEED = 0x5EED0001
games = 1000
for T in [int(a) for a in sys.argv[1:]] or [1, 2]:
    per = games // T
    slots = per // 2
    rules = api.default_rules(mcts_simulations=16, threads_per_mcts=2, concurrent_descents=2)
    parts = []
    for k in range(T):
        env = api.Env(slots, rules=rules, first_game_id=k * slots)
        net = api.Net(blocks=5, seed=1234)
        mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
        arena = api.Arena(mc, api.OPPONENT_SCRIPT, mirror_games=True)
        st = torch.cuda.Stream()
        arena.play(32, SEED + 7, stream=st.cuda_stream)
        parts.append([env, mc, arena, st, None, net])

    def run(p):
        torch.cuda.set_device(0)
        p[4] = p[2].play(per, SEED, stream=p[3].cuda_stream)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(p,)) for p in parts]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    tot = sum(p[4]["count"] for p in parts)
    print("threads %d x %d slots: %.3f s, %5.1f games/s, az wins %d, ticks %s" % (T, slots, dt, tot / dt, sum(p[4]["win"][0] for p in parts), [p[4]["ticks"] for p in parts]))
    for p in parts:
        p[2].close(); p[1].close(); p[0].close(); p[5].close()
