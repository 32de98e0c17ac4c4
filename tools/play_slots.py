"""configs[0] (-m play --mcts=16 --cg=1000) with different slot counts: time, ticks, share of slot-ticks that still had a game"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_risk_b200 import api

SEED = 0x5EED0001
games = 1000
net = api.Net(blocks=5, seed=1234)
for slots in [int(a) for a in sys.argv[1:]] or [500, 370, 250]:
    rules = api.default_rules(mcts_simulations=16, threads_per_mcts=2, concurrent_descents=2)
    env = api.Env(slots, rules=rules)
    mc = api.Mcts(env, net=net, evaluator=api.EVAL_NN, precision=api.BF16)
    arena = api.Arena(mc, api.OPPONENT_SCRIPT, mirror_games=True)
    arena.play(64, SEED + 7)
    t0 = time.perf_counter()
    r = arena.play(games, SEED)
    dt = time.perf_counter() - t0
    print("slots %4d: %.3f s, %5.1f games/s, ticks %d, %.2f ms per tick, busy slot-ticks %.0f %%, az wins %d"
          % (slots, dt, r["count"] / dt, r["ticks"], dt * 1e3 / r["ticks"], 100.0 * r["az_moves"] / (r["ticks"] * slots), r["win"][0]))
    arena.close(); mc.close(); env.close()
net.close()
