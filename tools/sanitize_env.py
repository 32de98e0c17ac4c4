"""small env / scripted-turn / recorder workload for compute-sanitizer (memcheck, racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_env.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from alphazero_risk_b200 import api

n = 300                                         # not a multiple of the block size: exercises the tail block
env = api.Env(n, first_game_id=11)
env.reset(1)
env.rollout(700)
print("rollout", env.counters())
env.reset(2)
env.record_turns(capacity_samples=n * 2500, max_samples_per_game=4096)
script = np.full((n, 2), api.SCRIPT_INIT, np.uint32)
for ply in range(140):
    st = env.play_turn(api.OPPONENT_SCRIPT, api.OPPONENT_RANDOM, script)
    if (st != -1).all():
        break
recs, dropped = env.turn_samples()
print("turn samples", recs.shape, dropped, "running", int((st == -1).sum()))
mc = api.Mcts(env, evaluator=api.EVAL_PSEUDO)
env.reset(3)
for _ in range(6):
    mc.search(pick_mode=api.PICK_SELFPLAY, apply_move=True)
print("mcts", mc.counters())
mc.close(); env.close()
print("ok")
