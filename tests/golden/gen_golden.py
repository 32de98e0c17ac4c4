"""Generates the golden fixtures in this directory FROM THE COMPILED REFERENCE
(oracle/_ref/libref_oracle.so = the unmodified /root/reference sources + the overlays of
oracle/ref/).  The reference ships no golden vectors of its own, so these
are the pinned outputs of its code on seeded inputs.

    python tests/golden/gen_golden.py        # needs /root/reference (builds oracle/_ref if missing)

Fixtures (all numpy .npz, a few hundred KB in total):
  env_trace.npz   per step of seeded random-play games: Data image before (160 B, padding
                  zeroed), legal-move mask, action, the dice the reference consumed, Data after,
                  game status after; plus the 42 deal draws and the dealt position of every game
  encode.npz      Data image -> the reference's [7,6,13] input tensor for sampled states
  mcts_trace.npz  self-play / play-mode games searched with the pseudo network: root state,
                  N/Q/P/pi (bit patterns), sumN, root value, chosen move, table size per move
  mcts_trace_deep.npz  (`--deep`) the same for one full self-play game at 200 and one at 800
                  simulations per move
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x5EED0001


def philox_dice(seed, game, ply, sim, n):
    """first n dice of the (seed, game, ply, sim) stream: same arithmetic as include/az_philox.h"""
    out = []
    for j in range(n):
        blk = philox4x32_10(game, ply, sim, j // 20, seed & 0xFFFFFFFF, seed >> 32)
        w = blk[(j // 5) % 4]
        for _ in range(j % 5 + 1):
            p = w * 6
            dgt, w = p >> 32, p & 0xFFFFFFFF
        out.append(dgt + 1)
    return out


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & M, (p0 >> 32) ^ c3 ^ k1, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c0, c1, c2, c3


def deal_draws(seed, game, ply):
    out = []
    for i in range(42):
        blk = philox4x32_10(game, ply, po.STREAM_DEAL, i >> 2, seed & 0xFFFFFFFF, seed >> 32)
        out.append((blk[i & 3] * (42 - i)) >> 32)
    return out


def gen_env(n_games=6):
    mask = po.data_byte_mask()
    rules = po.default_rules()
    po.ref_apply_rules(rules)
    rec = dict(before=[], valid=[], action=[], dice=[], n_dice=[], after=[], status=[], game=[], ply=[])
    deals, dealt = [], []
    enc_data, enc_x = [], []
    for g in range(n_games):
        r = po.RefGame()
        draws = deal_draws(SEED, g, 0)
        r.new_game_tape(draws)
        chk = po.RefGame(); chk.new_game(SEED, g, 0)
        assert (chk.data()[mask] == r.data()[mask]).all()
        deals.append(draws); dealt.append(r.data() * mask)
        ply = 0
        while r.status() == -1:
            before, valid = r.data() * mask, r.valid()
            a = r.random_action(SEED, g, ply)
            dice = philox_dice(SEED, g, ply, po.STREAM_REAL, 5)
            rc, used = r.move_tape(a, dice)
            assert rc == 0
            assert r.violations() == 0
            rec["before"].append(before); rec["valid"].append(valid); rec["action"].append(a)
            rec["dice"].append(dice); rec["n_dice"].append(used); rec["after"].append(r.data() * mask)
            rec["status"].append(r.status()); rec["game"].append(g); rec["ply"].append(ply)
            if ply % 5 == 0:
                enc_data.append(r.data() * mask); enc_x.append(r.encode())
            ply += 1
    np.savez_compressed(os.path.join(HERE, "env_trace.npz"),
                        before=np.array(rec["before"], np.uint8), valid=np.array(rec["valid"], np.uint64),
                        action=np.array(rec["action"], np.uint8), dice=np.array(rec["dice"], np.uint8),
                        n_dice=np.array(rec["n_dice"], np.uint8), after=np.array(rec["after"], np.uint8),
                        status=np.array(rec["status"], np.int8), game=np.array(rec["game"], np.uint32),
                        ply=np.array(rec["ply"], np.uint32), deal_draws=np.array(deals, np.int32),
                        dealt=np.array(dealt, np.uint8), seed=np.uint64(SEED))
    np.savez_compressed(os.path.join(HERE, "encode.npz"), data=np.array(enc_data, np.uint8),
                        x=np.array(enc_x, np.float32).reshape(-1, 7, 6, 13))
    print("env_trace: %d steps, encode: %d states" % (len(rec["action"]), len(enc_x)))


def gen_mcts(deep=False):
    """deep = the 200 / 800 simulation traces (SURVEY 8c iv), kept in their own file: minutes of reference time"""
    mask = po.data_byte_mask()
    cases = [("selfplay16", 16, 1, False, 0), ("selfplay64", 64, 1, False, 1), ("play32t2", 33, 2, True, 2)]
    if deep:
        cases = [("selfplay200", 200, 1, False, 3), ("selfplay800", 800, 1, False, 4)]
    out = {}
    for name, sims, T, play_mode, g in cases:
        rules = po.default_rules(mcts_simulations=sims, threads_per_mcts=T)
        po.ref_apply_rules(rules)
        r, m = po.RefGame(), po.RefMcts("pseudo")
        r.new_game(SEED, g, 0)
        ply, last_cur = 0, None
        rec = dict(root=[], N=[], Q=[], P=[], pi=[], sumN=[], value=[], move=[], table=[], trimmed=[])
        while r.status() == -1:
            cur = int(r.data()[146])
            trimmed = play_mode and cur != last_cur
            if trimmed:  # AlphaZeroPlayer::takeTurn trims once at the start of a turn
                m.trim(); last_cur = cur
            root = r.data() * mask
            res = m.search(r, SEED, g, ply)
            rnd = int(root[144]) | int(root[145]) << 8
            sample = (not play_mode) and rnd <= rules.temperature_threshold
            mv = m.pick(res["pi"], sample, SEED, g, ply)
            assert r.move(mv, SEED, g, ply) == 0
            rec["root"].append(root); rec["N"].append(res["N"]); rec["Q"].append(res["Q"].view(np.uint32))
            rec["P"].append(res["P"].view(np.uint32)); rec["pi"].append(res["pi"].view(np.uint32))
            rec["sumN"].append(res["sumN"]); rec["value"].append(np.float32(res["value"]).view(np.uint32))
            rec["move"].append(mv); rec["table"].append(m.table_size()); rec["trimmed"].append(trimmed)
            ply += 1
        out[name + "_cfg"] = np.array([sims, T, int(play_mode), g], np.int64)
        for k, v in rec.items():
            out[name + "_" + k] = np.array(v)
        out[name + "_final"] = r.data() * mask
        out[name + "_status"] = np.int64(r.status())
        print(name, "moves", ply, "status", r.status())
    out["seed"] = np.uint64(SEED)
    np.savez_compressed(os.path.join(HERE, "mcts_trace_deep.npz" if deep else "mcts_trace.npz"), **out)


if __name__ == "__main__":
    if not po.ref_available():
        po.build_ref()
    if "--deep" in sys.argv:
        gen_mcts(deep=True)
    else:
        gen_env()
        gen_mcts()
