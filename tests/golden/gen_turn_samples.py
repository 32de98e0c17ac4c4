"""Generates turn_samples.npz FROM THE COMPILED REFERENCE (oracle/_ref/libref_oracle.so): the bytes
NNTrainDataStorage::saveTrainingSamples writes after Script-vs-Script and Script-vs-Random games played by the
reference's own ScriptPlayer / RandomPlayer objects with one shared storage attached (the set-up of
AlphaZeroTrainer::trainOnGeneratedData, alphazero_trainer.cpp:242-268), driven by the RNG contract of
include/az_philox.h.  NNInputData's three padding bytes (43, 46, 47 of the image) are zeroed: they are
uninitialised in the reference.

    python tests/golden/gen_turn_samples.py        # needs /root/reference (oracle/_ref built)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x5EED0001
GAMES = {"script_vs_script": [700, 701], "script_vs_random": [710, 711]}


def play(pairing, game):
    L = po.ref_lib()
    ref = po.RefGame()
    ref.new_game(SEED, game, 0)
    st = L.ref_storage_new()
    rs = [L.ref_script_new(), L.ref_script_new()]
    rr = L.ref_random_new(1)
    for h in rs:
        L.ref_script_set_storage(h, st)
    L.ref_random_set_storage(rr, st)
    ply = 0
    while ref.status() == -1:
        cur = int(ref.data()[146])                                   # Data::currentPlayerTurn
        if pairing == "script_vs_script" or cur == 0:
            assert ref.script_turn(rs[cur], SEED, game, ply) == 0
        else:
            assert ref.random_turn(rr, SEED, game, ply) == 0
        ply += 1
    status = ref.status()
    rounds = int(ref.data()[144:146].view(np.uint16)[0])
    L.ref_script_game_finished(rs[0], status, rounds)
    if pairing == "script_vs_script":
        L.ref_script_game_finished(rs[1], status, rounds)
    else:
        L.ref_random_game_finished(rr, status, rounds)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.bin")
        assert L.ref_storage_save(st, path.encode()) == 0
        raw = np.fromfile(path, np.uint8)
    n = int(raw[:8].view(np.uint64)[0])
    recs = raw[8:].reshape(n, 265).copy()
    recs[:, [1 + 43, 1 + 46, 1 + 47]] = 0
    return recs, status, ply


def main():
    po.ref_apply_rules(po.default_rules())
    out = {}
    for pairing, games in GAMES.items():
        for g in games:
            recs, status, plies = play(pairing, g)
            out["%s_%d_records" % (pairing, g)] = recs
            out["%s_%d_meta" % (pairing, g)] = np.array([status, plies, len(recs)], np.int64)
            print(pairing, g, "status", status, "plies", plies, "samples", len(recs))
    np.savez_compressed(os.path.join(HERE, "turn_samples.npz"), **out)


if __name__ == "__main__":
    main()
