"""TEST INFRASTRUCTURE — ctypes bindings for the CPU oracle (oracle/librisk_oracle.so, the C
restatement in risk_oracle.c) and, when it has been built, for the compiled reference
(oracle/_ref/libref_oracle.so, see oracle/ref/build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package alphazero_risk_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "librisk_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")

LANDS, MOVES, SKIP, NONE = 42, 43, 42, 43
DATA_BYTES, INPUT_FLOATS = 160, 546
STREAM_REAL, STREAM_DEAL = 0xFFFFFFFF, 0xFFFFFFFE
PHASES = ("SETUP", "SETUP_NEUTRAL", "REINFORCEMENT", "ATTACK", "ATTACK_MOBILIZATION", "FORTIFY")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "librisk_oracle.so"])


def build_ref():
    subprocess.check_call(["bash", os.path.join(HERE, "ref", "build_ref.sh")])


class RoState(C.Structure):
    _fields_ = [("land", C.c_uint8 * 42), ("cards", C.c_uint8 * 2), ("round", C.c_uint16), ("cur", C.c_int8),
                ("card_sets", C.c_uint8), ("reinf", C.c_uint8), ("phase", C.c_uint8), ("mob_from", C.c_uint8),
                ("mob_to", C.c_uint8), ("allow_draw", C.c_uint8), ("attacks", C.c_uint8)]


class RoRules(C.Structure):
    _fields_ = [("allow_yield", C.c_int), ("limit_reinforcement", C.c_int), ("limit_attack", C.c_int),
                ("max_game_rounds", C.c_int), ("min_unit_move", C.c_int), ("mcts_simulations", C.c_int),
                ("threads_per_mcts", C.c_int), ("cpuct", C.c_float), ("dir_noise_value", C.c_float),
                ("dir_noise_epsi", C.c_float), ("temperature_threshold", C.c_int)]


class RoScript(C.Structure):
    _fields_ = [("set", C.c_int8), ("to", C.c_int8), ("from_", C.c_int8), ("from_army", C.c_uint8)]


class RoDice(C.Structure):
    _fields_ = [("use_tape", C.c_int), ("seed", C.c_uint64), ("game", C.c_uint32), ("ply", C.c_uint32),
                ("sim", C.c_uint32), ("j", C.c_uint32), ("tape", C.POINTER(C.c_int32)), ("tape_len", C.c_int),
                ("tape_pos", C.c_int)]


class RoTurnSink(C.Structure):
    _fields_ = [("states", C.c_void_p), ("moves", C.c_void_p), ("cap", C.c_int), ("n", C.c_int)]


class TurnSink:
    """what Player::addTrainingSample received during scripted / random turns: (state before the move, move) pairs"""

    def __init__(self, cap=8192):
        self.states = (RoState * cap)()
        self.moves = np.zeros(cap, np.uint8)
        self.c = RoTurnSink(C.cast(self.states, C.c_void_p), self.moves.ctypes.data, cap, 0)

    def __len__(self):
        assert self.c.n <= self.c.cap, "TurnSink overflow"
        return int(self.c.n)

    def records(self, status, oracle_game):
        """the 265-byte records NNTrainDataStorage::saveTrainingSamples writes for these samples once the game ended with `status`"""
        out = np.zeros((len(self), 265), np.uint8)
        keep = RoState.from_buffer_copy(oracle_game.s)
        for i in range(len(self)):
            C.memmove(C.byref(oracle_game.s), C.byref(self.states[i]), C.sizeof(RoState))
            pi = np.zeros(MOVES, np.float32)
            pi[self.moves[i]] = 1.0
            out[i] = oracle_game.sample_record(pi, status)
        C.memmove(C.byref(oracle_game.s), C.byref(keep), C.sizeof(RoState))
        return out


class BenchOut(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("games", C.c_uint64), ("sims", C.c_uint64), ("evals", C.c_uint64),
                ("moves", C.c_uint64), ("seconds", C.c_double)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        sp, rp, dp = C.POINTER(RoState), C.POINTER(RoRules), C.POINTER(RoDice)
        L.ro_default_rules.argtypes = [rp]
        L.ro_dice_philox.argtypes = [dp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ro_export_data.argtypes = [sp, u8p]
        L.ro_import_data.argtypes = [sp, u8p]
        L.ro_data_byte_mask.argtypes = [u8p]
        L.ro_new_game.argtypes = [sp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ro_new_game_tape.argtypes = [sp, i32p]
        L.ro_valid_moves.argtypes = [sp, rp]
        L.ro_valid_moves.restype = C.c_uint64
        L.ro_game_status.argtypes = [sp, rp]
        L.ro_reinforcement_value.argtypes = [C.c_uint64]
        L.ro_make_move.argtypes = [sp, C.c_int, rp, dp]
        L.ro_random_action.argtypes = [sp, rp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ro_encode.argtypes = [sp, f32p]
        L.ro_normalize_policy.argtypes = [f32p, C.c_uint64]
        L.ro_nn_input.argtypes = [sp, u8p]
        L.ro_script_init.argtypes = [C.POINTER(RoScript)]
        L.ro_script_turn.argtypes = [sp, C.POINTER(RoScript), rp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ro_invert_players.argtypes = [sp]
        L.ro_script_turn_rec.argtypes = [sp, C.POINTER(RoScript), rp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(RoTurnSink)]
        L.ro_random_turn_rec.argtypes = [sp, rp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(RoTurnSink)]
        L.ro_random_turn.argtypes = [sp, rp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ro_sample_record.argtypes = [sp, f32p, C.c_int, u8p]
        L.ro_mcts_new.restype = C.c_void_p
        L.ro_mcts_new.argtypes = [C.c_void_p, C.c_void_p]
        for fn in ("ro_mcts_free", "ro_mcts_clear", "ro_mcts_trim", "ro_mcts_table_size"):
            getattr(L, fn).argtypes = [C.c_void_p]
        for fn in ("ro_mcts_vl_skips", "ro_mcts_vl_duplicates"):
            getattr(L, fn).argtypes = [C.c_void_p]
            getattr(L, fn).restype = C.c_uint64
        L.ro_mcts_search.argtypes = [C.c_void_p, sp, rp, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p, f32p, f32p,
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ro_mcts_search_lockstep.argtypes = [C.c_void_p, sp, rp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, u32p, f32p, f32p, f32p,
                                              C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ro_pick_move.argtypes = [f32p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ro_bench_env.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(BenchOut)]
        _oracle = L
    return _oracle


def data_byte_mask():
    m = np.zeros(DATA_BYTES, np.uint8)
    oracle_lib().ro_data_byte_mask(m)
    return m.astype(bool)


def default_rules(**kw):
    r = RoRules()
    oracle_lib().ro_default_rules(C.byref(r))
    for k, v in kw.items():
        setattr(r, k, v)
    return r


class OracleGame:
    """One game driven through the C restatement."""

    def __init__(self, rules=None):
        self.L = oracle_lib()
        self.s = RoState()
        self.rules = rules if rules is not None else default_rules()

    def new_game(self, seed, game, ply=0):
        self.L.ro_new_game(C.byref(self.s), seed, game, ply)

    def new_game_tape(self, draws):
        self.L.ro_new_game_tape(C.byref(self.s), np.ascontiguousarray(draws, np.int32))

    def data(self):
        out = np.zeros(DATA_BYTES, np.uint8)
        self.L.ro_export_data(C.byref(self.s), out)
        return out

    def set_data(self, data):
        return self.L.ro_import_data(C.byref(self.s), np.ascontiguousarray(data, np.uint8))

    def valid(self):
        return int(self.L.ro_valid_moves(C.byref(self.s), C.byref(self.rules)))

    def status(self):
        return int(self.L.ro_game_status(C.byref(self.s), C.byref(self.rules)))

    def random_action(self, seed, game, ply):
        return int(self.L.ro_random_action(C.byref(self.s), C.byref(self.rules), seed, game, ply))

    def move(self, action, seed, game, ply, sim=STREAM_REAL):
        d = RoDice()
        self.L.ro_dice_philox(C.byref(d), seed, game, ply, sim)
        return int(self.L.ro_make_move(C.byref(self.s), action, C.byref(self.rules), C.byref(d)))

    def move_tape(self, action, dice):
        tape = np.ascontiguousarray(dice, np.int32)
        d = RoDice()
        d.use_tape = 1
        d.tape = tape.ctypes.data_as(C.POINTER(C.c_int32))
        d.tape_len = len(tape)
        rc = int(self.L.ro_make_move(C.byref(self.s), action, C.byref(self.rules), C.byref(d)))
        return rc, int(d.tape_pos)

    def encode(self):
        x = np.zeros(INPUT_FLOATS, np.float32)
        self.L.ro_encode(C.byref(self.s), x)
        return x

    def script_turn(self, script, seed, game, ply):
        """ScriptPlayer::takeTurn on this game; `script` = RoScript carrying the player's members between turns"""
        return int(self.L.ro_script_turn(C.byref(self.s), C.byref(script), C.byref(self.rules), seed, game, ply))

    def random_turn(self, seed, game, ply):
        """RandomPlayer::takeTurn on this game"""
        return int(self.L.ro_random_turn(C.byref(self.s), C.byref(self.rules), seed, game, ply))

    def script_turn_rec(self, script, seed, game, ply, sink):
        """ScriptPlayer::takeTurn with its Player::addTrainingSample calls collected by `sink` (TurnSink)"""
        return int(self.L.ro_script_turn_rec(C.byref(self.s), C.byref(script), C.byref(self.rules), seed, game, ply, C.byref(sink.c)))

    def random_turn_rec(self, seed, game, ply, sink):
        return int(self.L.ro_random_turn_rec(C.byref(self.s), C.byref(self.rules), seed, game, ply, C.byref(sink.c)))

    def invert_players(self):
        self.L.ro_invert_players(C.byref(self.s))

    def nn_input(self):
        """NNInputData(const State&) as its 88-byte image"""
        out = np.zeros(88, np.uint8)
        self.L.ro_nn_input(C.byref(self.s), out)
        return out

    def sample_record(self, pi, status):
        """one 265-byte training sample of the reference's file format for the current state, policy `pi`, final `status`"""
        out = np.zeros(265, np.uint8)
        self.L.ro_sample_record(C.byref(self.s), np.ascontiguousarray(pi, np.float32), int(status), out)
        return out


class OracleMcts:
    def __init__(self, rules=None, evaluator="pseudo"):
        self.L = oracle_lib()
        self.rules = rules if rules is not None else default_rules()
        fn = {"pseudo": self.L.ro_eval_pseudo, "uniform": self.L.ro_eval_uniform}[evaluator]
        self.h = self.L.ro_mcts_new(C.cast(fn, C.c_void_p), None)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ro_mcts_free(self.h)
            self.h = None

    def clear(self):
        self.L.ro_mcts_clear(self.h)

    def trim(self):
        self.L.ro_mcts_trim(self.h)

    def table_size(self):
        return int(self.L.ro_mcts_table_size(self.h))

    def vl_counts(self):
        """(moves passed over by the active_N rule, duplicate requests) so far"""
        return int(self.L.ro_mcts_vl_skips(self.h)), int(self.L.ro_mcts_vl_duplicates(self.h))

    def search(self, game_obj, seed, game, ply, lockstep=1):
        """lockstep = K > 1: K descents select (virtual-loss rule) before any is evaluated, ro_mcts_search_lockstep"""
        N = np.zeros(MOVES, np.uint32)
        Q = np.zeros(MOVES, np.float32)
        P = np.zeros(MOVES, np.float32)
        pi = np.zeros(MOVES, np.float32)
        sumN, val = C.c_uint32(0), C.c_float(0)
        if lockstep > 1:
            rc = self.L.ro_mcts_search_lockstep(self.h, C.byref(game_obj.s), C.byref(self.rules), seed, game, ply, lockstep, N, Q, P,
                                                pi, C.byref(sumN), C.byref(val))
        else:
            rc = self.L.ro_mcts_search(self.h, C.byref(game_obj.s), C.byref(self.rules), seed, game, ply, N, Q, P, pi,
                                       C.byref(sumN), C.byref(val))
        assert rc == 0, rc
        return dict(N=N, Q=Q, P=P, pi=pi, sumN=int(sumN.value), value=float(val.value))

    def pick(self, pi, sample, seed, game, ply):
        return int(self.L.ro_pick_move(np.ascontiguousarray(pi, np.float32), int(sample), seed, game, ply))


# --------------------------------------------------------------------------- compiled reference
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_state_new.restype = vp
        L.ref_state_free.argtypes = [vp]
        L.ref_state_get.argtypes = [vp, u8p]
        L.ref_state_set.argtypes = [vp, u8p]
        L.ref_state_newgame_philox.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ref_state_newgame_tape.argtypes = [vp, i32p, C.c_int]
        L.ref_valid_moves.argtypes = [vp]
        L.ref_valid_moves.restype = C.c_uint64
        L.ref_game_status.argtypes = [vp]
        L.ref_make_move_philox.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ref_make_move_tape.argtypes = [vp, C.c_int, i32p, C.c_int]
        L.ref_random_action_philox.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ref_encode.argtypes = [vp, f32p]
        L.ref_normalize_policy.argtypes = [f32p, C.c_uint64]
        L.ref_nn_input.argtypes = [vp, u8p]
        L.ref_script_new.restype = vp
        L.ref_script_free.argtypes = [vp]
        L.ref_script_turn.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ref_state_invert_players.argtypes = [vp]
        L.ref_storage_new.restype = vp
        L.ref_storage_free.argtypes = [vp]
        L.ref_script_set_storage.argtypes = [vp, vp]
        L.ref_random_set_storage.argtypes = [vp, vp]
        L.ref_script_game_finished.argtypes = [vp, C.c_int, C.c_int]
        L.ref_random_game_finished.argtypes = [vp, C.c_int, C.c_int]
        L.ref_storage_count.argtypes = [vp]
        L.ref_storage_count.restype = C.c_long
        L.ref_storage_save.argtypes = [vp, C.c_char_p]
        L.ref_random_new.restype = vp
        L.ref_random_new.argtypes = [C.c_int]
        L.ref_random_free.argtypes = [vp]
        L.ref_random_turn.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ref_state_set_current_player.argtypes = [vp, C.c_int]
        L.ref_save_samples.argtypes = [C.c_char_p, C.c_int, np.ctypeslib.ndpointer(np.int8), u8p, f32p, C.c_int, C.c_int]
        L.ref_consistency_violations.argtypes = [vp]
        L.ref_get_map.argtypes = [np.ctypeslib.ndpointer(np.uint64), np.ctypeslib.ndpointer(np.int8),
                                  np.ctypeslib.ndpointer(np.uint64), i32p]
        L.ref_get_default_settings.argtypes = [f32p]
        L.ref_set_settings.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int]
        L.ref_mcts_new.restype = vp
        L.ref_mcts_new.argtypes = [vp, vp]
        for fn in ("ref_mcts_free", "ref_mcts_clear", "ref_mcts_trim", "ref_mcts_table_size"):
            getattr(L, fn).argtypes = [vp]
        L.ref_mcts_evals.argtypes = [vp]
        L.ref_mcts_evals.restype = C.c_uint64
        L.ref_mcts_search.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p, f32p, f32p,
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ref_mcts_search_lockstep.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, u32p, f32p, f32p, f32p,
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ref_pick_move.argtypes = [vp, f32p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ref_bench_env.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.POINTER(BenchOut)]
        L.ref_bench_selfplay.argtypes = [C.c_int, C.c_uint64, C.c_uint32, vp, vp, C.POINTER(BenchOut)]
        _ref = L
    return _ref


def ref_apply_rules(rules):
    ref_lib().ref_set_settings(rules.mcts_simulations, rules.threads_per_mcts, rules.cpuct, rules.dir_noise_value,
                               rules.dir_noise_epsi, rules.allow_yield, rules.limit_reinforcement, rules.limit_attack,
                               rules.temperature_threshold, rules.max_game_rounds, rules.min_unit_move)


class RefGame:
    """One game driven through the UNMODIFIED reference classes (State / UtilityNN)."""

    def __init__(self):
        self.L = ref_lib()
        self.s = self.L.ref_state_new()

    def __del__(self):
        if getattr(self, "s", None):
            self.L.ref_state_free(self.s)
            self.s = None

    def new_game(self, seed, game, ply=0):
        assert self.L.ref_state_newgame_philox(self.s, seed, game, ply) == 0

    def new_game_tape(self, draws):
        t = np.ascontiguousarray(draws, np.int32)
        assert self.L.ref_state_newgame_tape(self.s, t, len(t)) == 42

    def data(self):
        out = np.zeros(DATA_BYTES, np.uint8)
        self.L.ref_state_get(self.s, out)
        return out

    def set_data(self, data):
        self.L.ref_state_set(self.s, np.ascontiguousarray(data, np.uint8))

    def valid(self):
        return int(self.L.ref_valid_moves(self.s))

    def status(self):
        return int(self.L.ref_game_status(self.s))

    def random_action(self, seed, game, ply):
        return int(self.L.ref_random_action_philox(self.s, seed, game, ply))

    def move(self, action, seed, game, ply):
        return int(self.L.ref_make_move_philox(self.s, action, seed, game, ply))

    def move_tape(self, action, dice):
        t = np.ascontiguousarray(dice, np.int32)
        n = int(self.L.ref_make_move_tape(self.s, action, t, len(t)))
        return (0, n) if n >= 0 else (-1, 0)

    def encode(self):
        x = np.zeros(INPUT_FLOATS, np.float32)
        self.L.ref_encode(self.s, x)
        return x

    def nn_input(self):
        out = np.zeros(88, np.uint8)
        self.L.ref_nn_input(self.s, out)
        return out

    def script_turn(self, script, seed, game, ply):
        """`script` = handle from ref_lib().ref_script_new() (a reference ScriptPlayer object)"""
        return int(self.L.ref_script_turn(script, self.s, seed, game, ply))

    def random_turn(self, player, seed, game, ply):
        """`player` = handle from ref_lib().ref_random_new(side)"""
        return int(self.L.ref_random_turn(player, self.s, seed, game, ply))

    def invert_players(self):
        self.L.ref_state_invert_players(self.s)

    def set_current_player(self, p):
        self.L.ref_state_set_current_player(self.s, int(p))

    def violations(self):
        return int(self.L.ref_consistency_violations(self.s))


class RefMcts:
    def __init__(self, evaluator="pseudo"):
        self.L = ref_lib()
        fn = {"pseudo": self.L.ref_eval_pseudo, "uniform": self.L.ref_eval_uniform}[evaluator]
        self.h = self.L.ref_mcts_new(C.cast(fn, C.c_void_p), None)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_mcts_free(self.h)
            self.h = None

    def clear(self):
        self.L.ref_mcts_clear(self.h)

    def trim(self):
        self.L.ref_mcts_trim(self.h)

    def table_size(self):
        return int(self.L.ref_mcts_table_size(self.h))

    def search(self, game_obj, seed, game, ply, lockstep=1):
        """lockstep = K > 1: K real search threads of the reference taking turns (ref_mcts_search_lockstep)"""
        N = np.zeros(MOVES, np.uint32)
        Q = np.zeros(MOVES, np.float32)
        P = np.zeros(MOVES, np.float32)
        pi = np.zeros(MOVES, np.float32)
        sumN, val = C.c_uint32(0), C.c_float(0)
        if lockstep > 1:
            rc = self.L.ref_mcts_search_lockstep(self.h, game_obj.s, seed, game, ply, lockstep, N, Q, P, pi, C.byref(sumN), C.byref(val))
        else:
            rc = self.L.ref_mcts_search(self.h, game_obj.s, seed, game, ply, N, Q, P, pi, C.byref(sumN), C.byref(val))
        assert rc == 0, self.L.ref_last_error()
        return dict(N=N, Q=Q, P=P, pi=pi, sumN=int(sumN.value), value=float(val.value))

    def pick(self, pi, sample, seed, game, ply):
        return int(self.L.ref_pick_move(self.h, np.ascontiguousarray(pi, np.float32), int(sample), seed, game, ply))


def new_script():
    sp = RoScript()
    oracle_lib().ro_script_init(C.byref(sp))
    return sp


def ref_save_samples(path, players, nn_inputs, policies, status, rounds):
    """the reference's own NNTrainDataStorage: push the samples, updateValues(status), saveTrainingSamples(path)"""
    p = np.ascontiguousarray(players, np.int8)
    rc = ref_lib().ref_save_samples(path.encode(), len(p), p, np.ascontiguousarray(nn_inputs, np.uint8).reshape(-1),
                                    np.ascontiguousarray(policies, np.float32).reshape(-1), int(status), int(rounds))
    if rc != 0:
        raise RuntimeError(ref_lib().ref_last_error().decode())


# --------------------------------------------------------------------------- six-player extension (oracle/risk6_oracle.c)
R6_STATE_BYTES, R6_PLAYERS = 108, 6


class R6State(C.Structure):
    _fields_ = [("army", C.c_uint8 * 42), ("owner", C.c_uint8 * 42), ("cards", C.c_uint8 * 6), ("pool", C.c_uint8 * 6),
                ("round", C.c_uint16), ("cur", C.c_uint8), ("card_sets", C.c_uint8), ("reinf", C.c_uint8), ("phase", C.c_uint8),
                ("mob_from", C.c_uint8), ("mob_to", C.c_uint8), ("allow_draw", C.c_uint8), ("attacks", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Oracle6Game:
    """one six-player game (SIXPLAYER.md) driven through the C statement of the extension; parity unpinned (no reference semantics)"""

    def __init__(self, rules=None):
        self.L = oracle_lib()
        self.L.r6_new_game.argtypes = [C.POINTER(R6State), C.c_uint64, C.c_uint32, C.c_uint32]
        self.L.r6_valid_moves.argtypes = [C.POINTER(R6State), C.POINTER(RoRules)]
        self.L.r6_valid_moves.restype = C.c_uint64
        self.L.r6_game_status.argtypes = [C.POINTER(R6State), C.POINTER(RoRules)]
        self.L.r6_make_move.argtypes = [C.POINTER(R6State), C.c_int, C.POINTER(RoRules), C.c_uint64, C.c_uint32, C.c_uint32]
        self.L.r6_random_action.argtypes = [C.POINTER(R6State), C.POINTER(RoRules), C.c_uint64, C.c_uint32, C.c_uint32]
        self.s = R6State()
        self.rules = rules if rules is not None else default_rules()
        assert C.sizeof(R6State) == R6_STATE_BYTES

    def new_game(self, seed, game, ply=0):
        self.L.r6_new_game(C.byref(self.s), seed, game, ply)

    def valid(self):
        return int(self.L.r6_valid_moves(C.byref(self.s), C.byref(self.rules)))

    def status(self):
        return int(self.L.r6_game_status(C.byref(self.s), C.byref(self.rules)))

    def random_action(self, seed, game, ply):
        return int(self.L.r6_random_action(C.byref(self.s), C.byref(self.rules), seed, game, ply))

    def move(self, action, seed, game, ply):
        return int(self.L.r6_make_move(C.byref(self.s), int(action), C.byref(self.rules), seed, game, ply))

    def image(self):
        return np.frombuffer(bytes(self.s), np.uint8).copy()

    def set_image(self, img):
        C.memmove(C.byref(self.s), np.ascontiguousarray(img, np.uint8).ctypes.data, R6_STATE_BYTES)


class Oracle6Mcts:
    """the six-player search of SIXPLAYER.md on the oracle (table cleared per search, pseudo evaluator unless `eval_fn` is given)"""

    def __init__(self, rules=None, eval_fn=None):
        self.L = oracle_lib()
        self.rules = rules if rules is not None else default_rules()
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C")
        self.L.r6_mcts_new.restype = C.c_void_p
        self.L.r6_mcts_new.argtypes = [C.c_void_p, C.c_void_p]
        self.L.r6_mcts_free.argtypes = [C.c_void_p]
        self.L.r6_mcts_table_size.argtypes = [C.c_void_p]
        self.L.r6_mcts_search.argtypes = [C.c_void_p, C.POINTER(R6State), C.POINTER(RoRules), C.c_uint64, C.c_uint32, C.c_uint32,
                                          np.ctypeslib.ndpointer(np.uint32, flags="C"), f32, f32, f32, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        self.L.r6_encode.argtypes = [C.POINTER(R6State), f32]
        self.h = self.L.r6_mcts_new(C.cast(eval_fn, C.c_void_p) if eval_fn is not None else None, None)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.r6_mcts_free(self.h)
            self.h = None

    def table_size(self):
        return int(self.L.r6_mcts_table_size(self.h))

    def search(self, game_obj, seed, game, ply):
        N = np.zeros(MOVES, np.uint32)
        Q, P, pi = np.zeros(MOVES, np.float32), np.zeros(MOVES, np.float32), np.zeros(MOVES, np.float32)
        sumN, val = C.c_uint32(0), C.c_float(0)
        rc = self.L.r6_mcts_search(self.h, C.byref(game_obj.s), C.byref(self.rules), seed, game, ply, N, Q, P, pi, C.byref(sumN), C.byref(val))
        assert rc == 0, rc
        return dict(N=N, Q=Q, P=P, pi=pi, sumN=int(sumN.value), value=float(val.value))

    def pick(self, pi, sample, seed, game, ply):
        return int(self.L.ro_pick_move(np.ascontiguousarray(pi, np.float32), int(sample), seed, game, ply))

    def encode(self, game_obj):
        x = np.zeros(INPUT_FLOATS, np.float32)
        self.L.r6_encode(C.byref(game_obj.s), x)
        return x
