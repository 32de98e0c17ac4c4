/*
 * TEST INFRASTRUCTURE — C-ABI shim around the UNMODIFIED reference classes, linked
 * into oracle/_ref/libref_oracle.so by oracle/ref/build_ref.sh.  Nothing in the
 * product path (alphazero_risk_b200/) may load this library; it exists so that the
 * oracle restatement, the golden vectors and the CUDA kernels can be checked against
 * the reference's own code, and so that bench.py can time the reference's CPU path.
 *
 * Compiled with -fno-access-control so it can reach State::data and
 * AlphaZeroMCTS::search/setRootState without touching the reference sources.
 *
 * What it drives (reference file:line):
 *   State::newGame                 state/state.cpp:137-167
 *   UtilityNN::getValidMoves       player/alpha_zero/alphazero_moves.cpp:3-70
 *   UtilityNN::makeMove            player/alpha_zero/alphazero_moves.cpp:72-233
 *   State::gameStatus              state/state.cpp:518-565
 *   NNInputData(State) + setInStateTensor   alphazero_nn_data.cpp:165-196, alphazero_nn.cpp:31-67
 *   NNOutputData::normalize        alphazero_nn_data.cpp:3-27
 *   AlphaZeroMCTS::setRootState / search    alphazero_mcts.cpp:289-307, 322-377
 *   StateSimulations::calculateMoveProbability, pickHigestWeightedMove, pickRandomWeightedMove
 *                                  alphazero_mcts.cpp:121-148, 397-412, 379-395
 */
#include <cstring>
#include <cstdint>
#include <chrono>
#include <atomic>
#include <thread>
#include <vector>
#include <mutex>
#include <condition_variable>

#include "risk_game/player/alpha_zero/alphazero_mcts.h"
#include "risk_game/player/alpha_zero/neural_network/alphazero_nn.h"
#include "risk_game/player/script/script_player.h"
#include "risk_game/player/random/random_player.h"

#include "az_philox.h"
#include "az_pseudo_net.h"

#define REF_API extern "C" __attribute__((visibility("default")))

static thread_local char g_err[512];

REF_API const char* ref_last_error() { return g_err; }

REF_API int ref_sizeof_data() { return (int)sizeof(Data); }
REF_API int ref_sizeof_player_status() { return (int)sizeof(PlayerStatus); }
REF_API int ref_sizeof_nn_input() { return (int)sizeof(NNInputData); }

/* ---------------------------------------------------------------- settings */
REF_API void ref_set_settings(int mcts_sims, int threads_per_mcts, float cpuct, float dir_noise_value,
                              float dir_noise_epsi, int allow_yield, int limit_reinforcement, int limit_attack,
                              int temperature_threshold, int max_game_rounds, int min_unit_move)
{
	SETTINGS.MCTS_SIMULATIONS = mcts_sims;
	SETTINGS.THREADS_PER_MCTS = threads_per_mcts;
	SETTINGS.HP_EXPLORATION = cpuct;
	SETTINGS.DIR_NOISE_VALUE = dir_noise_value;
	SETTINGS.DIR_NOISE_EPSI = dir_noise_epsi;
	SETTINGS.ALLOW_YIELD = allow_yield != 0;
	SETTINGS.LIMIT_REINFORCEMENT_MOVES = limit_reinforcement != 0;
	SETTINGS.LIMIT_ATTACK_MOVES = limit_attack != 0;
	SETTINGS.TEMPERATURE_TRESHOLD = temperature_threshold;
	SETTINGS.MAX_GAME_ROUNDS = max_game_rounds;
	SETTINGS.MIN_UNIT_MOVE = min_unit_move;
}

/* defaults of settings.h:40-62 as a flat array, so tests can assert them */
REF_API void ref_get_default_settings(float* out11)
{
	Settings s = Settings();
	out11[0] = (float)s.MCTS_SIMULATIONS; out11[1] = (float)s.THREADS_PER_MCTS; out11[2] = s.HP_EXPLORATION;
	out11[3] = s.DIR_NOISE_VALUE; out11[4] = s.DIR_NOISE_EPSI; out11[5] = s.ALLOW_YIELD; out11[6] = s.LIMIT_REINFORCEMENT_MOVES;
	out11[7] = s.LIMIT_ATTACK_MOVES; out11[8] = (float)s.TEMPERATURE_TRESHOLD; out11[9] = (float)s.MAX_GAME_ROUNDS;
	out11[10] = (float)s.MIN_UNIT_MOVE;
}

/* ---------------------------------------------------------------- map tables */
REF_API void ref_get_map(uint64_t* nbr_mask42, int8_t* nbr_list42x6, uint64_t* continent_mask6, int32_t* continent_bonus6)
{
	for (int i = 0; i < LAND_INDEX_SIZE; i++)
	{
		const Land* l = Land::getLand((uint8_t)i);
		nbr_mask42[i] = l->neighboursLandIndexBitMask;
		for (int k = 0; k < 6; k++)
			nbr_list42x6[i * 6 + k] = k < (int)l->neihboursLandIndex.size() ? (int8_t)Utility::li2i(l->neihboursLandIndex[k]) : (int8_t)-1;
	}
	const LandSet* sets[6] = { &LandSet::NORTH_AMERICA, &LandSet::SOUTH_AMERICA, &LandSet::EUROPE, &LandSet::AFRICA, &LandSet::ASIA, &LandSet::AUSTRALIA };
	const int bonus[6] = { NORTH_AMERICA_REINFORCEMENT, SOUTH_AMERICA_REINFORCEMENT, EUROPE_REINFORCEMENT, AFRICA_REINFORCEMENT, ASIA_REINFORCEMENT, AUSTRALIA_REINFORCEMENT };
	for (int i = 0; i < 6; i++) { continent_mask6[i] = sets[i]->landSetIndexBitMask; continent_bonus6[i] = bonus[i]; }
}

/* ---------------------------------------------------------------- rng context */
static void rng_philox(uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim)
{
	RefRngCtx& c = ref_rng_ctx();
	c.mode = REF_RNG_PHILOX; c.seed = seed; c.game = game; c.ply = ply; c.sim = sim; c.die_j = 0; c.deal_i = 0;
}

static void rng_tape(const int32_t* tape, int n)
{
	RefRngCtx& c = ref_rng_ctx();
	c.mode = REF_RNG_TAPE; c.tape = tape; c.tape_len = (size_t)n; c.tape_pos = 0;
}

static void rng_engine(uint32_t seed)
{
	RefRngCtx& c = ref_rng_ctx();
	c.mode = REF_RNG_ENGINE; c.engine.seed(seed);
}

/* ---------------------------------------------------------------- state */
REF_API void* ref_state_new() { State* s = new State(); s->setLog(false); return s; }
REF_API void ref_state_free(void* s) { delete (State*)s; }
REF_API void ref_state_get(void* s, uint8_t* data) { memcpy(data, &((State*)s)->data, sizeof(Data)); }
REF_API void ref_state_set(void* s, const uint8_t* data)
{
	State* st = (State*)s;
	memcpy(&st->data, data, sizeof(Data));
	st->hash = 0;
}
REF_API void ref_state_copy(void* dst, void* src) { *(State*)dst = *(State*)src; }

REF_API int ref_state_newgame_philox(void* s, uint64_t seed, uint32_t game, uint32_t ply)
{
	try
	{
		State* st = (State*)s;
		*st = State();
		st->setLog(false);
		rng_philox(seed, game, ply, AZ_STREAM_DEAL);
		st->newGame();
		return 0;
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

REF_API int ref_state_newgame_tape(void* s, const int32_t* tape, int n)
{
	try
	{
		State* st = (State*)s;
		*st = State();
		st->setLog(false);
		rng_tape(tape, n);
		st->newGame();
		return (int)ref_rng_ctx().tape_pos;
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

REF_API uint64_t ref_valid_moves(void* s) { return UtilityNN::getValidMoves(*(State*)s); }
REF_API int ref_game_status(void* s) { return ((State*)s)->gameStatus(); }
REF_API uint64_t ref_state_hash(void* s) { State* st = (State*)s; st->hash = 0; return st->getHash(); }
REF_API int ref_state_equal(void* a, void* b) { return ((State*)a)->equalFields(*(State*)b) ? 1 : 0; }

/* action: 0..41 land, 42 skip.  Returns 0, or -1 if the reference threw (message in ref_last_error). */
REF_API int ref_make_move_philox(void* s, int action, uint64_t seed, uint32_t game, uint32_t ply)
{
	try
	{
		rng_philox(seed, game, ply, AZ_STREAM_REAL);
		UtilityNN::makeMove(*(State*)s, Utility::i2li((uint8_t)action));
		return 0;
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

/* dice served from an explicit tape; returns number of dice consumed or -1 */
REF_API int ref_make_move_tape(void* s, int action, const int32_t* dice, int n_dice)
{
	try
	{
		rng_tape(dice, n_dice);
		UtilityNN::makeMove(*(State*)s, Utility::i2li((uint8_t)action));
		return (int)ref_rng_ctx().tape_pos;
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

/* uniform-random legal action of the contract: k-th set bit, k = mulhi(word1, popcount) */
REF_API int ref_random_action_philox(void* s, uint64_t seed, uint32_t game, uint32_t ply)
{
	uint64_t m = UtilityNN::getValidMoves(*(State*)s);
	az_u32x4 b = az_rng_block(seed, game, ply, AZ_STREAM_REAL, 0);
	uint32_t k = az_mulhi32(b.y, (uint32_t)Utility::popcount(m));
	for (uint32_t i = 0; i < k; i++) m &= m - 1;
	return (int)Utility::lm2i(Utility::getFirstBitMask(m));
}

/* the reference's own encoder: NNInputData ctor (stage 1) + setInStateTensor (stage 2) -> [7,6,13] */
REF_API void ref_encode(void* s, float* x546)
{
	NNInputData in(*(State*)s);
	tensorflow::Tensor t = UtilityNN::buildInTensor(in);
	memcpy(x546, t.f->data(), sizeof(float) * TF_INPUT_TENSOR_SIZE);
}

REF_API void ref_nn_input(void* s, uint8_t* out) { NNInputData in(*(State*)s); memcpy(out, &in, sizeof in); }

/* ---- ScriptPlayer (player/script/script_player.cpp): one object per (game slot, side), like GameGroup::threadPlayGame keeps them */
REF_API void* ref_script_new() { return new ScriptPlayer(); }
REF_API void ref_script_free(void* p) { delete (ScriptPlayer*)p; }
/* ScriptPlayer::takeTurn(state) with dice / ints from the scripted-opponent streams of (seed, game, ply) */
REF_API int ref_script_turn(void* p, void* s, uint64_t seed, uint32_t game, uint32_t ply)
{
	RefRngCtx& c = ref_rng_ctx();
	c.mode = REF_RNG_PHILOX; c.seed = seed; c.game = game; c.ply = ply; c.sim = AZ_STREAM_OPP; c.die_j = 0; c.int_j = 0;
	try { ((ScriptPlayer*)p)->takeTurn(*(State*)s); }
	catch (std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); c.mode = REF_RNG_ENGINE; c.sim = AZ_STREAM_REAL; return -1; }
	c.mode = REF_RNG_ENGINE; c.sim = AZ_STREAM_REAL;
	return 0;
}
/* ---- RandomPlayer (player/random/random_player.cpp:22-111); playerIndexTurn = the side it plays (Game::addPlayer, game.cpp:86-90) */
REF_API void* ref_random_new(int side) { RandomPlayer* p = new RandomPlayer(); p->playerIndexTurn = (int8_t)side; return p; }
REF_API void ref_random_free(void* p) { delete (RandomPlayer*)p; }
REF_API int ref_random_turn(void* p, void* s, uint64_t seed, uint32_t game, uint32_t ply)
{
	RefRngCtx& c = ref_rng_ctx();
	c.mode = REF_RNG_PHILOX; c.seed = seed; c.game = game; c.ply = ply; c.sim = AZ_STREAM_OPP; c.die_j = 0; c.int_j = 0;
	try { ((RandomPlayer*)p)->takeTurn(*(State*)s); }
	catch (std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); c.mode = REF_RNG_ENGINE; c.sim = AZ_STREAM_REAL; return -1; }
	c.mode = REF_RNG_ENGINE; c.sim = AZ_STREAM_REAL;
	return 0;
}
/* ---- Player::addTrainingSample / gameFinished (player/base/player.cpp:9-17, script_player.cpp:229-235, random_player.cpp:113-119):
   a real NNTrainDataStorage attached to real players, as AlphaZeroTrainer::trainOnGeneratedData does (alphazero_trainer.cpp:227-275) */
REF_API void* ref_storage_new() { return new NNTrainDataStorage(); }
REF_API void ref_storage_free(void* st) { delete (NNTrainDataStorage*)st; }
REF_API void ref_script_set_storage(void* p, void* st) { ((ScriptPlayer*)p)->setTrainStorage((NNTrainDataStorage*)st); }
REF_API void ref_random_set_storage(void* p, void* st) { ((RandomPlayer*)p)->setTrainStorage((NNTrainDataStorage*)st); }
REF_API void ref_script_game_finished(void* p, int status, int rounds) { ((ScriptPlayer*)p)->gameFinished(status, rounds); }
REF_API void ref_random_game_finished(void* p, int status, int rounds) { ((RandomPlayer*)p)->gameFinished(status, rounds); }
REF_API long ref_storage_count(void* st) { return (long)((NNTrainDataStorage*)st)->data.size(); }
REF_API int ref_storage_save(void* st, const char* path)
{
	try { ((NNTrainDataStorage*)st)->saveTrainingSamples(path); return 0; }
	catch (std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

/* Game::newGame's mirror game (game/game.cpp:170-179): previous start state with the sides swapped */
REF_API void ref_state_invert_players(void* s) { ((State*)s)->invertPlayers(); }
REF_API void ref_state_set_current_player(void* s, int p) { ((State*)s)->setCurrentPlayerTurn(p); }

/* the reference's own sample store: push n samples (player, NNInputData image, policy), let updateValues(status) fill the values,
   saveTrainingSamples(path).  alphazero_trainer.cpp:108-114, alphazero_nn_data.cpp:51-65, 115-138 */
REF_API int ref_save_samples(const char* path, int n, const int8_t* players, const uint8_t* nn_inputs88, const float* policies43, int status, int rounds)
{
	try {
		NNTrainDataStorage st;
		for (int i = 0; i < n; i++) {
			NNInputData in;
			memcpy((void*)&in, nn_inputs88 + (size_t)i * sizeof(NNInputData), sizeof in);
			std::vector<float> pol(policies43 + (size_t)i * TF_OUTPUT_POLICY_TENSOR_SIZE, policies43 + (size_t)(i + 1) * TF_OUTPUT_POLICY_TENSOR_SIZE);
			st.data.push_back(NNTrainData((uint8_t)players[i], std::move(in), NNOutputData(std::move(pol))));
		}
		st.updateValues(status, rounds);
		st.saveTrainingSamples(path);
		return 0;
	} catch (std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

REF_API void ref_normalize_policy(float* policy43, uint64_t valid)
{
	NNOutputData o;
	o.policy.assign(policy43, policy43 + TF_OUTPUT_POLICY_TENSOR_SIZE);
	o.normalize(valid);
	memcpy(policy43, o.policy.data(), sizeof(float) * TF_OUTPUT_POLICY_TENSOR_SIZE);
}

REF_API int ref_consistency_violations(void* s)
{
	/* consistencyCheck only prints; recompute the invariant it states (state.cpp:1209-1429):
	   all masks and totalArmy are pure functions of landArmy[] */
	State* st = (State*)s;
	const Data& d = st->data;
	int bad = 0;
	for (int p = 0; p < 2; p++)
	{
		uint64_t owned = 0, withArmy = 0, full = 0; int total = 0;
		for (int i = 0; i < LAND_INDEX_SIZE; i++)
			if (d.landArmy[i].playerIndex == p)
			{
				owned |= 1ull << i; total += d.landArmy[i].army;
				if (d.landArmy[i].army > 1) withArmy |= 1ull << i;
				if (d.landArmy[i].army == LAND_ARMY_MAX) full |= 1ull << i;
			}
		uint64_t att = 0, attArmy = 0;
		for (int i = 0; i < LAND_INDEX_SIZE; i++)
		{
			if (owned >> i & 1) att |= Land::getLand((uint8_t)i)->neighboursLandIndexBitMask;
			if (withArmy >> i & 1) attArmy |= Land::getLand((uint8_t)i)->neighboursLandIndexBitMask;
		}
		att &= ~owned; attArmy &= ~owned;
		const PlayerStatus& ps = d.playerStatus[p];
		bad += ps.ownedLands != owned; bad += ps.ownedLandsWithArmy != withArmy; bad += ps.ownedFullLands != full;
		bad += ps.attackLands != att; bad += ps.attackLandsWithArmy != attArmy; bad += ps.totalArmy != total;
	}
	return bad;
}

/* ---------------------------------------------------------------- evaluators */
static int phase_of(const NNInputData* in)
{
	if (in->featureIsPhaseSetup > 0.5f) return 0;
	if (in->featureIsPhaseSetupNeutral > 0.5f) return 1;
	if (in->featureIsPhaseReinforcement > 0.5f) return 2;
	if (in->featureIsPhaseAttack > 0.5f) return 3;
	if (in->featureIsPhaseAttackMobilization > 0.5f) return 4;
	return 5;
}

REF_API void ref_eval_pseudo(const NNInputData* in, float* policy43, float* value, void*)
{
	uint64_t key = az_pn_key((const uint8_t*)in->land, in->playerIndex, in->round, phase_of(in));
	for (int i = 0; i < 43; i++) policy43[i] = az_pn_policy(key, i);
	*value = az_pn_value(key);
}

REF_API void ref_eval_uniform(const NNInputData*, float* policy43, float* value, void*)
{
	for (int i = 0; i < 43; i++) policy43[i] = 1.0f / 43.0f;
	*value = 0.0f;
}

/* ---------------------------------------------------------------- MCTS */
struct RefMcts
{
	AlphaZeroMCTS mcts;
	std::shared_ptr<AlphaZeroNNId> nn;
};

REF_API void* ref_mcts_new(RefEvalFn fn, void* user)
{
	RefMcts* m = new RefMcts();
	m->nn = std::make_shared<AlphaZeroNNId>(fn ? fn : ref_eval_pseudo, user);
	return m;
}
REF_API void ref_mcts_free(void* m) { delete (RefMcts*)m; }
REF_API void ref_mcts_clear(void* m) { ((RefMcts*)m)->mcts.getStorage()->clearNodes(); }   /* AlphaZeroPlayer::newGame, alphazero_player.cpp:31-34 */
REF_API void ref_mcts_trim(void* m) { ((RefMcts*)m)->mcts.getStorage()->trimNodes(); }     /* AlphaZeroPlayer::takeTurn first line, alphazero_player.cpp:5 */
REF_API int ref_mcts_table_size(void* m) { return (int)((RefMcts*)m)->mcts.store.state_map.size(); }
REF_API uint64_t ref_mcts_evals(void* m) { return ((RefMcts*)m)->nn->evals; }

static void root_stats(RefMcts* h, State& root, uint32_t* N43, float* Q43, float* P43, float* pi43, uint32_t* sumN, float* root_value)
{
	std::shared_ptr<StateSimulations> ss = h->mcts.getStorage()->getStateSimulation(root);
	std::vector<float> policy = ss->calculateMoveProbability(1.0f);
	for (int i = 0; i < ALL_MOVES; i++)
	{
		LandIndex li = Utility::i2li((uint8_t)i);
		if (ss->moveValues.contains(li))
		{
			const SimulationValue& sv = ss->moveValues.at(li);
			N43[i] = sv.N; Q43[i] = sv.Q; P43[i] = sv.P;
		}
		else { N43[i] = 0; Q43[i] = 0.0f; P43[i] = 0.0f; }
		pi43[i] = policy[i];
	}
	*sumN = ss->sumN;
	*root_value = ss->value;
}

/*
 * One AlphaZeroMCTS::simulate (alphazero_mcts.cpp:255-287) with THREADS_PER_MCTS = 1
 * semantics, driven from the calling thread so that every simulation can be given
 * its own (game, ply, sim) dice stream: setRootState, then `sims - sims % T` times
 * search(copy of root).  Outputs the root statistics (zeros for illegal moves).
 */
REF_API int ref_mcts_search(void* m, void* s, uint64_t seed, uint32_t game, uint32_t ply,
                            uint32_t* N43, float* Q43, float* P43, float* pi43, uint32_t* sumN, float* root_value)
{
	try
	{
		RefMcts* h = (RefMcts*)m;
		State& root = *(State*)s;
		rng_philox(seed, game, ply, 0);
		h->mcts.setRootState(root, h->nn);
		int count = SETTINGS.MCTS_SIMULATIONS - (SETTINGS.MCTS_SIMULATIONS % SETTINGS.THREADS_PER_MCTS);
		for (int i = 0; i < count; i++)
		{
			rng_philox(seed, game, ply, (uint32_t)i);
			State copyState = root;
			copyState.setLog(false);
			h->mcts.search(copyState, h->nn);
		}
		root_stats(h, root, N43, Q43, P43, pi43, sumN, root_value);
		return 0;
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

/*
 * AlphaZeroMCTS::simulate with K REAL search threads (threadSimulateJob, alphazero_mcts.cpp:310-320) running the
 * UNMODIFIED AlphaZeroMCTS::search, forced into ONE of the interleavings the reference's locks allow — the "lockstep"
 * schedule the CUDA search implements:
 *   round r = simulations r*K .. r*K+K-1; thread j runs simulation r*K+j (its dice stream is (game, ply, r*K+j));
 *   selection: thread 0 descends first, thread j+1 starts its descent when thread j has either reached predictFuture
 *     (an unseen state: it parks there, like the reference's thread blocks on the batch queue) or finished (terminal
 *     state: search() has already backed the result up).  Every getNextBestMoveAndSetVisited therefore sees the
 *     active_N marks of the descents before it in the round (the "virtual loss" rule, alphazero_mcts.cpp:91-107);
 *   completion: when all K have selected, the parked threads are released one at a time in thread order: evaluator,
 *     normalize, StateSimulationsStorage::add (drops a state a previous thread of the round has added), addValue up
 *     the path.
 * The threads only take turns; every tree operation is the reference's own code.
 */
struct Lockstep
{
	std::mutex mu; std::condition_variable cv;
	int K = 1, round = 0, sel_turn = 0, fin_turn = -1;
};
static thread_local int tl_slot = -1;
static thread_local bool tl_parked = false;

static void lockstep_selected(Lockstep* L, std::unique_lock<std::mutex>& lk)
{
	L->sel_turn++;
	if (L->sel_turn == L->K) L->fin_turn = 0;
	L->cv.notify_all();
	L->cv.wait(lk, [&] { return L->fin_turn == tl_slot; });
}

static void lockstep_park(void* user)
{
	Lockstep* L = (Lockstep*)user;
	std::unique_lock<std::mutex> lk(L->mu);
	tl_parked = true;
	lockstep_selected(L, lk);
}

REF_API int ref_mcts_search_lockstep(void* m, void* s, uint64_t seed, uint32_t game, uint32_t ply, int K,
                                     uint32_t* N43, float* Q43, float* P43, float* pi43, uint32_t* sumN, float* root_value)
{
	try
	{
		RefMcts* h = (RefMcts*)m;
		State& root = *(State*)s;
		rng_philox(seed, game, ply, 0);
		h->mcts.setRootState(root, h->nn);
		int count = SETTINGS.MCTS_SIMULATIONS - (SETTINGS.MCTS_SIMULATIONS % SETTINGS.THREADS_PER_MCTS);
		if (K < 1 || count % K != 0) { snprintf(g_err, sizeof g_err, "simulation count %d is not a multiple of K = %d", count, K); return -1; }
		const int rounds = count / K;
		Lockstep L; L.K = K;
		h->nn->park = lockstep_park; h->nn->park_user = &L;
		std::atomic<int> failed{ 0 };
		std::string what;                                   /* g_err is thread-local: carry a worker's message back */
		std::vector<std::thread> threads;
		for (int j = 0; j < K; j++)
		{
			threads.push_back(std::thread([&, j] {
				tl_slot = j;
				for (int r = 0; r < rounds; r++)
				{
					{
						std::unique_lock<std::mutex> lk(L.mu);
						L.cv.wait(lk, [&] { return L.round == r && L.sel_turn == j; });
					}
					tl_parked = false;
					rng_philox(seed, game, ply, (uint32_t)(r * K + j));
					State copyState = root;
					copyState.setLog(false);
					try { h->mcts.search(copyState, h->nn); }
					catch (const std::exception& e) { std::lock_guard<std::mutex> g(L.mu); what = e.what(); failed = 1; }
					std::unique_lock<std::mutex> lk(L.mu);
					if (!tl_parked) lockstep_selected(&L, lk);     /* terminal descent: backed up already, keeps its place in the order */
					L.fin_turn++;
					if (L.fin_turn == K) { L.round++; L.sel_turn = 0; L.fin_turn = -1; }
					L.cv.notify_all();
				}
			}));
		}
		for (auto& t : threads) t.join();
		h->nn->park = nullptr; h->nn->park_user = nullptr;
		if (failed) { snprintf(g_err, sizeof g_err, "%s", what.c_str()); return -1; }
		root_stats(h, root, N43, Q43, P43, pi43, sumN, root_value);
		return 0;
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

/* move choice from pi: play mode = argmax (alphazero_player.cpp:12), self-play = sample while
   round <= TEMPERATURE_TRESHOLD (alphazero_trainer.cpp:98-106) with the contract's float */
REF_API int ref_pick_move(void* m, const float* pi43, int sample, uint64_t seed, uint32_t game, uint32_t ply)
{
	try
	{
		RefMcts* h = (RefMcts*)m;
		std::vector<float> probs(pi43, pi43 + ALL_MOVES);
		LandIndex li;
		if (sample) { rng_philox(seed, game, ply, AZ_STREAM_REAL); li = h->mcts.pickRandomWeightedMove(probs); }
		else li = h->mcts.pickHigestWeightedMove(probs);
		return (int)Utility::li2i(li);
	}
	catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

/* ---------------------------------------------------------------- CPU baseline timing (reference's own RNG) */
struct BenchOut { uint64_t steps, games, sims, evals, moves; double seconds; };

/* config 2 on the CPU: every thread loops newGame -> {getValidMoves, uniform-random legal move,
   makeMove, gameStatus}; returns total env steps and wall seconds */
REF_API void ref_bench_env(int n_threads, uint64_t steps_per_thread, uint32_t seed, BenchOut* out)
{
	std::vector<std::thread> th;
	std::vector<BenchOut> res(n_threads);
	auto t0 = std::chrono::steady_clock::now();
	for (int t = 0; t < n_threads; t++)
		th.emplace_back([&, t]() {
			rng_engine(seed + 7919u * (uint32_t)t);
			uint64_t steps = 0, games = 0;
			while (steps < steps_per_thread)
			{
				State st; st.setLog(false); st.newGame(); games++;
				while (st.gameStatus() == State::NOT_ENDED && steps < steps_per_thread)
				{
					uint64_t vm = UtilityNN::getValidMoves(st);
					uint64_t mv = Utility::randomMask(vm);
					UtilityNN::makeMove(st, Utility::lm2li(mv));
					steps++;
				}
			}
			res[t].steps = steps; res[t].games = games;
		});
	for (auto& x : th) x.join();
	auto t1 = std::chrono::steady_clock::now();
	memset(out, 0, sizeof *out);
	for (auto& r : res) { out->steps += r.steps; out->games += r.games; }
	out->seconds = std::chrono::duration<double>(t1 - t0).count();
}

/* configs 1/3/5 on the CPU: one self-play game loop per thread through the reference's own
   AlphaZeroMCTS::simulate (spawns THREADS_PER_MCTS search threads per call, alphazero_mcts.cpp:269-276),
   following threadExecuteTrainingGame (alphazero_trainer.cpp:80-119); fn = evaluator (null net by default) */
REF_API void ref_bench_selfplay(int n_threads, uint64_t moves_per_thread, uint32_t seed, RefEvalFn fn, void* user, BenchOut* out)
{
	std::vector<std::thread> th;
	std::vector<BenchOut> res(n_threads);
	auto t0 = std::chrono::steady_clock::now();
	for (int t = 0; t < n_threads; t++)
		th.emplace_back([&, t]() {
			rng_engine(seed + 7919u * (uint32_t)t);
			std::shared_ptr<AlphaZeroNNId> nn = std::make_shared<AlphaZeroNNId>(fn ? fn : ref_eval_uniform, user);
			uint64_t moves = 0, games = 0, steps = 0;
			int count = SETTINGS.MCTS_SIMULATIONS - (SETTINGS.MCTS_SIMULATIONS % SETTINGS.THREADS_PER_MCTS);
			while (moves < moves_per_thread)
			{
				AlphaZeroMCTS mcts = AlphaZeroMCTS();
				State root; root.setLog(false); root.newGame(); games++;
				int8_t gs = -1;
				while (gs == -1 && moves < moves_per_thread)
				{
					mcts.simulate(root, nn);
					std::shared_ptr<StateSimulations> ss = mcts.getStorage()->getStateSimulation(root);
					std::vector<float> policy = ss->calculateMoveProbability(1.0f);
					LandIndex li = root.getRound() > SETTINGS.TEMPERATURE_TRESHOLD ? mcts.pickHigestWeightedMove(policy) : mcts.pickRandomWeightedMove(policy);
					UtilityNN::makeMove(root, li);
					gs = root.gameStatus();
					moves++; steps++;
				}
			}
			res[t].moves = moves; res[t].games = games; res[t].sims = moves * (uint64_t)count; res[t].evals = nn->evals; res[t].steps = steps;
		});
	for (auto& x : th) x.join();
	auto t1 = std::chrono::steady_clock::now();
	memset(out, 0, sizeof *out);
	for (auto& r : res) { out->moves += r.moves; out->games += r.games; out->sims += r.sims; out->evals += r.evals; out->steps += r.steps; }
	out->seconds = std::chrono::duration<double>(t1 - t0).count();
}
