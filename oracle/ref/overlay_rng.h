#pragma once
/*
 * TEST INFRASTRUCTURE — overlay for /root/reference/src/rng.h, used only when the
 * UNMODIFIED reference sources are compiled into oracle/_ref/ by oracle/ref/build_ref.sh.
 *
 * It keeps the reference's Rng API (rInt / rDice / rFloat / getInstance / getEngine,
 * rng.h:22-47) but serves the numbers from a per-thread context that the C shim
 * (oracle/ref/ref_shim.cpp) sets up, so a run of the reference is reproducible and
 * can be compared bit-for-bit with the CUDA path:
 *
 *   REF_RNG_PHILOX : the (seed, game, ply, sim) counter-based contract of include/az_philox.h
 *   REF_RNG_TAPE   : explicit integer tape (golden vectors with hand-picked dice)
 *   REF_RNG_ENGINE : the reference's own behaviour (std::default_random_engine +
 *                    uniform distributions), thread-local, used for CPU-baseline timing
 */
#include <random>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <atomic>

#include "az_philox.h"

enum RefRngMode { REF_RNG_ENGINE = 0, REF_RNG_PHILOX = 1, REF_RNG_TAPE = 2 };

inline uint32_t nextEngineSeed()
{
	/* every thread (incl. the search threads AlphaZeroMCTS::simulate spawns) gets its own stream */
	static std::atomic<uint32_t> n{ 0 };
	return 12345u + 7919u * n.fetch_add(1);
}

struct RefRngCtx
{
	int mode = REF_RNG_ENGINE;
	/* philox contract */
	uint64_t seed = 0;
	uint32_t game = 0, ply = 0, sim = AZ_STREAM_REAL;
	uint32_t die_j = 0;   /* dice drawn so far in this (game, ply, sim) stream */
	uint32_t deal_i = 0;  /* rInt() draws so far in this deal                  */
	uint32_t int_j = 0;   /* rInt() draws so far in this scripted-opponent turn */
	/* tape */
	const int32_t* tape = nullptr;
	size_t tape_len = 0, tape_pos = 0;
	/* counters (diagnostics) */
	uint64_t n_dice = 0, n_int = 0, n_float = 0;
	/* engine */
	std::default_random_engine engine{ nextEngineSeed() };
	std::uniform_int_distribution<int> dist_int{ 0, RAND_MAX };
	std::uniform_int_distribution<int> dist_dice{ 1, 6 };
	std::uniform_real_distribution<float> dist_float{ 0.0f, 1.0f };
};

inline RefRngCtx& ref_rng_ctx()
{
	static thread_local RefRngCtx ctx;
	return ctx;
}

class Rng
{
	Rng() {}

	static int popTape(RefRngCtx& c)
	{
		if (c.tape_pos >= c.tape_len)
		{
			fprintf(stderr, "ref overlay rng: tape exhausted at %zu\n", c.tape_pos);
			abort();
		}
		return c.tape[c.tape_pos++];
	}
public:
	int rInt()
	{
		RefRngCtx& c = ref_rng_ctx();
		c.n_int++;
		switch (c.mode)
		{
		case REF_RNG_PHILOX:
		{
			/* only the initial deal draws ints on the hot path: Utility::randomMask,
			   land.cpp:100-112 computes rInt() % remaining; remaining = 42 - i */
			if (c.sim == AZ_STREAM_OPP) return (int)az_rng_opp_int(c.seed, c.game, c.ply, c.int_j++);
			if (c.deal_i >= 42) { fprintf(stderr, "ref overlay rng: rInt outside a deal\n"); abort(); }
			uint32_t k = az_rng_deal_draw(c.seed, c.game, c.ply, c.deal_i);
			c.deal_i++;
			return (int)k;
		}
		case REF_RNG_TAPE: return popTape(c);
		default: return c.dist_int(c.engine);
		}
	}

	int rDice()
	{
		RefRngCtx& c = ref_rng_ctx();
		c.n_dice++;
		switch (c.mode)
		{
		case REF_RNG_PHILOX: return az_rng_die(c.seed, c.game, c.ply, c.sim, c.die_j++);
		case REF_RNG_TAPE: return popTape(c);
		default: return c.dist_dice(c.engine);
		}
	}

	float rFloat()
	{
		RefRngCtx& c = ref_rng_ctx();
		c.n_float++;
		switch (c.mode)
		{
		case REF_RNG_PHILOX:
		{
			if (c.sim == AZ_STREAM_OPP) return az_rng_unit_float(az_rng_opp_word(c.seed, c.game, c.ply, c.int_j++));
			az_u32x4 b = az_rng_block(c.seed, c.game, c.ply, AZ_STREAM_REAL, 0);
			return az_rng_unit_float(b.z);
		}
		case REF_RNG_TAPE: return (float)popTape(c) * (1.0f / 16777216.0f);
		default: return c.dist_float(c.engine);
		}
	}

	static Rng& getInstance()
	{
		static Rng INSTANCE;
		return INSTANCE;
	}

	std::default_random_engine& getEngine()
	{
		return ref_rng_ctx().engine;
	}
};

static Rng& RNG = Rng::getInstance();
