/*
 * az_philox.h — the random-number CONTRACT shared by the CUDA kernels, the CPU
 * oracle (oracle/risk_oracle.c) and the overlay that drives the compiled
 * reference (oracle/ref/overlay_rng.h).
 *
 * The reference draws every random number from one process-global
 * std::default_random_engine seeded from std::random_device
 * (/root/reference/src/rng.h:5-50), so its stream is not reproducible.  Parity is
 * therefore defined on a counter-based stream that both sides evaluate:
 *
 *     block(seed, game, ply, sim, b) = Philox4x32-10(key = {seed_lo, seed_hi},
 *                                                    ctr = {game, ply, sim, b})
 *
 *   game : global game-slot id (independent of how games are sharded over GPUs)
 *   ply  : number of real moves this slot has made so far (never reset on re-deal)
 *   sim  : MCTS simulation index inside the search for that move, or one of the
 *          reserved stream ids below for things that are not a simulation
 *   b    : block index inside the stream
 *
 * Dice (reference: State::getDiceRolls, state/state.cpp:645-684, one
 * RNG.rDice() per die, attacker dice first then defender dice, state.cpp:832-833)
 * are consumed SEQUENTIALLY inside one (game, ply, sim) stream, exactly like the
 * reference consumes its global engine: the j-th die drawn is base-6 digit
 * (j % 5) of word ((j / 5) % 4) of block (j / 20).  Digit extraction is the
 * exact multiply-shift chain  d = (w*6)>>32 ; w = (uint32)(w*6) ; value = d+1.
 * A real move needs at most 5 dice => word 0 of block 0.
 *
 * For the real move (sim = AZ_STREAM_REAL) block 0 additionally provides
 *   word 1 : uniform-random legal action selector, k = mulhi(word1, popcount(mask))
 *            (mirrors Utility::randomMask, land/land.cpp:100-112: k-th set bit)
 *   word 2 : self-play move sampling float, (word2 >> 8) * 2^-24
 *            (mirrors RNG.rFloat() in pickRandomWeightedMove, alphazero_mcts.cpp:379-395)
 * The initial deal (State::newGame, state.cpp:137-167: 42 draws of
 * rInt() % remaining) uses sim = AZ_STREAM_DEAL: draw i = mulhi(word(i%4) of
 * block (i/4), 42 - i).
 *
 * A scripted opponent (ScriptPlayer::takeTurn, player/script/script_player.cpp:162-227) plays a WHOLE turn in one call
 * and draws an unbounded number of dice: they are consumed sequentially from the stream sim = AZ_STREAM_OPP of the
 * (game, ply) at which the turn starts (one ply per opponent turn), exactly like the reference consumes its engine.
 * Its rInt() draws (Utility::randomMask in the setup phase, land/land.cpp:100-112: rInt() % count) come from
 * sim = AZ_STREAM_OPP_INT: draw i = word (i % 4) of block (i / 4), shifted right by one (a non-negative int); RandomPlayer's
 * rFloat() (player/random/random_player.cpp:66) takes the next word of the same sequence as (word >> 8) * 2^-24.
 *
 * Plain C99 / CUDA; no dependencies.
 */
#ifndef AZ_PHILOX_H
#define AZ_PHILOX_H

#include <stdint.h>

#if defined(__CUDACC__)
#define AZ_HD __host__ __device__ __forceinline__
#else
#define AZ_HD static inline
#endif

#define AZ_STREAM_REAL 0xFFFFFFFFu /* dice / action / sampling float of a real move */
#define AZ_STREAM_DEAL 0xFFFFFFFEu /* initial deal of a (re)started game             */
#define AZ_STREAM_OPP 0xFFFFFFFDu  /* dice of one scripted-opponent turn               */
#define AZ_STREAM_OPP_INT 0xFFFFFFFCu /* rInt() draws of one scripted-opponent turn   */

typedef struct az_u32x4 { uint32_t x, y, z, w; } az_u32x4;

AZ_HD uint32_t az_mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* Philox4x32-10 (Salmon et al., SC'11): multipliers 0xD2511F53 / 0xCD9E8D57,
 * Weyl key increments 0x9E3779B9 / 0xBB67AE85. */
AZ_HD az_u32x4 az_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                uint32_t k0, uint32_t k1)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = az_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = az_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    az_u32x4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

AZ_HD az_u32x4 az_rng_block(uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim, uint32_t b)
{
    return az_philox4x32_10(game, ply, sim, b, (uint32_t)seed, (uint32_t)(seed >> 32));
}

AZ_HD uint32_t az_u32x4_word(az_u32x4 v, int i)
{
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

/* digit-th base-6 digit (0-based) of the 32-bit fraction w/2^32, returned as a die 1..6 */
AZ_HD int az_die_from_word(uint32_t w, int digit)
{
    uint32_t d = 0;
    for (int i = 0; i <= digit; ++i) {
        uint64_t p = (uint64_t)w * 6u;
        d = (uint32_t)(p >> 32);
        w = (uint32_t)p;
    }
    return (int)d + 1;
}

/* j-th die of the stream (seed, game, ply, sim) — reference order of consumption */
AZ_HD int az_rng_die(uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim, uint32_t j)
{
    az_u32x4 blk = az_rng_block(seed, game, ply, sim, j / 20u);
    return az_die_from_word(az_u32x4_word(blk, (int)((j / 5u) & 3u)), (int)(j % 5u));
}

/* i-th draw (0..41) of the initial deal: index into the remaining-lands set */
AZ_HD uint32_t az_rng_deal_draw(uint64_t seed, uint32_t game, uint32_t ply, uint32_t i)
{
    az_u32x4 blk = az_rng_block(seed, game, ply, AZ_STREAM_DEAL, i >> 2);
    return az_mulhi32(az_u32x4_word(blk, (int)(i & 3u)), 42u - i);
}

/* i-th non-dice draw of a scripted opponent's turn: rInt() = word >> 1, rFloat() = az_rng_unit_float(word); both advance i */
AZ_HD uint32_t az_rng_opp_word(uint64_t seed, uint32_t game, uint32_t ply, uint32_t i)
{
    az_u32x4 blk = az_rng_block(seed, game, ply, AZ_STREAM_OPP_INT, i >> 2);
    return az_u32x4_word(blk, (int)(i & 3u));
}
AZ_HD uint32_t az_rng_opp_int(uint64_t seed, uint32_t game, uint32_t ply, uint32_t i)
{
    return az_rng_opp_word(seed, game, ply, i) >> 1;
}

AZ_HD float az_rng_unit_float(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

#endif /* AZ_PHILOX_H */
