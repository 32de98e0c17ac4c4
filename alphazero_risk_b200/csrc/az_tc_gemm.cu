// az_tc_gemm.cu — bf16 tcgen05 GEMM for the training step's three convolution-shaped contractions.
//
//   C[split][m][n] = sum over the split's k of A[m][k] * B[n][k]          (fp32 accumulate in TMEM, fp32 out)
//
// profiles/README.md (round 1, third pass): 92 % of a training step is the forward convolution, its data gradient and its weight
// gradient at 26-28 TFLOP/s on the fp32 pipes.  All three are GEMMs once the 3x3 taps are unrolled into K (forward, data gradient:
// M = board cells, K = 9 x channels) or into M (weight gradient: M = 9 x input channels, K = board cells):
//   forward        z[r][co]   = sum_{t,ci} a[nb(r,t)][ci] * w[t][ci][co]            A = im2col(a)        B = w^T
//   data gradient  da[r][ci]  = sum_{t,co} dz[nb(r,t)][co] * w[8-t][ci][co]         A = im2col(dz)       B = flipped w
//   weight grad.   dw[t][ci][co] = sum_r a[nb(r,t)][ci] * dz[r][co]                 A = im2col(a)^T      B = dz^T    (split over r)
// Operands live in global memory in the "chunked K-major" order the UMMA SWIZZLE_NONE descriptor reads: [K/8][rows][8 bf16], so a
// stage of K = 64 is 8 contiguous slabs per operand (bulk copies, no tensor maps) and lands in shared memory already in core-matrix
// order (8 rows x 16 bytes contiguous; leading-dimension offset = rows x 16 B between K chunks, stride offset = 128 B between 8-row
// groups) — the same descriptors the inference tower uses.
//
// Kernel: one CTA per 128 x 256 output tile and K split, cta_group::1, M = 128, N = 256, K = 16 per instruction, 4-stage ring of
// 48 KB stages; warp 0 streams, warp 1 issues, warps 2-5 drain the 256 TMEM columns (32 per tcgen05.ld) to fp32 rows.
// The producer kernels (k_tg_*) build the operands from the fp32 activations / gradients / weights of az_nn_train.cu.
#include <algorithm>
#include <vector>

#include <cuda_bf16.h>

#include "az_common.cuh"
#include "az_tc_ptx.cuh"
#include "az_tc_gemm.cuh"

#define TG_THREADS 192
#define TG_STAGES 4
#define TG_BM 128
#define TG_BN 256
#define TG_BK 64
#define TG_A_STAGE (TG_BK / 8 * TG_BM * 16)       // 16 KB
#define TG_B_STAGE (TG_BK / 8 * TG_BN * 16)       // 32 KB
#define TG_SMEM (TG_STAGES * (TG_A_STAGE + TG_B_STAGE) + 16 * 8 + 16)

// A: [Kp/8][Mp][8] bf16, B: [Kp/8][256][8] bf16; Mp multiple of 128, Kp multiple of 64; C: [splits][Mp][256] fp32.
// kb_per_split K blocks of 64 per split (the last split may get fewer, never zero).
__global__ void __launch_bounds__(TG_THREADS, 1)
k_tc_gemm(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ C, int Mp, int Kp, int kb_per_split)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + TG_STAGES * TG_A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + TG_STAGES * TG_B_STAGE);
    uint64_t* bar_full = bars;                     // [TG_STAGES]
    uint64_t* bar_empty = bars + TG_STAGES;        // [TG_STAGES]
    uint64_t* bar_acc = bars + 2 * TG_STAGES;      // accumulator complete
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TG_BM, split = blockIdx.y;
    const int kb_total = Kp / TG_BK;
    const int kb0 = split * kb_per_split;
    const int nkb = min(kb_per_split, kb_total - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t* a8 = reinterpret_cast<const uint8_t*>(A);
            const uint8_t* b8 = reinterpret_cast<const uint8_t*>(B);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TG_STAGES, k = i / TG_STAGES;
                if (k > 0) mbar_wait(bar_empty + s, (uint32_t)((k - 1) & 1));
                mbar_expect_tx(bar_full + s, TG_A_STAGE + TG_B_STAGE);
                const size_t kc0 = (size_t)(kb0 + i) * (TG_BK / 8);
#pragma unroll
                for (int c = 0; c < TG_BK / 8; ++c) {
                    bulk_g2s(sA + (size_t)s * TG_A_STAGE + (size_t)c * TG_BM * 16, a8 + ((kc0 + c) * (size_t)Mp + (size_t)m0) * 16, TG_BM * 16, bar_full + s);
                    bulk_g2s(sB + (size_t)s * TG_B_STAGE + (size_t)c * TG_BN * 16, b8 + (kc0 + c) * (size_t)TG_BN * 16, TG_BN * 16, bar_full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TG_STAGES, k = i / TG_STAGES;
                mbar_wait(bar_full + s, (uint32_t)(k & 1));
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < TG_BK / 16; ++kk) {
                    const uint64_t adesc = umma_desc(a_base + (uint32_t)(s * TG_A_STAGE + kk * 2 * TG_BM * 16), TG_BM * 16, 128);
                    const uint64_t bdesc = umma_desc(b_base + (uint32_t)(s * TG_B_STAGE + kk * 2 * TG_BN * 16), TG_BN * 16, 128);
                    tc_mma_bf16(tmem_base, adesc, bdesc, TC_IDESC, (i > 0 || kk > 0) ? 1u : 0u);
                }
                tc_commit(bar_empty + s);
            }
            tc_commit(bar_acc);
        }
    } else {
        const int q = warp & 3;                    // TMEM lanes 32q .. 32q+31 (warps 2,3,4,5 -> 2,3,0,1)
        mbar_wait(bar_acc, 0u);
        tc_fence_after();
        float* crow = C + ((size_t)split * Mp + (size_t)(m0 + q * 32 + lane)) * TG_BN;
#pragma unroll 1
        for (int c = 0; c < TG_BN / 32; ++c) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            tc_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e)
                *reinterpret_cast<float4*>(crow + c * 32 + e * 4) = make_float4(__uint_as_float(v[e * 4]), __uint_as_float(v[e * 4 + 1]),
                                                                               __uint_as_float(v[e * 4 + 2]), __uint_as_float(v[e * 4 + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---------------------------------------------------------------- operand producers
__constant__ int8_t c_gnb[42 * 9];          // neighbour of board cell p for tap t, -1 outside the board

__device__ __forceinline__ uint4 pack8(const float (&f)[8])
{
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    return o;
}

// fp32 row-major [rows][cin] -> bf16 chunked [cpad/8][rows][8] (channels >= cin are zero).  One block = 32 rows x up to 32 chunks
// through shared memory, so both the global reads (a row's channels) and the global writes (a chunk's rows) are contiguous.
__global__ void __launch_bounds__(256) k_tg_chunk(const float* __restrict__ src, int rows, int cin, int cpad, __nv_bfloat16* __restrict__ out)
{
    __shared__ uint4 tile[32][33];
    const int chunks = cpad / 8, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.x; j < 1024; j += 256) {
        const int rl = j >> 5, cc = c0 + (j & 31), r = r0 + rl;
        if (r < rows && cc < chunks) {
            float f[8];
            const float* s = src + (size_t)r * cin + cc * 8;
            if ((cin & 7) == 0) {
                const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = cc * 8 + e < cin ? s[e] : 0.0f;
            }
            tile[rl][j & 31] = pack8(f);
        }
    }
    __syncthreads();
    uint4* o = reinterpret_cast<uint4*>(out);
    for (int j = threadIdx.x; j < 1024; j += 256) {
        const int cl = j >> 5, rl = j & 31, cc = c0 + cl, r = r0 + rl;
        if (r < rows && cc < chunks) o[(size_t)cc * rows + r] = tile[rl][cl];
    }
}

// im2col, chunked K-major: out[kc][r][8], k = t * cpad + ci (cpad = channels padded to a multiple of 8), value = src[nb(r,t)][ci] or 0,
// from the bf16 chunked copy of the source (k_tg_chunk): one 16-byte load and one 16-byte store per thread, both contiguous along r.
// Rows >= rows (up to Mp) and k >= 9 * cpad (up to Kp) are zero.  grid = (Mp / 256, Kp / 8): block = 256 rows of one K chunk.
__global__ void __launch_bounds__(256) k_tg_im2col(const __nv_bfloat16* __restrict__ src16, int rows, int cpad, int Mp, int Kp, __nv_bfloat16* __restrict__ out)
{
    const int r = blockIdx.x * 256 + threadIdx.x, kc = blockIdx.y;
    if (r >= Mp) return;
    const int chunks = cpad / 8, t = kc / chunks, cc = kc - t * chunks;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows && t < 9) {
        const int b = r / 42, p = r - b * 42, y = p / 6, x = p - y * 6;
        const int dy = t / 3 - 1, dx = t - (t / 3) * 3 - 1;
        if ((unsigned)(y + dy) < 7u && (unsigned)(x + dx) < 6u)
            v = reinterpret_cast<const uint4*>(src16)[(size_t)cc * rows + (size_t)(r + dy * 6 + dx)];
    }
    reinterpret_cast<uint4*>(out)[(size_t)kc * Mp + r] = v;
}

// transposed im2col for the weight gradient: out[rc][m][8], k = board cell row r = rc * 8 + e, m = t * cin + ci (no channel padding:
// m indexes the gradient's rows directly), value = src[nb(r,t)][ci] or 0, from the bf16 chunked copy of the source (k_tg_chunk).
// Thread = (rc, t, channel chunk): eight 16-byte loads (rows r .. r+7 of the chunk, shifted by the tap), an 8 x 8 transpose in
// registers, eight 16-byte stores to consecutive m.  Rows >= rows are zero; m >= 9 * cin (up to Mp) is left alone (those output rows
// are dropped by the reduction).  grid = (ceil(9 * chunks / 96), Kp / 8), block = 96.
__global__ void __launch_bounds__(96) k_tg_im2col_t(const __nv_bfloat16* __restrict__ src16, int rows, int cin, int cpad, int Mp, int Kp, __nv_bfloat16* __restrict__ out)
{
    const int chunks = cpad / 8, tc = blockIdx.x * 96 + threadIdx.x, rc = blockIdx.y;
    if (tc >= 9 * chunks) return;
    const int t = tc / chunks, cc = tc - t * chunks;
    const int dy = t / 3 - 1, dx = t - (t / 3) * 3 - 1;
    const uint4* s4 = reinterpret_cast<const uint4*>(src16) + (size_t)cc * rows;
    uint32_t w[8][4];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int r = rc * 8 + e;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows) {
            const int b = r / 42, p = r - b * 42, y = p / 6, x = p - y * 6;
            if ((unsigned)(y + dy) < 7u && (unsigned)(x + dx) < 6u) v = s4[r + dy * 6 + dx];
        }
        w[e][0] = v.x; w[e][1] = v.y; w[e][2] = v.z; w[e][3] = v.w;
    }
    uint4* o = reinterpret_cast<uint4*>(out) + (size_t)rc * Mp + (size_t)t * cin + cc * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (cc * 8 + j >= cin) break;
        const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
        uint4 v;
        v.x = __byte_perm(w[0][j >> 1], w[1][j >> 1], sel); v.y = __byte_perm(w[2][j >> 1], w[3][j >> 1], sel);
        v.z = __byte_perm(w[4][j >> 1], w[5][j >> 1], sel); v.w = __byte_perm(w[6][j >> 1], w[7][j >> 1], sel);
        o[j] = v;
    }
}

// the same for cin a multiple of 8 (m = tc * 8 + j: a block's 96 x 8 output chunks are one contiguous 12 KB run): the transposed chunks
// go through shared memory (XOR-swizzled against bank conflicts) so that every store instruction writes 512 contiguous bytes per warp
__global__ void __launch_bounds__(96) k_tg_im2col_t8(const __nv_bfloat16* __restrict__ src16, int rows, int chunks, int Mp, int Kp, __nv_bfloat16* __restrict__ out)
{
    __shared__ uint4 stage[96 * 8];
    const int tc = blockIdx.x * 96 + threadIdx.x, rc = blockIdx.y, total = 9 * chunks;
    if (tc < total) {
        const int t = tc / chunks, cc = tc - t * chunks;
        const int dy = t / 3 - 1, dx = t - (t / 3) * 3 - 1;
        const uint4* s4 = reinterpret_cast<const uint4*>(src16) + (size_t)cc * rows;
        uint32_t w[8][4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int r = rc * 8 + e;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r < rows) {
                const int b = r / 42, p = r - b * 42, y = p / 6, x = p - y * 6;
                if ((unsigned)(y + dy) < 7u && (unsigned)(x + dx) < 6u) v = s4[r + dy * 6 + dx];
            }
            w[e][0] = v.x; w[e][1] = v.y; w[e][2] = v.z; w[e][3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
            uint4 v;
            v.x = __byte_perm(w[0][j >> 1], w[1][j >> 1], sel); v.y = __byte_perm(w[2][j >> 1], w[3][j >> 1], sel);
            v.z = __byte_perm(w[4][j >> 1], w[5][j >> 1], sel); v.w = __byte_perm(w[6][j >> 1], w[7][j >> 1], sel);
            stage[threadIdx.x * 8 + (j ^ (threadIdx.x & 7))] = v;
        }
    }
    __syncthreads();
    uint4* o = reinterpret_cast<uint4*>(out) + (size_t)rc * Mp + (size_t)blockIdx.x * 96 * 8;
    const int valid = min(96, total - blockIdx.x * 96) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int q = k * 96 + threadIdx.x;
        if (q < valid) o[q] = stage[(q & ~7) | ((q & 7) ^ ((q >> 3) & 7))];
    }
}

// dz^T for the weight gradient: out[rc][co][8] = dz[rc * 8 + e][co]; thread = (rc, co)
__global__ void __launch_bounds__(256) k_tg_rows_t(const float* __restrict__ dz, int rows, int Kp, __nv_bfloat16* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)(Kp / 8) * 256;
    if (i >= total) return;
    const int rc = (int)(i >> 8), co = (int)(i & 255);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { const int r = rc * 8 + e; f[e] = r < rows ? dz[(size_t)r * 256 + co] : 0.0f; }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(f);
}

// weights as the B operand: out[kc][n][8], k = t * cpad + c.
//   flip = 0 (forward):        n = co, c = ci: value = w[t][ci][co]
//   flip = 1 (data gradient):  n = ci, c = co: value = w[8 - t][ci][co]
// w: fp32 [9][cin][256]; thread = (kc, n)
__global__ void __launch_bounds__(256) k_tg_weights(const float* __restrict__ w, int cin, int cpad, int Kp, int flip, __nv_bfloat16* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)(Kp / 8) * 256;
    if (i >= total) return;
    const int kc = (int)(i >> 8), n = (int)(i & 255);
    const int k0 = kc * 8, t = k0 / cpad, c0 = k0 - t * cpad;
    float f[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    if (t < 9) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c0 + e;
            if (!flip) { if (c < cin) f[e] = w[((size_t)t * cin + c) * 256 + n]; }
            else f[e] = w[((size_t)(8 - t) * cin + n) * 256 + c];          // cin == 256 here
        }
    }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(f);
}

// sum of the K splits: out[m][n] = sum_s part[s][m][n] for m < rows_out (rows of the padded tile beyond are dropped)
__global__ void __launch_bounds__(256) k_tg_reduce(const float* __restrict__ part, int splits, int Mp, int rows_out, float* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows_out * 256) return;
    float s = 0.0f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * Mp * 256 + i];
    out[i] = s;
}

// ---------------------------------------------------------------- host side
static unsigned long long g_tables_devs = 0;       // bit d: device d holds the tap table and the kernel attribute

int az_tg_init()
{
    int dev = 0;
    AZ_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && ((g_tables_devs >> dev) & 1ull)) return AZ_OK;
    int8_t nb[42 * 9];
    for (int p = 0; p < 42; ++p)
        for (int k = 0; k < 9; ++k) {
            int y = p / 6 + k / 3 - 1, x = p % 6 + k % 3 - 1;
            nb[p * 9 + k] = (y < 0 || y >= 7 || x < 0 || x >= 6) ? (int8_t)-1 : (int8_t)(y * 6 + x);
        }
    AZ_CUDA(cudaMemcpyToSymbol(c_gnb, nb, sizeof nb));
    AZ_CUDA(cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
    if (dev < 64) g_tables_devs |= 1ull << dev;
    return AZ_OK;
}

int az_tg_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int Mp, int Kp, int splits, cudaStream_t s)
{
    AZ_REQUIRE(Mp % TG_BM == 0 && Kp % TG_BK == 0 && splits >= 1, "az_tg_gemm: Mp must be a multiple of 128, Kp of 64");
    const int kb_total = Kp / TG_BK;
    if (splits > kb_total) splits = kb_total;
    const int per = (kb_total + splits - 1) / splits;
    const int used = (kb_total + per - 1) / per;                 // every launched split has at least one K block
    AZ_REQUIRE(used == splits, "az_tg_gemm: split count leaves an empty split (use az_tg_splits)");
    k_tc_gemm<<<dim3((unsigned)(Mp / TG_BM), (unsigned)splits), TG_THREADS, TG_SMEM, s>>>(A, B, C, Mp, Kp, per);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

// largest split count <= want for which no split is empty
int az_tg_splits(int Kp, int want)
{
    const int kb_total = Kp / TG_BK;
    if (want > kb_total) want = kb_total;
    if (want < 1) want = 1;
    const int per = (kb_total + want - 1) / want;
    return (kb_total + per - 1) / per;
}

static unsigned blocks_for(size_t total) { return (unsigned)((total + 255) / 256); }

int az_tg_im2col(const float* src, int rows, int cin, int cpad, int Mp, int Kp, __nv_bfloat16* scratch16, __nv_bfloat16* out, cudaStream_t s)
{
    k_tg_chunk<<<dim3((unsigned)((rows + 31) / 32), (unsigned)((cpad / 8 + 31) / 32)), 256, 0, s>>>(src, rows, cin, cpad, scratch16);
    k_tg_im2col<<<dim3((unsigned)((Mp + 255) / 256), (unsigned)(Kp / 8)), 256, 0, s>>>(scratch16, rows, cpad, Mp, Kp, out);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
int az_tg_im2col_t(const float* src, int rows, int cin, int Mp, int Kp, __nv_bfloat16* scratch16, __nv_bfloat16* out, cudaStream_t s)
{
    const int cpad = (cin + 7) / 8 * 8;
    k_tg_chunk<<<dim3((unsigned)((rows + 31) / 32), (unsigned)((cpad / 8 + 31) / 32)), 256, 0, s>>>(src, rows, cin, cpad, scratch16);
    const dim3 grid((unsigned)((9 * (cpad / 8) + 95) / 96), (unsigned)(Kp / 8));
    if (cin % 8 == 0) k_tg_im2col_t8<<<grid, 96, 0, s>>>(scratch16, rows, cpad / 8, Mp, Kp, out);
    else k_tg_im2col_t<<<grid, 96, 0, s>>>(scratch16, rows, cin, cpad, Mp, Kp, out);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
int az_tg_rows_t(const float* dz, int rows, int Kp, __nv_bfloat16* out, cudaStream_t s)
{
    k_tg_rows_t<<<blocks_for((size_t)(Kp / 8) * 256), 256, 0, s>>>(dz, rows, Kp, out);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
int az_tg_weights(const float* w, int cin, int cpad, int Kp, int flip, __nv_bfloat16* out, cudaStream_t s)
{
    k_tg_weights<<<blocks_for((size_t)(Kp / 8) * 256), 256, 0, s>>>(w, cin, cpad, Kp, flip, out);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
int az_tg_reduce(const float* part, int splits, int Mp, int rows_out, float* out, cudaStream_t s)
{
    k_tg_reduce<<<blocks_for((size_t)rows_out * 256), 256, 0, s>>>(part, splits, Mp, rows_out, out);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

// test / bench entry point: C[m][n] = sum_k A[m][k] * B[n][k] for row-major fp32 host matrices (rounded to bf16 on the way in),
// through the chunked layout and the tcgen05 kernel with `splits` K splits; ms = device time of the GEMM launches alone (median of reps)
extern "C" int az_tc_gemm_test(const float* h_a, const float* h_b, int M, int K, int splits, int reps, float* h_c, float* ms)
{
    AZ_REQUIRE(h_a && h_b && h_c && M >= 1 && K >= 1 && reps >= 1, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); az_set_error("no CUDA device: libaz_b200 has no CPU fallback"); return AZ_ERR_NO_DEVICE; }
    int rc = az_tg_init(); if (rc) return rc;
    const int Mp = (M + TG_BM - 1) / TG_BM * TG_BM, Kp = (K + TG_BK - 1) / TG_BK * TG_BK;
    splits = az_tg_splits(Kp, splits);
    std::vector<__nv_bfloat16> a((size_t)(Kp / 8) * Mp * 8, __float2bfloat16(0.0f)), b((size_t)(Kp / 8) * 256 * 8, __float2bfloat16(0.0f));
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) a[((size_t)(k / 8) * Mp + m) * 8 + k % 8] = __float2bfloat16(h_a[(size_t)m * K + k]);
    for (int n = 0; n < 256; ++n) for (int k = 0; k < K; ++k) b[((size_t)(k / 8) * 256 + n) * 8 + k % 8] = __float2bfloat16(h_b[(size_t)n * K + k]);
    __nv_bfloat16 *d_a = nullptr, *d_b = nullptr; float *d_part = nullptr, *d_c = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t ce = cudaMalloc(&d_a, a.size() * 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_b, b.size() * 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_part, sizeof(float) * (size_t)splits * Mp * 256);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_c, sizeof(float) * (size_t)M * 256);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_b, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    std::vector<float> times;
    for (int i = 0; i < reps + 1 && ce == cudaSuccess && rc == AZ_OK; ++i) {
        cudaEventRecord(e0, 0);
        rc = az_tg_gemm(d_a, d_b, d_part, Mp, Kp, splits, 0);
        cudaEventRecord(e1, 0);
        ce = cudaEventSynchronize(e1);
        float t = 0.0f; cudaEventElapsedTime(&t, e0, e1);
        if (i > 0) times.push_back(t);
    }
    if (ce == cudaSuccess && rc == AZ_OK) rc = az_tg_reduce(d_part, splits, Mp, M, d_c, 0);
    if (ce == cudaSuccess && rc == AZ_OK) ce = cudaMemcpy(h_c, d_c, sizeof(float) * (size_t)M * 256, cudaMemcpyDeviceToHost);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_part); cudaFree(d_c);
    if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
    if (ce != cudaSuccess) { az_set_error("az_tc_gemm_test: %s", cudaGetErrorString(ce)); cudaGetLastError(); return AZ_ERR_CUDA; }
    if (rc) return rc;
    if (ms && !times.empty()) { std::sort(times.begin(), times.end()); *ms = times[times.size() / 2]; }
    return AZ_OK;
}
