// Codebook nearest-neighbour search, coarse pass: one layer of  score[n,k] = ||c_k||^2 - 2 r_n.c_k  as an sm_100a
// tcgen05 GEMM (fp16 operands, fp32 accumulators in TMEM) with the row top-6 kept in the epilogue.
//
// Replaces the `torch.cdist` + `argmin` pair of VectorQuantizer.forward (nat.py:2146, 2157).  The [N,K] distance
// matrix is never materialised: each epilogue thread owns one frame (one TMEM lane) and folds its 256 fresh
// accumulators per chunk into six packed (score|index) keys.  The exact decision is taken afterwards by
// rvq_rows.cuh from these candidates under a proven error bound (DESIGN.md "Exactness").
//
// Shape of the kernel (one CTA per SM, persistent over 128-frame tiles):
//   warp 0      TMA producer: A tile [128 frames x 64] and B tile [256 codes x 64] per K-block, 4-stage ring
//   warp 1      TMEM owner + single-thread tcgen05.mma issuer, two 256-column accumulator stages
//   warps 2..5  epilogue: tcgen05.ld 32 columns at a time, score, key, branch-free top-6 insertion
#pragma once

#include "nat_common.cuh"

namespace nat {
namespace gemm {

constexpr int BLOCK_M = 128;     // frames per tile == TMEM lanes
constexpr int BLOCK_N = 256;     // codes per accumulator stage
constexpr int BLOCK_K = 64;      // fp16 elements per K-block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 2 * BLOCK_N;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 /*barriers*/ + 1024 /*alignment slack*/;
constexpr int KEY_INVALID = 0x7FFFFFFF;

// Candidates kept per frame. With the fp16 window about 1/6 of the score spacing at the minimum (randn data), the
// chance that the last kept candidate is still inside the window falls ~6x per extra candidate: 4 -> 1e-3 of the
// frames need the exact full scan, 6 -> 1e-6. Each candidate costs two integer min/max per score in the epilogue.
#ifndef NAT_NCAND
#define NAT_NCAND 6
#endif
constexpr int NCAND = NAT_NCAND;
#ifndef NAT_GEMM_MINBLOCKS
#define NAT_GEMM_MINBLOCKS 1
#endif

// Candidates of one frame after the coarse pass: keys ascending; key = (score bits & ~0xFF) | column-in-chunk,
// idx = global code index (16 bit). The alignment pads the struct to a multiple of 16 bytes (48 for six).
struct __align__(16) Cand {
    int key[NCAND];
    unsigned short idx[NCAND];
};
static_assert(sizeof(Cand) % 16 == 0, "Cand layout");

// Branch-free insertion of v into the ascending list m[0..NCAND): 2*NCAND-1 integer min/max.
__device__ __forceinline__ void topk_insert(int v, int (&m)[NCAND]) {
#pragma unroll
    for (int i = 0; i < NCAND - 1; ++i) {
        const int a = min(m[i], v);
        v = max(m[i], v);
        m[i] = a;
    }
    m[NCAND - 1] = min(m[NCAND - 1], v);
}

// DUMP=true writes the raw accumulators instead of candidates: validation of the MMA path, and the score matrix of the
// Philox sampling mode (rows::sample_from_acc_kernel).
// NAT_GEMM_MINBLOCKS = 2 would only cap registers at 168 per thread (shared memory still admits one CTA per SM) so
// that row kernels of another stream can co-reside; measured slower on B200, default 1.
template <bool DUMP>
__global__ void __launch_bounds__(NUM_THREADS, NAT_GEMM_MINBLOCKS)
rvq_gemm_topk_kernel(const __grid_constant__ CUtensorMap map_a,   // fp16 [rows, Dp], box 64 x 128, SWIZZLE_128B
                     const __grid_constant__ CUtensorMap map_b,   // fp16 [L*Kp, Dp], box 64 x 256, SWIZZLE_128B
                     int n_rows, int n_tiles, int n_chunks, int n_kblocks, int b_row0,
                     int n_splits,                                // DUMP only: a tile's chunks are dealt to n_splits CTAs
                     const float4* __restrict__ rowinfo,          // per frame {alpha, bias, window, -}
                     const float* __restrict__ cn,                // [Kp] ||c_k||^2 of this layer, +inf in the padding
                     Cand* __restrict__ cand, float* __restrict__ dump, int dump_ld) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* full = bars;                  // TMA -> MMA
    uint64_t* empty = bars + STAGES;        // MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;    // MMA -> epilogue (accumulator stage ready)
    uint64_t* tempty = tfull + 2;           // epilogue -> MMA (accumulator stage drained)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (!DUMP) n_splits = 1;                                      // candidate lists are merged across a tile's chunks
    n_splits = max(1, min(n_splits, n_chunks));
    const int cps = (n_chunks + n_splits - 1) / n_splits;        // chunks per work item

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_mbar_init();
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int w = blockIdx.x; w < n_tiles * n_splits; w += gridDim.x) {
                const int tile = w / n_splits, c0 = (w % n_splits) * cps, c1 = min(n_chunks, c0 + cps);
                for (int chunk = c0; chunk < c1; ++chunk) {
                    for (int kb = 0; kb < n_kblocks; ++kb) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full[s], A_STAGE_BYTES + B_STAGE_BYTES);
                        tma_load_2d(smem_a + s * A_STAGE_BYTES, &map_a, &full[s], kb * BLOCK_K, tile * BLOCK_M);
                        tma_load_2d(smem_b + s * B_STAGE_BYTES, &map_b, &full[s], kb * BLOCK_K,
                                    b_row0 + chunk * BLOCK_N);
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(BLOCK_M, BLOCK_N);
            uint32_t s = 0, ph = 0, it = 0;
            for (int w = blockIdx.x; w < n_tiles * n_splits; w += gridDim.x) {
                const int c0 = (w % n_splits) * cps, c1 = min(n_chunks, c0 + cps);
                for (int chunk = c0; chunk < c1; ++chunk, ++it) {
                    const uint32_t as = it & 1, aph = (it >> 1) & 1;
                    mbar_wait(&tempty[as], aph ^ 1);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                    for (int kb = 0; kb < n_kblocks; ++kb) {
                        mbar_wait(&full[s], ph);
                        tcgen05_fence_after();
                        const uint64_t adesc = umma_desc_kmajor_sw128(smem_a + s * A_STAGE_BYTES);
                        const uint64_t bdesc = umma_desc_kmajor_sw128(smem_b + s * B_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            // +32 bytes per UMMA_K inside the 128-byte swizzle row: start-address field += 2
                            umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(&empty[s]);
                        if (kb == n_kblocks - 1) umma_commit(&tfull[as]);
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: 4 warps, one frame per thread
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int row_in_tile = q * 32 + lane;
        uint32_t it = 0;
        for (int w = blockIdx.x; w < n_tiles * n_splits; w += gridDim.x) {
            const int tile = w / n_splits, c0 = (w % n_splits) * cps, c1 = min(n_chunks, c0 + cps);
            const long long row = static_cast<long long>(tile) * BLOCK_M + row_in_tile;
            float alpha = 0.f, bias = 0.f;
            if (row < n_rows) {
                const float4 ri = __ldg(&rowinfo[row]);
                alpha = ri.x;
                bias = ri.y;
            }
            int gk[NCAND], gi[NCAND];
#pragma unroll
            for (int i = 0; i < NCAND; ++i) { gk[i] = KEY_INVALID; gi[i] = 0; }
            for (int chunk = c0; chunk < c1; ++chunk, ++it) {
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                mbar_wait(&tfull[as], aph);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
                int lk[NCAND];
#pragma unroll
                for (int i = 0; i < NCAND; ++i) lk[i] = KEY_INVALID;
                const float4* cn4 = reinterpret_cast<const float4*>(cn + chunk * BLOCK_N);
#pragma unroll
                for (int g = 0; g < BLOCK_N / 32; ++g) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + g * 32, v);
                    tmem_wait_ld();
                    if (DUMP) {
                        if (row < n_rows) {
                            // one full 128-byte line per thread, as four 32-byte stores: whole sectors (dump_ld is a
                            // multiple of 256; two 16-byte halves of a sector from two instructions cost L2 a merge)
                            float* out = dump + static_cast<long long>(row) * dump_ld + chunk * BLOCK_N + g * 32;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                             ::"l"(out + 8 * j), "r"(v[8 * j]), "r"(v[8 * j + 1]), "r"(v[8 * j + 2]), "r"(v[8 * j + 3]),
                                               "r"(v[8 * j + 4]), "r"(v[8 * j + 5]), "r"(v[8 * j + 6]), "r"(v[8 * j + 7]) : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 c = __ldg(cn4 + g * 8 + j4);
                            const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int col = g * 32 + j4 * 4 + e;
                                const float s = fmaf(__uint_as_float(v[j4 * 4 + e]), alpha, cc[e]) + bias;
                                const int key = (__float_as_int(s) & 0xFFFFFF00) | col;
                                topk_insert(key, lk);
                            }
                        }
                    }
                }
                // accumulator stage drained: hand it back to the MMA warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);

                if (!DUMP) {   // merge the chunk's keys into the running global list (stable: earlier chunk wins ties)
#pragma unroll
                    for (int i = 0; i < NCAND; ++i) {
                        int key = lk[i];
                        int idx = chunk * BLOCK_N + (key & 0xFF);
#pragma unroll
                        for (int p = 0; p < NCAND; ++p) {
                            const bool lt = key < gk[p];
                            const int tk = gk[p], ti = gi[p];
                            gk[p] = lt ? key : tk;
                            gi[p] = lt ? idx : ti;
                            key = lt ? tk : key;
                            idx = lt ? ti : idx;
                        }
                    }
                }
            }
            if (!DUMP && row < n_rows) {
                Cand c;
#pragma unroll
                for (int i = 0; i < NCAND; ++i) { c.key[i] = gk[i]; c.idx[i] = static_cast<unsigned short>(gi[i]); }
                cand[row] = c;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace gemm
}  // namespace nat
