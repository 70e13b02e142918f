// One persistent sm_100a kernel for a whole RVQ stack: every layer's codebook search (tcgen05 GEMM, fp32
// accumulators in TMEM), the exact decision, the residual update and the next layer's tensor-core operand.
//
// Replaces ResidualVectorQuantizer.forward's layer loop (nat.py:1398-1408) around VectorQuantizer.forward
// (nat.py:2146-2167): cdist + argmin + gather + straight-through update, for all L layers of one stack, in one
// launch. A CTA owns the 128-frame tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  through ALL layers, so the
// only cross-layer dependency (tile t of layer l+1 needs the update of tile t of layer l) stays inside the CTA and
// is a shared-memory counter, not a grid-wide barrier or a kernel boundary.
//
// Order of work inside a CTA: its tiles are taken in groups of `group` tiles (3 by default); inside a group the jobs
// run layer-major  (t0,l0) (t1,l0) (t2,l0) (t0,l1) ...  so the update of one tile runs under the GEMMs of the others.
// The live residual / operand rows of all CTAs together exceed L2, so they travel through HBM between layers; the
// write-back policy (`store_mask`) keeps that traffic to one residual write per four-layer stack.
//
// Two stacks of one tokenizer (S0-S3, A0-A3) that quantise the same frames share ONE layer-0 preparation: layer 0 of a
// tile reads the prepared rows (r0 / its fp16 operand / rowinfo0), which this kernel never writes when the rows it
// writes (r / a / rowinfo / rowamax) are separate buffers, and the second stack's launch is a programmatic dependent
// launch: every CTA releases its dependents at once (the stacks share nothing this kernel writes), so the second
// stack's CTAs take over the SMs one by one as the first stack's CTAs run out of tiles, instead of waiting for its
// slowest CTA. (Both stacks in one launch -- the grid split between them, or (stack, tile) units dealt over the whole
// grid -- was built and measured in round 2: no faster than two launches, and the per-stack indirection cost 3-7 %
// in the update warps; profiles/r2_*.log, DESIGN.md.)
//
// Warp roles (768 threads, launched as clusters of two CTAs: see the kernel's comment):
//   warps 0..3    candidates: tcgen05.ld of the accumulators, coarse score, running-threshold candidate list
//   warps 4..19   update: exact decision (fp64 re-rank where the window demands it), residual update in the
//                 reference's op order, next-layer fp16 operand + error window, index streams
//   warp 20       TMA producer (A tile 128 frames x 64, B tile 256 / PAIR codes x 64 per K-block, 4- or 6-stage ring)
//   warp 21       TMEM owner + single-thread tcgen05.mma issuer (two 256-column accumulator stages)
//   warps 22, 23  idle (they complete the sixth warpgroup that setmaxnreg needs)
//
// Coarse pass and certificate (DESIGN.md "Exactness"): for frame n the kept set is every code whose coarse score
// s_k = acc_k * alpha + ||c_k||^2 lies within the frame's proven window W of the running minimum; the final filter
// against the final minimum leaves {k : s_k <= min_k s + W}, which must contain the true fp32-data argmin. One
// survivor: certified. Several (up to 15 travel in the hand-off record): exact fp64 re-rank. More, or list overflow
// (adversarial orderings only): exact full scan by the update warp. Ties go to the lowest index (nat.py:2157).
#pragma once

#include "nat_common.cuh"
#include "rvq_rows.cuh"

#ifndef NAT_UPD_WARPS
#define NAT_UPD_WARPS 16
#endif
#ifndef NAT_REGS_EPI
#define NAT_REGS_EPI 88
#define NAT_REGS_UPD 88
#endif

namespace nat {
namespace stack {

constexpr int BLOCK_M = 128;     // frames per tile == TMEM lanes
constexpr int BLOCK_N = 256;     // codes per accumulator stage
constexpr int BLOCK_K = 64;      // fp16 elements per K-block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
constexpr int EPI_WARPS = 4;
constexpr int UPD_WARPS = NAT_UPD_WARPS;
constexpr int WARP_EPI0 = 0;
constexpr int WARP_UPD0 = EPI_WARPS;
constexpr int WARP_TMA = EPI_WARPS + UPD_WARPS;
constexpr int WARP_MMA = WARP_TMA + 1;
constexpr int NUM_THREADS = (EPI_WARPS + UPD_WARPS + 4) * 32;  // warpgroups: candidates | update ... | TMA, MMA and two idle warps
constexpr int LAUNCH_REGS = (65536 / NUM_THREADS) / 8 * 8;     // registers per thread the launch bound leaves ptxas
// 80 registers per thread at launch (768 threads); after setmaxnreg: 128 * (88 + 4 * 88 + 40) <= 64 K.
// NAT_UPD_WARPS / NAT_REGS_* are A/B build knobs: 8 update warps with 120 registers measured 12 % slower than 16 with 88
// (the update is bound by rows in flight, not by registers).
#ifndef NAT_REGS_AUX
#define NAT_REGS_AUX 40
#endif
constexpr int REGS_EPI = NAT_REGS_EPI, REGS_UPD = NAT_REGS_UPD, REGS_AUX = NAT_REGS_AUX;
static_assert(WARP_MMA + 1 <= NUM_THREADS / 32 && EPI_WARPS == 4 && UPD_WARPS % 4 == 0 && BLOCK_M % UPD_WARPS == 0 &&
              2 * (BLOCK_M / UPD_WARPS) <= 32, "warpgroup layout");
static_assert(REGS_EPI + (UPD_WARPS / 4) * REGS_UPD + REGS_AUX <= (NUM_THREADS / 128) * LAUNCH_REGS, "setmaxnreg only moves registers inside the CTA: the total must not exceed the launch allocation");
constexpr int TMEM_COLS = 2 * BLOCK_N;
constexpr int ROWS_PER_UPD_WARP = BLOCK_M / UPD_WARPS;
constexpr int LCAP = 16;         // eight-code groups listed per frame (compacted when full and at the end)
constexpr int HCAP = 15;         // candidates handed to the update warps per frame (one 32-byte record: count + 15 indices)
constexpr int HAND_BYTES = 32;
constexpr unsigned HAND_SCAN = 0xFFFFu;

#ifndef NAT_STAGES_PAIR
#define NAT_STAGES_PAIR 6
#endif
// Shared-memory carve-up. A CTA of a pair stages half of every B tile, so the same bytes buy a deeper ring.
template <int PAIR>
struct Smem {
    static constexpr int STAGES = PAIR == 2 ? NAT_STAGES_PAIR : nat::stack::STAGES;
    static constexpr int B_BYTES = B_STAGE_BYTES / PAIR;          // bytes of a B stage held by one CTA
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_STAGE_BYTES;
    static constexpr int OFF_LIST_S = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_LIST_I = OFF_LIST_S + LCAP * BLOCK_M * 4;
    static constexpr int OFF_HAND = OFF_LIST_I + LCAP * BLOCK_M * 4;
    static constexpr int OFF_CN = OFF_HAND + 2 * BLOCK_M * HAND_BYTES;    // [2][BLOCK_N] ||c||^2 of the chunk being scored / the next
    static constexpr int OFF_BARS = OFF_CN + 2 * BLOCK_N * 4;
    static constexpr int BYTES = OFF_BARS + 256 + 1024 /*alignment slack*/;
    static_assert(BYTES <= 227 * 1024, "shared memory budget");
    static_assert((2 * STAGES + 8) * 8 + 4 + UPD_WARPS * 4 <= 256, "barrier block");
};
constexpr int SMEM_BYTES = Smem<1>::BYTES > Smem<2>::BYTES ? Smem<1>::BYTES : Smem<2>::BYTES;

struct StackRef {
    const float* cbf;            // [L, K, dp] fp32 codebooks, zero padded
    const double* cn64;          // [L, K]
    const float* cn32;           // [L, kp], +inf in the padding
    const rows::LayerConst* lc;  // [L]
    const float* r0;             // [rows, dp] rows entering layer 0, written by the preparation kernel
    float* r;                    // [rows, dp] residual rows written back by the update warps (may alias r0: in place)
    __half* a;                   // [rows, dp] fp16 operand rows written by the update warps (layer 0's come from the
                                 //   preparation kernel through its own tensor map; may be the same buffer)
    float4* rowinfo;             // [rows] {alpha, -, window, sx}: layer 0's from the preparation kernel (rowinfo0), then
    const float4* rowinfo0;      //   written per layer by the update warps into `rowinfo`
    float* rowamax;              // [rows] max |r| of the row, ditto (rowamax0 / rowamax)
    const float* rowamax0;
    void* codes;                 // [L, codes_ld] index streams
    long long codes_ld, code_off;
    double* row_loss;            // [L, loss_ld] per-frame sum of t^2, or nullptr
    long long loss_ld;
    unsigned long long* stats;   // [L, 4] counters or nullptr
    int L;
    int store_mask;              // bit l: the residual leaving layer l is written back (hot form only; a clear bit
                                 // means later layers replay that update from the emitted code instead)
};

struct StackArgs {
    StackRef s;
    int n_rows, n_tiles, K, kp, dp, code_dtype;
    int group;                   // tiles per group (>= 1); >= tiles per CTA means plain layer-major order
    int dbg_mode;                // timing experiments only (results invalid when != 0)
    unsigned long long* dbg;     // optional [grid][DBG_SLOTS] cycle counters (nat_debug_stack_counters), or nullptr
};

// Cycle counters per CTA when StackArgs::dbg is set: where each role waits.
enum { DBG_TMA_WAIT_READY = 0, DBG_TMA_WAIT_EMPTY, DBG_MMA_WAIT_TEMPTY, DBG_MMA_WAIT_FULL, DBG_EPI_WAIT_TFULL,
       DBG_EPI_WAIT_CEMPTY, DBG_EPI_TOTAL, DBG_UPD_WAIT_CFULL, DBG_UPD_TOTAL, DBG_KERNEL_TOTAL, DBG_EPI_WAIT_LD, DBG_EPI_EVENTS, DBG_UPD_DECIDE, DBG_UPD_RESID, DBG_UPD_FENCE, DBG_SLOTS = 16 };

// Instrumentation exists only in the DBG instantiation of the kernel: the production build carries no counters,
// no clock reads and no experiment switches (they cost registers in the 40-register roles).
template <bool DBG>
struct WaitClockT {
    long long acc = 0;
    bool on;
    __device__ explicit WaitClockT(bool enabled) : on(enabled) {}
    __device__ __forceinline__ long long begin() const { return on ? clock64() : 0; }
    __device__ __forceinline__ void end(long long t0) { if (on) acc += clock64() - t0; }
};
template <>
struct WaitClockT<false> {
    static constexpr long long acc = 0;
    __device__ explicit WaitClockT(bool) {}
    __device__ __forceinline__ long long begin() const { return 0; }
    __device__ __forceinline__ void end(long long) {}
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st_release(unsigned* smem_counter, unsigned v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(smem_counter)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* smem_counter) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(smem_counter)) : "memory");
    return v;
}
// Bounded like mbar_wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void wait_counter(const unsigned* smem_counter, unsigned need) {
    if (ld_acquire(smem_counter) >= need) return;
    const long long t0 = clock64();
    while (ld_acquire(smem_counter) < need) {
        __nanosleep(32);
        if (clock64() - t0 > 60000000000LL) {
            printf("nat_b200: update counter wait timed out (block %d, need %u, have %u)\n", blockIdx.x, need,
                   ld_acquire(smem_counter));
            __trap();
        }
    }
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// fp64 score ||c||^2 - 2 r.c of one code against the frame held in registers (lane-strided float4s). Same
// accumulation order as rows::exact_score, so both device paths produce identical decisions.
template <int NV>
__device__ __forceinline__ double score_regs(const float4 (&rv)[NV], const float4* __restrict__ c4, int dp4,
                                             double cn64, int lane) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int q = i * 32 + lane;
        if (q < dp4) {
            const float4 b = __ldcg(c4 + q);
            acc = fma(static_cast<double>(rv[i].x), static_cast<double>(b.x), acc);
            acc = fma(static_cast<double>(rv[i].y), static_cast<double>(b.y), acc);
            acc = fma(static_cast<double>(rv[i].z), static_cast<double>(b.z), acc);
            acc = fma(static_cast<double>(rv[i].w), static_cast<double>(b.w), acc);
        }
    }
    acc = warp_sum(acc);
    return cn64 - 2.0 * acc;
}

// fp32 screening of one code: r.c and sum |r_i c_i| (what bounds its rounding error). Two FMA chains per lane and a
// five-level tree: no partial sum is more than 2 NV + 6 <= 22 additions deep.
template <int NV>
__device__ __forceinline__ void score32_regs(const float4 (&rv)[NV], const float4 (&cv)[NV], float& dot, float& adot) {
    float d0 = 0.f, d1 = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        {
            const float4 b = cv[i];                    // zero beyond the row: adds nothing
            d0 = fmaf(rv[i].x, b.x, d0); a0 = fmaf(fabsf(rv[i].x), fabsf(b.x), a0);
            d1 = fmaf(rv[i].y, b.y, d1); a1 = fmaf(fabsf(rv[i].y), fabsf(b.y), a1);
            d0 = fmaf(rv[i].z, b.z, d0); a0 = fmaf(fabsf(rv[i].z), fabsf(b.z), a0);
            d1 = fmaf(rv[i].w, b.w, d1); a1 = fmaf(fabsf(rv[i].w), fabsf(b.w), a1);
        }
    }
    dot = warp_sum(d0 + d1);
    adot = warp_sum(a0 + a1);
}

template <int NV>
__device__ __forceinline__ void load_row(float4 (&rv)[NV], const float4* __restrict__ r4, int dp4, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int q = k * 32 + lane;
        rv[k] = q < dp4 ? __ldcg(r4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int NV>
__device__ __forceinline__ void load_code(float4 (&cv)[NV], const float4* __restrict__ c4, int dp4, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int q = k * 32 + lane;
        cv[k] = q < dp4 ? __ldcg(c4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): half the memory instructions of the 128-bit forms. The
// update warps are bound by memory instructions in flight, not by bytes.
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldcg256(const float* p) {
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void stg256(float* p, const F8& r) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}
// Residual rows are streamed (read once per layer, written once per stack): with an evict-first priority in L2 they
// stop displacing the fp16 operand rows, which the TMA reads back two jobs after the update warps wrote them.
#ifndef NAT_L2_HINTS
#define NAT_L2_HINTS 1
#endif
__device__ __forceinline__ F8 ldg256_stream(const float* p) {
#if NAT_L2_HINTS
    F8 r;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p) : "memory");
    return r;
#else
    return ldcg256(p);
#endif
}
__device__ __forceinline__ void stg256_stream(float* p, const F8& r) {
#if NAT_L2_HINTS
    asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
#else
    stg256(p, r);
#endif
}
// The same update on a row held as NH groups of eight consecutive floats per lane (group g = floats (g * 32 + lane) * 8 ...).
template <int NH>
__device__ __forceinline__ void replay_update8(F8 (&rv)[NH], const float* __restrict__ c, int lane) {
#pragma unroll
    for (int g = 0; g < NH; ++g) {
        const F8 cv = ldcg256(c + (g * 32 + lane) * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float t = __fsub_rn(cv.v[e], rv[g].v[e]);
            rv[g].v[e] = __fsub_rn(rv[g].v[e], __fadd_rn(rv[g].v[e], t));
        }
    }
}

// r <- r - (r + (c - r)) for one code vector: the reference's update (nat.py:2159, 2167, 1405) on a row held in registers.
template <int NV>
__device__ __forceinline__ void replay_update(float4 (&rv)[NV], const float4* __restrict__ c4, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float4 cv = __ldcg(c4 + k * 32 + lane);
        float t;
        t = __fsub_rn(cv.x, rv[k].x); rv[k].x = __fsub_rn(rv[k].x, __fadd_rn(rv[k].x, t));
        t = __fsub_rn(cv.y, rv[k].y); rv[k].y = __fsub_rn(rv[k].y, __fadd_rn(rv[k].y, t));
        t = __fsub_rn(cv.z, rv[k].z); rv[k].z = __fsub_rn(rv[k].z, __fadd_rn(rv[k].z, t));
        t = __fsub_rn(cv.w, rv[k].w); rv[k].w = __fsub_rn(rv[k].w, __fadd_rn(rv[k].w, t));
    }
}
// a - b on both halves, rounded exactly like two __fsub_rn (one fused multiply by -1, one rounding)
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }

// Two sums and one maximum over the warp in 8 shuffles instead of 15: the lanes first split the quantities between
// them (bit 0 of the lane: which sum; bit 1: sum or maximum), then reduce each over the lanes that hold it.
// Results are valid in lane 0. (Shuffles travel through the same shared-memory pipe as the MMA's operand reads.)
#ifndef NAT_REDUCE3
#define NAT_REDUCE3 0
#endif
__device__ __forceinline__ void warp_reduce_2sum_1max(float& s0, float& s1, float& m, int lane) {
#if NAT_REDUCE3
    const bool b0 = lane & 1, b1 = lane & 2;
    float v = (b0 ? s1 : s0) + __shfl_xor_sync(0xffffffffu, b0 ? s0 : s1, 1);       // pair sums: even lanes s0, odd lanes s1
    float mm = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    const float recv = __shfl_xor_sync(0xffffffffu, b1 ? v : mm, 2);
    v = b1 ? fmaxf(mm, recv) : v + recv;                                             // lanes with bit 1: the maximum
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = b1 ? fmaxf(v, t) : v + t;
    }
    s0 = v;                                                                          // lane 0: s0, lane 1: s1, lane 2: max
    s1 = __shfl_sync(0xffffffffu, v, 1);
    m = __shfl_sync(0xffffffffu, v, 2);
#else
    s0 = warp_sum(s0); s1 = warp_sum(s1); m = warp_max(m);
#endif
}

// PAIR = 1: one CTA per tile, tcgen05.mma.cta_group::1 (M 128 x N 256).
// PAIR = 2: launched as clusters of two CTAs (the two SMs of a TPC). Each CTA still owns its own 128-frame tiles,
//   candidate lists, update warps and accumulators, but the pair runs ONE tcgen05.mma.cta_group::2 (M 256 x N 256)
//   issued by the even CTA: each CTA stages only the 128 codes x 64 half of every B tile, so the codebook traffic
//   from L2 per frame halves. Both CTAs walk the same number of jobs (the odd CTA may end on a phantom tile whose
//   frames are all out of range). Cross-CTA protocol: both TMA producers signal the leader's `full` barrier, the
//   leader's commits arrive on `empty` / `tfull` in both CTAs, both candidate warpgroups arrive on the leader's
//   `tempty`.
template <int NV, int PAIR, bool DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
rvq_stack_kernel(const __grid_constant__ CUtensorMap map_a0,  // fp16 [rows, dp], box 64 x 128, SWIZZLE_128B: the prepared operand rows (layer 0)
                 const __grid_constant__ CUtensorMap map_a,   // ditto: the operand rows the update warps write (layers >= 1)
                 const __grid_constant__ CUtensorMap map_b,   // fp16 [L*kp, dp], box 64 x (256 / PAIR), SWIZZLE_128B: the stack's codebooks
                 const __grid_constant__ StackArgs p) {
    static_assert(PAIR == 1 || PAIR == 2, "one CTA or a CTA pair");
    // Dependents (the next stack of the same tokenizer, launched programmatically) share nothing this kernel writes:
    // they may take an SM as soon as one of this grid's CTAs leaves it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const StackRef& sk = p.s;
    const int lb = static_cast<int>(blockIdx.x);
    const int part = static_cast<int>(gridDim.x);
    using SM = Smem<PAIR>;
    using WaitClock = WaitClockT<DBG>;
    const int dbg_mode = DBG ? p.dbg_mode : 0;
    unsigned long long* const dbg_out = DBG ? p.dbg : nullptr;
    constexpr int STAGES = SM::STAGES, B_BYTES = SM::B_BYTES;
    constexpr int OFF_A = SM::OFF_A, OFF_B = SM::OFF_B, OFF_LIST_S = SM::OFF_LIST_S, OFF_LIST_I = SM::OFF_LIST_I,
                  OFF_HAND = SM::OFF_HAND, OFF_CN = SM::OFF_CN, OFF_BARS = SM::OFF_BARS;
    const uint32_t crank = PAIR == 2 ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem + OFF_A;
    uint8_t* smem_b = smem + OFF_B;
    float* list_s = reinterpret_cast<float*>(smem + OFF_LIST_S);                    // [LCAP][128]
    uint32_t* list_i = reinterpret_cast<uint32_t*>(smem + OFF_LIST_I);              // [LCAP][128] group id << 8 | mask
    uint4* hand = reinterpret_cast<uint4*>(smem + OFF_HAND);                        // [2][128] records of two uint4
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* full = bars;                  // TMA -> MMA
    uint64_t* empty = bars + STAGES;        // MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;    // MMA -> candidates (accumulator stage ready)
    uint64_t* tempty = tfull + 2;           // candidates -> MMA (accumulator stage drained)
    uint64_t* cfull = tempty + 2;           // candidates -> update (hand-off slot written)
    uint64_t* cempty = cfull + 2;           // update -> candidates (hand-off slot read)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + 2);
    unsigned* ready = tmem_slot + 1;        // [UPD_WARPS] jobs finished by each update warp

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_chunks = p.kp / BLOCK_N;
    const int n_kblocks = p.dp / BLOCK_K;
    // PAIR = 2: both CTAs take the leader's count, so the odd CTA may run one phantom tile (index >= n_tiles)
    const int my_tiles = (p.n_tiles - (lb - static_cast<int>(crank)) + part - 1) / part;
    const int group = max(1, p.group);

    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EPI_WARPS * PAIR);
            mbar_init(&cfull[i], EPI_WARPS);
            mbar_init(&cempty[i], UPD_WARPS);
        }
        for (int i = 0; i < UPD_WARPS; ++i) ready[i] = 0;
        fence_mbar_init();
    } else if (warp == WARP_MMA) {
        if (PAIR == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS);
        else tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tcgen05_fence_before();
    if (PAIR == 2) cluster_sync_all();      // the peer's barriers must exist before anything arrives on them
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Registers follow the work: the update warps hold two residual rows and two code vectors in flight.
    // (setmaxnreg sits at the top of each role's branch so that ptxas budgets the branch accordingly.)
    // Every role walks the same job sequence:  for each group g0 .. : for each layer l : for each tile i of the group.
    if (warp >= WARP_TMA) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    if (warp == WARP_TMA) {
        // ------------------------------------------------------------------ TMA producer (lane 0 issues)
        uint32_t s = 0, ph = 0;
        unsigned job = 0;
        WaitClock w_ready(dbg_out != nullptr), w_empty(dbg_out != nullptr);
        const long long t_start = w_ready.begin();
        for (int g0 = 0; g0 < my_tiles; g0 += group) {
            const int gs = min(group, my_tiles - g0);
            for (int l = 0; l < sk.L; ++l) {
                const CUtensorMap* map_al = l == 0 ? &map_a0 : &map_a;
                for (int i = g0; i < g0 + gs; ++i, ++job) {
                    // a phantom tile re-reads the last real tile's operand (its results are never used)
                    const int tile = min(lb + i * part, p.n_tiles - 1);
                    if (l > 0) {
                        const long long t0 = w_ready.begin();
                        // operand rows of (tile, l) are written by this CTA's update warps in job (tile, l-1):
                        // every update warp must have finished that job (they progress independently)
                        if (lane < UPD_WARPS) wait_counter(&ready[lane], job - gs + 1);
                        __syncwarp();
                        fence_proxy_async();
                        w_ready.end(t0);
                    }
                    if (lane == 0) {
                        for (int chunk = 0; chunk < n_chunks; ++chunk) {
                            for (int kb = 0; kb < n_kblocks; ++kb) {
                                const long long t1 = w_empty.begin();
                                mbar_wait(&empty[s], ph ^ 1);
                                w_empty.end(t1);
                                if (PAIR == 2) {
                                    // both CTAs' bytes are counted on the leader's barrier
                                    const bool skip_a = (dbg_mode & 16) && chunk > 0;      // timing experiment only
                                    if (leader) mbar_arrive_expect_tx(&full[s], 2 * ((skip_a ? 0 : A_STAGE_BYTES) + B_BYTES));
                                    const uint32_t bar = mapa_rank(smem_u32(&full[s]), 0);
                                    if (!skip_a)
                                    tma_load_2d_pair(smem_a + s * A_STAGE_BYTES, map_al, bar, kb * BLOCK_K, tile * BLOCK_M);
                                    tma_load_2d_pair(smem_b + s * B_BYTES, &map_b, bar, kb * BLOCK_K,
                                                     l * p.kp + chunk * BLOCK_N + crank * (BLOCK_N / 2));
                                } else {
                                    mbar_arrive_expect_tx(&full[s], A_STAGE_BYTES + B_BYTES);
                                    tma_load_2d(smem_a + s * A_STAGE_BYTES, map_al, &full[s], kb * BLOCK_K, tile * BLOCK_M);
                                    tma_load_2d(smem_b + s * B_BYTES, &map_b, &full[s], kb * BLOCK_K,
                                                l * p.kp + chunk * BLOCK_N);
                                }
                                if (++s == STAGES) { s = 0; ph ^= 1; }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (dbg_out != nullptr && lane == 0) {
            unsigned long long* d = dbg_out + static_cast<size_t>(blockIdx.x) * DBG_SLOTS;
            d[DBG_TMA_WAIT_READY] = w_ready.acc;
            d[DBG_TMA_WAIT_EMPTY] = w_empty.acc;
            d[DBG_KERNEL_TOTAL] = clock64() - t_start;
        }
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(BLOCK_M * PAIR, BLOCK_N);
            uint32_t s = 0, ph = 0;
            const int total = sk.L * my_tiles * n_chunks;
            WaitClock w_tempty(dbg_out != nullptr), w_full(dbg_out != nullptr);
            for (int it = 0; it < total; ++it) {
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                const long long t0 = w_tempty.begin();
                mbar_wait(&tempty[as], aph ^ 1);
                w_tempty.end(t0);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    const long long t1 = w_full.begin();
                    mbar_wait(&full[s], ph);
                    w_full.end(t1);
                    tcgen05_fence_after();
                    const uint64_t adesc = umma_desc_kmajor_sw128(smem_a + s * A_STAGE_BYTES);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(smem_b + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        if (PAIR == 2) umma_f16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        else umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    if (PAIR == 2) {
                        umma_commit_pair(&empty[s]);
                        if (kb == n_kblocks - 1) umma_commit_pair(&tfull[as]);
                    } else {
                        umma_commit(&empty[s]);
                        if (kb == n_kblocks - 1) umma_commit(&tfull[as]);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
            if (dbg_out != nullptr) {
                unsigned long long* d = dbg_out + static_cast<size_t>(blockIdx.x) * DBG_SLOTS;
                d[DBG_MMA_WAIT_TEMPTY] = w_tempty.acc;
                d[DBG_MMA_WAIT_FULL] = w_full.acc;
            }
        }
        __syncwarp();
    } else if (warp < WARP_UPD0) {
        if (REGS_EPI > LAUNCH_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        else if (REGS_EPI < LAUNCH_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        // ------------------------------------------------------------------ candidates: one frame per thread
        // Scores are looked at eight at a time. Chunk 0 is read twice: first only for its minimum (branch-free), so
        // that the filter pass starts with a threshold that is already within a few records of the final one; every
        // later element then costs one FFMA and half a min, and the rare eight-group whose minimum is inside the
        // window appends {group minimum, group id, 8-bit mask of its elements inside the window} to the frame's list.
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int tid = q * 32 + lane;                // frame within the tile
        const uint32_t ls = smem_u32(list_s) + tid * 4;      // entry e at + e * 512
        const uint32_t li = smem_u32(list_i) + tid * 4;
        // ||c||^2 of a chunk is staged in shared memory one chunk ahead by these four warps themselves (two floats
        // per thread), so that scoring never waits on a global load.
        const uint32_t cnbuf = smem_u32(smem + OFF_CN);
        sts_f32x2(cnbuf + tid * 8, __ldg(reinterpret_cast<const float2*>(sk.cn32) + tid));
        uint32_t it = 0, job = 0;
        WaitClock w_tfull(dbg_out != nullptr), w_cempty(dbg_out != nullptr), w_ld(dbg_out != nullptr);
        [[maybe_unused]] unsigned long long n_events = 0;
        const long long t_epi = w_tfull.begin();
        for (int g0 = 0; g0 < my_tiles; g0 += group) {
            const int gs = min(group, my_tiles - g0);
            for (int l = 0; l < sk.L; ++l) {
                for (int i = g0; i < g0 + gs; ++i, ++job) {
                    const int tile = lb + i * part;
                    const long long row = static_cast<long long>(tile) * BLOCK_M + tid;
                    const bool valid = row < p.n_rows;
                    const float inf = __int_as_float(0x7F800000);
                    float alpha = 0.f, window = 0.f;
                    float m = inf;
                    float thr = valid ? inf : -inf;   // frames outside the input never collect anything
                    int cnt = 0;
                    bool overflow = false;

                    // drop list entries whose group minimum has fallen out of the (only ever shrinking) window
                    auto compact = [&]() {
                        int w = 0;
                        for (int e = 0; e < cnt; ++e) {
                            const float sv = lds_f32(ls + e * (BLOCK_M * 4));
                            const uint32_t iv = lds_u32(li + e * (BLOCK_M * 4));
                            if (sv <= thr) {
                                sts_f32(ls + w * (BLOCK_M * 4), sv);
                                sts_u32(li + w * (BLOCK_M * 4), iv);
                                ++w;
                            }
                        }
                        cnt = w;
                    };

                    for (int chunk = 0; chunk < n_chunks; ++chunk, ++it) {
                        const uint32_t as = it & 1, aph = (it >> 1) & 1;
                        const long long t0 = w_tfull.begin();
                        mbar_wait(&tfull[as], aph);
                        w_tfull.end(t0);
                        tcgen05_fence_after();
                        if (chunk == 0 && valid) {
                            // written by this CTA's update warps one layer ago: only now is it known to be there
                            const float4 ri = __ldcg((l == 0 ? sk.rowinfo0 : sk.rowinfo) + row);
                            alpha = ri.x;
                            window = ri.z;
                        }
                        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
                        // (layer, chunk) after this one in the job sequence; past the end it wraps to (0, 0)
                        int nl = l, nchunk = chunk + 1;
                        if (nchunk == n_chunks) {
                            nchunk = 0;
                            if (i + 1 >= g0 + gs) nl = (l + 1 < sk.L) ? l + 1 : 0;
                        }
                        const float2 cn_next = __ldg(reinterpret_cast<const float2*>(
                                                         sk.cn32 + static_cast<long long>(nl) * p.kp + nchunk * BLOCK_N) + tid);
                        bar_sync_named(1, EPI_WARPS * 32);         // this chunk's buffer is complete and visible
                        const uint32_t cn_s = cnbuf + (it & 1) * (BLOCK_N * 4);
                        const int gid0 = (chunk * BLOCK_N) >> 3;

                        // FILTER = false: running minimum only.  FILTER = true: minimum + list of groups in the window.
                        // All 32 scores and the four eight-group minima are formed first (independent instructions), and
                        // ONE test against the threshold guards the list code: thr only ever shrinks, so a block whose
                        // smallest score is above it has no entry to add. Inside, the groups are taken in order with the
                        // threshold updated after each entry, exactly as if they had been tested one by one.
                        auto scan32 = [&](const uint32_t (&v)[32], int g, bool filter) {
                            float s[32];
                            float mn[4];
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const float4 c0 = lds_f32x4(cn_s + (g * 32 + h * 8) * 4);
                                const float4 c1 = lds_f32x4(cn_s + (g * 32 + h * 8 + 4) * 4);
                                s[h * 8 + 0] = fmaf(__uint_as_float(v[h * 8 + 0]), alpha, c0.x);
                                s[h * 8 + 1] = fmaf(__uint_as_float(v[h * 8 + 1]), alpha, c0.y);
                                s[h * 8 + 2] = fmaf(__uint_as_float(v[h * 8 + 2]), alpha, c0.z);
                                s[h * 8 + 3] = fmaf(__uint_as_float(v[h * 8 + 3]), alpha, c0.w);
                                s[h * 8 + 4] = fmaf(__uint_as_float(v[h * 8 + 4]), alpha, c1.x);
                                s[h * 8 + 5] = fmaf(__uint_as_float(v[h * 8 + 5]), alpha, c1.y);
                                s[h * 8 + 6] = fmaf(__uint_as_float(v[h * 8 + 6]), alpha, c1.z);
                                s[h * 8 + 7] = fmaf(__uint_as_float(v[h * 8 + 7]), alpha, c1.w);
                                mn[h] = fminf(fminf(fminf(fminf(s[h * 8 + 0], s[h * 8 + 1]), s[h * 8 + 2]),
                                                    fminf(fminf(s[h * 8 + 3], s[h * 8 + 4]), s[h * 8 + 5])),
                                              fminf(fminf(s[h * 8 + 6], s[h * 8 + 7]), inf));
                            }
                            const float mall = fminf(fminf(mn[0], mn[1]), fminf(mn[2], mn[3]));
                            if (!filter) {
                                m = fminf(m, mall);
                                return;
                            }
                            if (!(mall <= thr)) return;
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                if (mn[h] <= thr) {
                                    if (DBG) ++n_events;
                                    m = fminf(m, mn[h]);
                                    // m + W rounded up: the kept set must be a superset of the exact window
                                    thr = fmaf(fabsf(m) + window, 2.4e-7f, m + window);
                                    uint32_t mask = 0;
#pragma unroll
                                    for (int e = 0; e < 8; ++e) mask |= (s[h * 8 + e] <= thr) ? (1u << e) : 0u;
                                    if (cnt == LCAP) compact();
                                    if (cnt < LCAP) {
                                        sts_f32(ls + cnt * (BLOCK_M * 4), mn[h]);
                                        sts_u32(li + cnt * (BLOCK_M * 4),
                                                (static_cast<uint32_t>(gid0 + g * 4 + h) << 8) | mask);
                                        ++cnt;
                                    } else {
                                        overflow = true;
                                    }
                                }
                            }
                        };

                        uint32_t va[32];
                        if (chunk == 0) {
#pragma unroll 1
                            for (int g = 0; g < BLOCK_N / 32; ++g) {
                                tmem_ld_32x32b_x32(taddr + g * 32, va);
                                { const long long tw = w_ld.begin(); tmem_wait_ld(); w_ld.end(tw); }
                                if ((dbg_mode & 15) < 6) scan32(va, g, false);
                            }
                            if (valid) thr = fmaf(fabsf(m) + window, 2.4e-7f, m + window);
                        }
#pragma unroll 1
                        for (int g = 0; g < BLOCK_N / 32; ++g) {
                            tmem_ld_32x32b_x32(taddr + g * 32, va);
                            { const long long tw = w_ld.begin(); tmem_wait_ld(); w_ld.end(tw); }
                            if ((dbg_mode & 15) < 6) scan32(va, g, true);
                        }
                        // accumulator stage drained: hand it back to the MMA warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR == 2) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[as]), 0));
                            else mbar_arrive(&tempty[as]);
                        }
                        sts_f32x2(cnbuf + ((it + 1) & 1) * (BLOCK_N * 4) + tid * 8, cn_next);
                    }
                    // hand the survivors to the update warps: u16 count + up to HCAP u16 code indices per frame
                    compact();
                    const uint32_t slot = job & 1;
                    const long long t1 = w_cempty.begin();
                    mbar_wait(&cempty[slot], ((job >> 1) & 1) ^ 1);
                    w_cempty.end(t1);
                    const uint32_t hrec = smem_u32(hand) + (slot * BLOCK_M + tid) * HAND_BYTES;
                    unsigned n = 0;
                    for (int e = 0; e < cnt; ++e) {
                        const uint32_t iv = lds_u32(li + e * (BLOCK_M * 4));
                        const uint32_t base = (iv >> 8) << 3;
                        uint32_t mask = iv & 0xFFu;
                        while (mask != 0) {
                            const uint32_t b = __ffs(mask) - 1;
                            mask &= mask - 1;
                            if (n < HCAP) sts_u16(hrec + 2 + n * 2, static_cast<unsigned short>(base + b));
                            ++n;
                        }
                    }
                    if ((dbg_mode & 15) >= 6) { n = 1; sts_u16(hrec + 2, 0); overflow = false; }
                    if (!valid) n = 0;
                    else if (overflow || n == 0 || n > HCAP) n = HAND_SCAN;
                    sts_u16(hrec, static_cast<unsigned short>(n));
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cfull[slot]);
                }
            }
        }
        if (dbg_out != nullptr && warp == WARP_EPI0 && lane == 0) {
            unsigned long long* d = dbg_out + static_cast<size_t>(blockIdx.x) * DBG_SLOTS;
            d[DBG_EPI_WAIT_TFULL] = w_tfull.acc;
            d[DBG_EPI_WAIT_CEMPTY] = w_cempty.acc;
            d[DBG_EPI_TOTAL] = clock64() - t_epi;
            d[DBG_EPI_WAIT_LD] = w_ld.acc;
            d[DBG_EPI_EVENTS] = n_events;
        }
    } else if (warp < WARP_TMA) {
        if (REGS_UPD > LAUNCH_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_UPD));
        else if (REGS_UPD < LAUNCH_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_UPD));
        // ------------------------------------------------------------------ update: one warp per frame
        // Phase 0 settles the frames the coarse pass could not certify (exact fp64 re-rank, or the exact scan).
        // Phase 1 is branch-free: row rr+1's residual and code vector are in flight while row rr is updated, and the
        // warp reductions + error window of row rr-1 are scheduled under row rr's arithmetic. The operand scale of
        // the next layer comes from a bound (max|r| + max|c|), not from the new row, so no element waits on a
        // reduction; the exact new max is reduced afterwards and kept for the next layer's bound.
        const int uw = warp - WARP_UPD0;
        const int dp4 = p.dp >> 2;
        uint32_t job = 0;
        WaitClock w_cfull(dbg_out != nullptr), w_res(dbg_out != nullptr), w_dec(dbg_out != nullptr), w_fence(dbg_out != nullptr);
        const long long t_upd = w_cfull.begin();
        for (int g0 = 0; g0 < my_tiles; g0 += group) {
            const int gs = min(group, my_tiles - g0);
            for (int l = 0; l < sk.L; ++l) {
                const float* cb_l = sk.cbf + static_cast<long long>(l) * p.K * p.dp;
                const double* cn64_l = sk.cn64 + static_cast<long long>(l) * p.K;
                const float* cn32_l = sk.cn32 + static_cast<long long>(l) * p.kp;   // == fl32(cn64) for every code that can be a candidate
                const bool last = l + 1 == sk.L;
                const rows::LayerConst* lc_next = last ? nullptr : sk.lc + l + 1;
                const float cabs = __ldg(&sk.lc[l].cabs);
                double* loss_l = sk.row_loss != nullptr ? sk.row_loss + static_cast<long long>(l) * sk.loss_ld : nullptr;
                const bool residual_needed = !last || loss_l != nullptr;
                // Hot form only: layers whose store bit is clear did not write their residual back; the stored row is the
                // one that entered the oldest such layer, and this layer replays their updates from the emitted codes.
                const int store_mask = (sk.row_loss == nullptr && dp4 == NV * 32) ? sk.store_mask : ~0;
                int n_replay = 0;
                for (int jj = l - 1; jj >= 0 && !((store_mask >> jj) & 1); --jj) ++n_replay;
                // rows this layer reads: the prepared ones until an earlier layer of the stack has written a residual back
                const float* r_src = n_replay == l ? sk.r0 : sk.r;
                char* codes_l = static_cast<char*>(sk.codes);
                const long long code_base = static_cast<long long>(l) * sk.codes_ld + sk.code_off;
                for (int i = g0; i < g0 + gs; ++i, ++job) {
                    const int tile = lb + i * part;
                    const int row0 = tile * BLOCK_M + uw * ROWS_PER_UPD_WARP;
                    const int nrows = max(0, min(ROWS_PER_UPD_WARP, p.n_rows - row0));
                    float am_old = 0.f;                // lane rr: max |r| of row rr before this layer
                    // (No early L2 prefetch of the residual rows here: issued a whole GEMM ahead, the lines were evicted
                    // again before use and cost a second DRAM read; measured 4.6 % slower. A late one, issued when the
                    // job's candidates arrive, measured 4.6 % slower than none as well.)
                    if (residual_needed && nrows > 0 && lane < nrows) am_old = __ldcg((l == 0 ? sk.rowamax0 : sk.rowamax) + row0 + lane);
                    const uint32_t slot = job & 1;
                    const long long t0 = w_cfull.begin();
                    mbar_wait(&cfull[slot], (job >> 1) & 1);
                    w_cfull.end(t0);
                    // lanes 2 rr and 2 rr + 1 hold the two halves of row rr's record
                    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
                    if (lane < 2 * ROWS_PER_UPD_WARP) rec = hand[(slot * BLOCK_M + uw * ROWS_PER_UPD_WARP) * 2 + lane];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cempty[slot]);
                    if ((dbg_mode & 15) == 3 || (dbg_mode & 15) == 7) {
                        __syncwarp();
                        if (lane == 0) st_release(&ready[uw], job + 1);
                        continue;
                    }

                    // ---- phase 0: decisions. lane rr ends up holding the code of row rr in jsel.
                    const long long tA = w_dec.begin();
                    const uint32_t first_word = __shfl_sync(0xffffffffu, rec.x, (lane * 2) & 31);   // row `lane`, word 0
                    const unsigned n_mine = first_word & 0xFFFFu;
                    int jsel = static_cast<int>(first_word >> 16);
                    unsigned todo = __ballot_sync(0xffffffffu, lane < nrows && n_mine != 1u);
                    unsigned n_rerank = 0, n_scan = 0, n_fp64 = 0;
                    const unsigned n_cert = static_cast<unsigned>(nrows) - __popc(todo);
                    while (todo != 0) {
                        const int rr = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const unsigned n = __shfl_sync(0xffffffffu, rec.x, 2 * rr) & 0xFFFFu;
                        // candidate c sits in 16-bit slot c + 1 of the record: word (c + 1) / 2 of lanes 2 rr, 2 rr + 1
                        auto candidate = [&](int c) -> int {
                            const int w = (c + 1) >> 1;
                            const uint32_t mine = (w & 3) == 0 ? rec.x : (w & 3) == 1 ? rec.y : (w & 3) == 2 ? rec.z : rec.w;
                            const uint32_t word = __shfl_sync(0xffffffffu, mine, 2 * rr + (w >> 2));
                            return static_cast<int>(((c + 1) & 1) ? (word >> 16) : (word & 0xFFFFu));
                        };
                        float4 rv[NV];
                        load_row<NV>(rv, reinterpret_cast<const float4*>(r_src + static_cast<long long>(row0 + rr) * p.dp), dp4, lane);
                        if constexpr (true) {
#pragma unroll 1
                            for (int back = n_replay; back > 0; --back) {
                                const int jp = rows::load_code(codes_l, p.code_dtype,
                                                               code_base - back * sk.codes_ld + row0 + rr);
                                if (dp4 == NV * 32)
                                    replay_update<NV>(rv, reinterpret_cast<const float4*>(
                                        cb_l + (static_cast<long long>(jp) - static_cast<long long>(back) * p.K) * p.dp), lane);
                            }
                        }
                        double best = 0.0;
                        int bestj = -1;
                        if (n == HAND_SCAN) {
                            // exact scan of every code, two in flight; k ascending so the first minimum is kept
                            for (int k0 = 0; k0 < p.K; k0 += 2) {
                                const int k1 = min(k0 + 1, p.K - 1);
                                const double s0 = score_regs<NV>(rv, reinterpret_cast<const float4*>(cb_l + static_cast<long long>(k0) * p.dp),
                                                                 dp4, cn64_l[k0], lane);
                                const double s1 = score_regs<NV>(rv, reinterpret_cast<const float4*>(cb_l + static_cast<long long>(k1) * p.dp),
                                                                 dp4, cn64_l[k1], lane);
                                if (bestj < 0 || s0 < best) { best = s0; bestj = k0; }
                                if (k1 > k0 && s1 < best) { best = s1; bestj = k1; }
                            }
                            ++n_scan;
                        } else {
                            // Screening in fp32 first: lane c keeps an interval that contains candidate c's exact score
                            // ||c||^2 - 2 r.c (|fl(r.c) - r.c| <= gamma_22 sum |r_i c_i|, gamma_22 < 1.4e-6, doubled by
                            // the factor 2 and widened to 4.9e-6 to cover the rounding of the bound itself; 1.3e-7 of
                            // ||c||^2 and |s| covers the roundings of fl32(||c||^2), of s and of the interval ends).
                            // If the interval with the lowest upper end lies strictly below all the others, its code is
                            // the fp64 answer. Otherwise (about one re-rank in 200) the candidates go through fp64.
                            float lo_c = 0.f, hi_c = 0.f;
                            int k_c = 0;
#pragma unroll 1
                            for (int c = 0; c < static_cast<int>(n); ++c) {        // warp-uniform trip count, <= HCAP
                                const int k = min(candidate(c), p.K - 1);
                                float dot, adot;
                                // (loading the first candidate together with the frame spills: 328 B, not adopted)
                                float4 cv0[NV];
                                load_code<NV>(cv0, reinterpret_cast<const float4*>(cb_l + static_cast<long long>(k) * p.dp), dp4, lane);
                                score32_regs<NV>(rv, cv0, dot, adot);
                                const float cn = __ldg(cn32_l + k);
                                const float sc = fmaf(-2.f, dot, cn);
                                const float e = fmaf(4.9e-6f, adot, 1.3e-7f * (cn + fabsf(sc))) + 1e-30f;
                                if (lane == c) { lo_c = sc - e; hi_c = sc + e; k_c = k; }
                            }
                            const bool mine = lane < static_cast<int>(n);
                            float hmin = mine ? hi_c : __int_as_float(0x7F800000);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) hmin = fminf(hmin, __shfl_xor_sync(0xffffffffu, hmin, o));
                            const unsigned win = __ballot_sync(0xffffffffu, mine && hi_c == hmin);
                            const int wl = __ffs(win) - 1;
                            const bool apart = !mine || lane == wl || lo_c > hmin;      // false on NaN: falls through
                            if (__popc(win) == 1 && __all_sync(0xffffffffu, apart)) {
                                bestj = __shfl_sync(0xffffffffu, k_c, wl);
                            } else {
#pragma unroll 1
                                for (int c = 0; c < static_cast<int>(n); ++c) {
                                    const int k = min(candidate(c), p.K - 1);
                                    const double s = score_regs<NV>(rv, reinterpret_cast<const float4*>(cb_l + static_cast<long long>(k) * p.dp),
                                                                    dp4, cn64_l[k], lane);
                                    if (bestj < 0 || s < best || (s == best && k < bestj)) { best = s; bestj = k; }
                                }
                                ++n_fp64;
                            }
                            ++n_rerank;
                        }
                        if (lane == rr) jsel = bestj;
                    }
                    jsel = max(0, min(jsel, p.K - 1));
                    if (lane < nrows) rows::store_code(codes_l, p.code_dtype, code_base + row0 + lane, jsel);

                    w_dec.end(tA);
                    // ---- phase 1: residual update, next operand, error window
                    const long long tB = w_res.begin();
                    if (!last && loss_l == nullptr && dp4 == NV * 32 && nrows > 0) {
                        // Hot form: every lane holds exactly NV / 2 groups of eight consecutive floats of a row (256-bit
                        // loads and stores), nothing is predicated. One row at a time per warp; the sixteen update warps
                        // of the CTA cover each other's memory latency.
                        // The residual entering the LAST layer is only ever seen through its fp16 operand (the last
                        // layer emits codes and, without a loss, nothing else): it is not written back.
                        const bool keep_r = (store_mask >> l) & 1;
#pragma unroll 1
                        for (int rr = 0; rr < nrows; ++rr) {
                            const int row = row0 + rr;
                            const int j = __shfl_sync(0xffffffffu, jsel, rr);
                            constexpr int NH = NV / 2;                 // groups of eight floats per lane
                            float* rrow = sk.r + static_cast<long long>(row) * p.dp;
                            const float* crow = cb_l + static_cast<long long>(j) * p.dp;
                            F8 cur[NH], cv[NH];
#pragma unroll
                            for (int g = 0; g < NH; ++g)
                                if (!(dbg_mode & 128)) cur[g] = ldg256_stream(r_src + static_cast<long long>(row) * p.dp + (g * 32 + lane) * 8);
                                else for (int e = 0; e < 8; ++e) cur[g].v[e] = 1.f;          // timing experiment only
#pragma unroll 1
                            for (int back = n_replay; back > 0; --back) {
                                const int jp = rows::load_code(codes_l, p.code_dtype, code_base - back * sk.codes_ld + row);
                                replay_update8<NH>(cur, cb_l + (static_cast<long long>(jp) - static_cast<long long>(back) * p.K) * p.dp, lane);
                            }
#pragma unroll
                            for (int g = 0; g < NH; ++g)
                                if (!(dbg_mode & 32)) cv[g] = ldcg256(crow + (g * 32 + lane) * 8);
                                else for (int e = 0; e < 8; ++e) cv[g].v[e] = 0.5f;          // timing experiment only
                            // operand scale from the bound  max|r'| <= max|r| + max|c|
                            const float bound = (__shfl_sync(0xffffffffu, am_old, rr) + cabs) * 1.00001f;
                            const float sx = rows::pow2_scale_for(bound);
                            const float2 sx2 = make_float2(sx, sx);
                            uint4* a_row = reinterpret_cast<uint4*>(sk.a + static_cast<long long>(row) * p.dp);
                            float2 lo2v = make_float2(0.f, 0.f), xh2v = make_float2(0.f, 0.f);
                            float amax = 0.f;
#pragma unroll
                            for (int g = 0; g < NH; ++g) {
                                // q = codebook[j]; t = q - r; q_ste = r + t; r' = r - q_ste   (nat.py:2159, 2167, 1405)
                                F8 nr;
                                uint32_t hp[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 x2 = make_float2(cur[g].v[2 * e], cur[g].v[2 * e + 1]);
                                    const float2 c2 = make_float2(cv[g].v[2 * e], cv[g].v[2 * e + 1]);
                                    const float2 t2 = sub2(c2, x2);
                                    const float2 n2 = sub2(x2, __fadd2_rn(x2, t2));
                                    nr.v[2 * e] = n2.x; nr.v[2 * e + 1] = n2.y;
                                    amax = fmaxf(amax, fmaxf(fabsf(n2.x), fabsf(n2.y)));
                                    const float2 s2 = __fmul2_rn(n2, sx2);
                                    const __half2 h2 = __float22half2_rn(s2);
                                    const float2 l2 = sub2(s2, __half22float2(h2));
                                    lo2v = __ffma2_rn(l2, l2, lo2v);
                                    xh2v = __ffma2_rn(s2, s2, xh2v);
                                    hp[e] = *reinterpret_cast<const uint32_t*>(&h2);
                                }
                                if (keep_r && !(dbg_mode & 64)) stg256_stream(rrow + (g * 32 + lane) * 8, nr);
                                if (!(dbg_mode & 64)) a_row[g * 32 + lane] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                            }
                            float lo2 = lo2v.x + lo2v.y, xh2 = xh2v.x + xh2v.y, amx = amax;
                            warp_reduce_2sum_1max(lo2, xh2, amx, lane);
                            if (lane == 0) {
                                sk.rowinfo[row] = rows::make_rowinfo_bound(sx, lo2, xh2, lc_next, p.dp);
                                sk.rowamax[row] = amx;
                            }
                        }
                    } else
                    if (residual_needed && nrows > 0) {
                        // General form (ragged D, last layer with loss): same arithmetic, predicated per float4.
#pragma unroll 1
                        for (int rr = 0; rr < nrows; ++rr) {
                            const int row = row0 + rr;
                            const int j = __shfl_sync(0xffffffffu, jsel, rr);
                            float4* r4 = reinterpret_cast<float4*>(sk.r + static_cast<long long>(row) * p.dp);
                            float4 cur[NV], cv[NV];
                            load_row<NV>(cur, reinterpret_cast<const float4*>(r_src + static_cast<long long>(row) * p.dp), dp4, lane);
                            load_code<NV>(cv, reinterpret_cast<const float4*>(cb_l + static_cast<long long>(j) * p.dp), dp4, lane);
                            const float bound = (__shfl_sync(0xffffffffu, am_old, rr) + cabs) * 1.00001f;
                            const float sx = rows::pow2_scale_for(bound);
                            const float2 sx2 = make_float2(sx, sx);
                            uint2* a_row = reinterpret_cast<uint2*>(sk.a + static_cast<long long>(row) * p.dp);
                            float2 lo2v = make_float2(0.f, 0.f), xh2v = make_float2(0.f, 0.f);
                            float amax = 0.f;
                            double loss = 0.0;
#pragma unroll
                            for (int k = 0; k < NV; ++k) {
                                const int qq = k * 32 + lane;
                                if (qq < dp4) {
                                    const float2 x01 = make_float2(cur[k].x, cur[k].y), x23 = make_float2(cur[k].z, cur[k].w);
                                    const float2 c01 = make_float2(cv[k].x, cv[k].y), c23 = make_float2(cv[k].z, cv[k].w);
                                    const float2 t01 = sub2(c01, x01), t23 = sub2(c23, x23);
                                    const float2 n01 = sub2(x01, __fadd2_rn(x01, t01)), n23 = sub2(x23, __fadd2_rn(x23, t23));
                                    if (loss_l != nullptr) {
                                        loss += static_cast<double>(__fmul_rn(t01.x, t01.x));
                                        loss += static_cast<double>(__fmul_rn(t01.y, t01.y));
                                        loss += static_cast<double>(__fmul_rn(t23.x, t23.x));
                                        loss += static_cast<double>(__fmul_rn(t23.y, t23.y));
                                    }
                                    if (!last) {
                                        r4[qq] = make_float4(n01.x, n01.y, n23.x, n23.y);
                                        amax = fmaxf(fmaxf(amax, fabsf(n01.x)), fmaxf(fabsf(n01.y), fmaxf(fabsf(n23.x), fabsf(n23.y))));
                                        const float2 s01 = __fmul2_rn(n01, sx2), s23 = __fmul2_rn(n23, sx2);
                                        const __half2 h01 = __float22half2_rn(s01), h23 = __float22half2_rn(s23);
                                        const float2 l01 = sub2(s01, __half22float2(h01)), l23 = sub2(s23, __half22float2(h23));
                                        lo2v = __ffma2_rn(l01, l01, lo2v);
                                        lo2v = __ffma2_rn(l23, l23, lo2v);
                                        xh2v = __ffma2_rn(s01, s01, xh2v);
                                        xh2v = __ffma2_rn(s23, s23, xh2v);
                                        uint2 packed;
                                        packed.x = *reinterpret_cast<const uint32_t*>(&h01);
                                        packed.y = *reinterpret_cast<const uint32_t*>(&h23);
                                        a_row[qq] = packed;
                                    }
                                }
                            }
                            if (loss_l != nullptr) {
                                loss = warp_sum(loss);
                                if (lane == 0) loss_l[row] = loss;
                            }
                            if (!last) {
                                const float lo2 = warp_sum(lo2v.x + lo2v.y), xh2 = warp_sum(xh2v.x + xh2v.y), amx = warp_max(amax);
                                if (lane == 0) {
                                    sk.rowinfo[row] = rows::make_rowinfo_bound(sx, lo2, xh2, lc_next, p.dp);
                                    sk.rowamax[row] = amx;
                                }
                            }
                        }
                    }
                    w_res.end(tB);
                    // publish this warp's share of the job: global writes -> visible to the TMA (async proxy)
                    const long long tD = w_fence.begin();
                    // fence.proxy.async carries a GPU-scope barrier of its own (MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC): the rows
                    // are in L2, where the TMA reads them, when the counter moves. (A __threadfence in front of it would
                    // add a second barrier and invalidate the SM's whole L1 -- CCTL.IVALL -- every job.)
                    fence_proxy_async();
                    __syncwarp();
                    w_fence.end(tD);
                    if (lane == 0) {
                        st_release(&ready[uw], job + 1);
                        if (sk.stats != nullptr) {
                            unsigned long long* st = sk.stats + l * 4;
                            if (n_cert) atomicAdd(st + 0, static_cast<unsigned long long>(n_cert));
                            if (n_rerank) atomicAdd(st + 1, static_cast<unsigned long long>(n_rerank));
                            if (n_scan) atomicAdd(st + 2, static_cast<unsigned long long>(n_scan));
                            if (n_fp64) atomicAdd(st + 3, static_cast<unsigned long long>(n_fp64));
                        }
                    }
                }
            }
        }
        if (dbg_out != nullptr && uw == 0 && lane == 0) {
            unsigned long long* d = dbg_out + static_cast<size_t>(blockIdx.x) * DBG_SLOTS;
            d[DBG_UPD_WAIT_CFULL] = w_cfull.acc;
            d[DBG_UPD_TOTAL] = clock64() - t_upd;
            d[DBG_UPD_RESID] = w_res.acc;
            d[DBG_UPD_DECIDE] = w_dec.acc;
            d[DBG_UPD_FENCE] = w_fence.acc;
        }
    }

    tcgen05_fence_before();
    if (PAIR == 2) cluster_sync_all();      // neither CTA may leave while the other can still reach into it
    else __syncthreads();
    if (warp == WARP_MMA) {
        tcgen05_fence_after();
        if (PAIR == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace stack
}  // namespace nat
