// Codebook preparation: fp32 snapshot (zero padded rows), power-of-two scaled fp16 copy for the tensor cores,
// ||c||^2 in fp32/fp64 and the per-layer norm maxima the error bound needs. Runs once per codebook change
// (the host mutates codebooks with copy_: nat.py:593, 1527, 2221), never on the per-frame path.
#pragma once

#include "nat_common.cuh"
#include "rvq_rows.cuh"

namespace nat {
namespace prepare {

// scratch ints per layer: [0] absmax bits, [1] max ||c_hat||^2, [2] max ||c_hat - fp16||^2, [3] max ||fp16||^2,
// [4] max ||c||^2   (all non-negative floats compared as ints)
constexpr int kScratchPerLayer = 8;

__global__ void __launch_bounds__(256)
pack_absmax_kernel(const float* __restrict__ src, int K, int D, int dp, float* __restrict__ dst,
                   int* __restrict__ scratch) {
    const long long total = static_cast<long long>(K) * dp;
    float amax = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long k = i / dp;
        const int d = static_cast<int>(i - k * dp);
        const float v = d < D ? __ldg(src + k * D + d) : 0.f;
        dst[i] = v;
        amax = fmaxf(amax, fabsf(v));
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(scratch + 0, __float_as_int(amax));
}

__global__ void __launch_bounds__(256)
convert_norms_kernel(const float* __restrict__ cbf, int K, int kp, int dp, __half* __restrict__ cbh,
                     float* __restrict__ cn32, double* __restrict__ cn64, int* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float sc = rows::pow2_scale_for(__int_as_float(scratch[0]));
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < kp; k += warps) {
        if (k >= K) {                                   // padding codes can never be selected
            if (lane == 0) cn32[k] = __int_as_float(0x7F800000);
            continue;
        }
        const float* c = cbf + static_cast<long long>(k) * dp;
        __half* h = cbh + static_cast<long long>(k) * dp;
        double n2 = 0.0, hat2 = 0.0, lo2 = 0.0, til2 = 0.0;
        for (int i = lane; i < dp; i += 32) {
            const float v = c[i];
            const float vs = v * sc;
            const __half hv = __float2half_rn(vs);
            const double hf = static_cast<double>(__half2float(hv));
            h[i] = hv;
            n2 += static_cast<double>(v) * v;
            hat2 += static_cast<double>(vs) * vs;
            lo2 += (static_cast<double>(vs) - hf) * (static_cast<double>(vs) - hf);
            til2 += hf * hf;
        }
        n2 = warp_sum(n2); hat2 = warp_sum(hat2); lo2 = warp_sum(lo2); til2 = warp_sum(til2);
        if (lane == 0) {
            cn64[k] = n2;
            cn32[k] = static_cast<float>(n2);
            const float up = 1.000001f;
            atomicMax(scratch + 1, __float_as_int(static_cast<float>(hat2) * up));
            atomicMax(scratch + 2, __float_as_int(static_cast<float>(lo2) * up));
            atomicMax(scratch + 3, __float_as_int(static_cast<float>(til2) * up));
            atomicMax(scratch + 4, __float_as_int(static_cast<float>(n2) * up));
        }
    }
}

// Exact duplicates. A code vector equal to an EARLIER one of its layer can never be the answer (ties go to the lowest
// index, nat.py:2157), but it sits inside every window its twin sits in: a codebook whose dead EMA entries have all
// collapsed onto the same vector (nat.py:2205-2221 divides a decayed-to-zero sum by a decayed-to-zero count) would hand
// dozens of candidates per frame to the exact re-rank, or send the frame to the exact scan. The argmin paths therefore
// read `cn32m`, a copy of ||c||^2 in which every later duplicate is +inf like a padding code; the sampling path keeps
// `cn32` (a duplicate has its own probability mass under multinomial).
__global__ void __launch_bounds__(256)
row_hash_kernel(const float* __restrict__ cbf, int K, int dp, unsigned long long* __restrict__ hash) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < K; k += warps) {
        const float* c = cbf + static_cast<long long>(k) * dp;
        unsigned long long h = 0x9E3779B97F4A7C15ull * (lane + 1);
        for (int i = lane; i < dp; i += 32) {
            unsigned int b = __float_as_uint(c[i]);
            if (b == 0x80000000u) b = 0u;                            // -0 == +0
            h = (h ^ b) * 0x100000001B3ull;
            h ^= h >> 29;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o) * 0xD6E8FEB86659FD93ull;
        h = __shfl_sync(0xffffffffu, h, 0);
        if (lane == 0) hash[k] = h;
    }
}

__global__ void __launch_bounds__(256)
mask_duplicates_kernel(const float* __restrict__ cbf, const unsigned long long* __restrict__ hash,
                       const float* __restrict__ cn32, float* __restrict__ cn32m, int K, int kp, int dp) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < kp; k += warps) {
        float out = cn32[k];
        if (k < K) {
            const unsigned long long hk = hash[k];
            const float* ck = cbf + static_cast<long long>(k) * dp;
            bool dup = false;
            for (int j0 = 0; j0 < k && !dup; j0 += 32) {
                const int j = j0 + lane;
                unsigned hit = __ballot_sync(0xffffffffu, j < k && hash[j] == hk);
                while (hit != 0 && !dup) {                          // hashes agree: compare the vectors themselves
                    const int jj = j0 + __ffs(hit) - 1;
                    hit &= hit - 1;
                    const float* cj = cbf + static_cast<long long>(jj) * dp;
                    bool same = true;
                    for (int i = lane; i < dp; i += 32) same = same && (ck[i] == cj[i]);
                    dup = __all_sync(0xffffffffu, same);
                }
            }
            if (dup) out = __int_as_float(0x7F800000);
        }
        if (lane == 0) cn32m[k] = out;
    }
}

__global__ void finish_consts_kernel(const int* __restrict__ scratch, int L, rows::LayerConst* __restrict__ lc) {
    const int l = threadIdx.x;
    if (l >= L) return;
    const int* s = scratch + l * kScratchPerLayer;
    rows::LayerConst c;
    const float up = 1.000001f;
    c.sc = rows::pow2_scale_for(__int_as_float(s[0]));
    c.chat_max = sqrtf(__int_as_float(s[1])) * up;
    c.clo_max = sqrtf(__int_as_float(s[2])) * up;
    c.ctil_max = sqrtf(__int_as_float(s[3])) * up;
    c.cmax2 = __int_as_float(s[4]);
    c.cabs = __int_as_float(s[0]);
    c.pad[0] = c.pad[1] = 0.f;
    lc[l] = c;
}

}  // namespace prepare
}  // namespace nat
