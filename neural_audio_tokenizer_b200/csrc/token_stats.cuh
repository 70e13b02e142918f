// Token statistics over the index streams: the integer part of what the reference computes
// on the host after moving every stream off the device --
//   * token diversity  len(torch.unique(all_tokens)) / len(all_tokens)          nat.py:4913-4917, 3442-3447
//   * entropy of the pooled token distribution (torch.unique(return_counts))     nat.py:3577-3584
//   * the 2-D histogram behind the mutual information (np.histogram2d)           nat.py:3586-3637
// The kernels produce exact integer histograms; the few floating-point operations on those counts stay on the host
// in the reference's own numpy / scipy calls (token_stats.py), so the results are identical, not merely close.
#pragma once

#include "nat_common.cuh"

namespace nat {
namespace stats {

__device__ __forceinline__ long long load_code(const void* base, int dtype, long long i) {
    return dtype == NAT_CODES_I64 ? static_cast<const long long*>(base)[i]
         : dtype == NAT_CODES_I32 ? static_cast<long long>(static_cast<const int*>(base)[i])
                                  : static_cast<long long>(static_cast<const short*>(base)[i]);
}

// counts[v] += #{i : codes[i] == v} for 0 <= v < vocab; tokens outside the vocabulary are counted in *outliers.
// Shared-memory privatised when the vocabulary fits (<= 12 288 bins), global atomics otherwise.
constexpr int kHistThreads = 512;
constexpr int kHistSmemBins = 12288;

__global__ void __launch_bounds__(kHistThreads)
token_histogram_kernel(const void* __restrict__ codes, int dtype, long long n, int vocab,
                       unsigned long long* __restrict__ counts, unsigned long long* __restrict__ outliers) {
    extern __shared__ unsigned int s_bins[];
    const bool priv = vocab <= kHistSmemBins;
    if (priv) {
        for (int v = threadIdx.x; v < vocab; v += kHistThreads) s_bins[v] = 0u;
        __syncthreads();
    }
    unsigned long long bad = 0;
    const long long stride = static_cast<long long>(gridDim.x) * kHistThreads;
    for (long long i = static_cast<long long>(blockIdx.x) * kHistThreads + threadIdx.x; i < n; i += stride) {
        const long long v = load_code(codes, dtype, i);
        if (v < 0 || v >= vocab) { ++bad; continue; }
        if (priv) atomicAdd(&s_bins[v], 1u);
        else atomicAdd(&counts[v], 1ULL);
    }
    if (bad) atomicAdd(outliers, bad);
    if (priv) {
        __syncthreads();
        for (int v = threadIdx.x; v < vocab; v += kHistThreads)
            if (s_bins[v]) atomicAdd(&counts[v], static_cast<unsigned long long>(s_bins[v]));
    }
}

// np.histogram2d(a, b, bins) on integer samples: per axis, bin = searchsorted(edges, v, side='right') - 1, a sample
// equal to the last edge goes to the last bin, samples outside [edges[0], edges[-1]] are dropped
// (numpy/lib/_histograms_impl.py, histogramdd). `edges_*` are the bins + 1 float64 edges numpy itself produced.
constexpr int kJointThreads = 256;
constexpr int kJointMaxBins = 64;

__device__ __forceinline__ int bin_of(const double* edges, int bins, double v) {
    if (!(v >= edges[0]) || v > edges[bins]) return -1;
    if (v == edges[bins]) return bins - 1;
    int lo = 0, hi = bins + 1;                       // first index with edges[idx] > v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

__global__ void __launch_bounds__(kJointThreads)
joint_histogram_kernel(const void* __restrict__ a, const void* __restrict__ b, int dtype, long long n,
                       const double* __restrict__ edges_a, const double* __restrict__ edges_b, int bins,
                       unsigned long long* __restrict__ hist) {
    __shared__ double s_ea[kJointMaxBins + 1], s_eb[kJointMaxBins + 1];
    __shared__ unsigned int s_h[kJointMaxBins * kJointMaxBins];
    for (int i = threadIdx.x; i <= bins; i += kJointThreads) { s_ea[i] = edges_a[i]; s_eb[i] = edges_b[i]; }
    for (int i = threadIdx.x; i < bins * bins; i += kJointThreads) s_h[i] = 0u;
    __syncthreads();
    const long long stride = static_cast<long long>(gridDim.x) * kJointThreads;
    for (long long i = static_cast<long long>(blockIdx.x) * kJointThreads + threadIdx.x; i < n; i += stride) {
        const int ia = bin_of(s_ea, bins, static_cast<double>(load_code(a, dtype, i)));
        const int ib = bin_of(s_eb, bins, static_cast<double>(load_code(b, dtype, i)));
        if (ia >= 0 && ib >= 0) atomicAdd(&s_h[ia * bins + ib], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins * bins; i += kJointThreads)
        if (s_h[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(s_h[i]));
}

}  // namespace stats
}  // namespace nat
