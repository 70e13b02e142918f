// Front-end kernels: framed 2048-point real DFT in shared memory with fused epilogues.
//
//   MEL mode       replaces torchaudio MelSpectrogram(n_fft=2048, hop, n_mels, normalized=True) as the reference
//                  builds it at nat.py:2281-2290: reflect padding, periodic Hann, /sum(w^2), |.|^2, banded
//                  (98.5 % sparse) HTK filterbank projection, optional 10*log10.
//   SPECTRAL mode  replaces the per-frame rfft loop of nat.py:2405-2430: |X| + 1e-12, centroid and bandwidth.
//
// Two real frames share one complex FFT (frame a in the real lane, frame b in the imaginary lane) and are separated
// afterwards with the conjugate-symmetry identities. The transform is an in-place radix-2 decimation-in-frequency
// pass over a 16 KB shared buffer (natural order in, bit-reversed out; the epilogue reads through __brev).
// A CTA owns 8 consecutive frames so that the [n_mels, T] output is written in 32-byte runs.
#pragma once

#include "nat_common.cuh"

namespace nat {
namespace fe {

constexpr int NFFT = 2048;
constexpr int LOG2N = 11;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int THREADS = 256;
constexpr int FRAMES_PER_CTA = 8;
constexpr int MAX_MELS = 256;

// One butterfly sweep of the DIF transform, callable from host code too (tests replay it on the CPU).
__host__ __device__ inline void dif_stage(float2* x, const float2* tw, int half, int tid, int nthreads) {
    const int tw_stride = (NFFT / 2) / half;
    for (int j = tid; j < NFFT / 2; j += nthreads) {
        const int pos = j & (half - 1);
        const int i0 = ((j - pos) << 1) + pos;
        const int i1 = i0 + half;
        const float2 u = x[i0], v = x[i1];
        const float2 w = tw[pos * tw_stride];
        const float dx = u.x - v.x, dy = u.y - v.y;
        x[i0] = make_float2(u.x + v.x, u.y + v.y);
        x[i1] = make_float2(dx * w.x - dy * w.y, dx * w.y + dy * w.x);
    }
}

__device__ __forceinline__ int brev11(int k) { return static_cast<int>(__brev(static_cast<unsigned>(k)) >> (32 - LOG2N)); }

__device__ __forceinline__ float hann_from_tw(const float2* tw, int n) {
    // periodic Hann: 0.5 - 0.5 cos(2 pi n / N); tw[k].x = cos(2 pi k / N) for k < N/2
    return n < NFFT / 2 ? 0.5f - 0.5f * tw[n].x : 0.5f + 0.5f * tw[n - NFFT / 2].x;
}

struct MelArgs {
    const float* wave;      // [B, S]
    long long S;
    long long T;            // frames per clip
    int hop;
    int n_mels;
    const float2* tw;       // [NFFT/2] exp(-2 pi i k / N)
    const float* fbT;       // [n_mels, NBINS] band-major filterbank
    const int2* band;       // [n_mels] {first bin, one past last bin}
    float* mel;             // [B, n_mels, T]
    float* logmel;          // optional
    float inv_wsum;         // 1 / sum(w^2)
};

template <bool SPECTRAL>
__device__ __forceinline__ float frame_sample(const float* __restrict__ w, long long S, long long pos) {
    if (SPECTRAL) return pos < S ? __ldg(w + pos) : 0.f;                    // zero-padded tail, nat.py:2410-2412
    long long j = pos - NFFT / 2;                                           // center=True, reflect
    if (j < 0) j = -j;
    if (j >= S) j = 2 * (S - 1) - j;
    return __ldg(w + j);
}

// Transform frames f and f+1 (f+1 may not exist) of one clip: smem x holds the bit-reversed spectrum of a + i b.
template <bool SPECTRAL>
__device__ __forceinline__ void fft_frame_pair(float2* x, const float2* tw_s, const float* __restrict__ wave,
                                               long long S, long long start_a, bool has_b, long long start_b) {
    for (int n = threadIdx.x; n < NFFT; n += THREADS) {
        const float wn = hann_from_tw(tw_s, n);
        const float a = frame_sample<SPECTRAL>(wave, S, start_a + n) * wn;
        const float b = has_b ? frame_sample<SPECTRAL>(wave, S, start_b + n) * wn : 0.f;
        x[n] = make_float2(a, b);
    }
    __syncthreads();
#pragma unroll 1
    for (int half = NFFT / 2; half >= 1; half >>= 1) {
        dif_stage(x, tw_s, half, threadIdx.x, THREADS);
        __syncthreads();
    }
}

// X_a[k], X_b[k] from Z = FFT(a + i b):  X_a = (Z[k] + conj Z[N-k]) / 2,  X_b = (Z[k] - conj Z[N-k]) / (2i)
__device__ __forceinline__ void split_bins(const float2* x, int k, float2& xa, float2& xb) {
    const float2 z = x[brev11(k)];
    const float2 y = x[brev11((NFFT - k) & (NFFT - 1))];
    xa = make_float2(0.5f * (z.x + y.x), 0.5f * (z.y - y.y));
    xb = make_float2(0.5f * (z.y + y.y), 0.5f * (y.x - z.x));
}

__global__ void __launch_bounds__(THREADS)
mel_power_kernel(MelArgs p, long long groups_per_clip, long long total_groups) {
    __shared__ float2 x[NFFT];
    __shared__ float2 tw_s[NFFT / 2];
    __shared__ float pw[2][NBINS + 3];
    __shared__ float out_tile[MAX_MELS][FRAMES_PER_CTA + 1];
    for (int i = threadIdx.x; i < NFFT / 2; i += THREADS) tw_s[i] = p.tw[i];
    __syncthreads();
    for (long long grp = blockIdx.x; grp < total_groups; grp += gridDim.x) {
        const long long b = grp / groups_per_clip;
        const long long f0 = (grp - b * groups_per_clip) * FRAMES_PER_CTA;
        const float* wave = p.wave + b * p.S;
        const int nf = static_cast<int>(min(static_cast<long long>(FRAMES_PER_CTA), p.T - f0));
        for (int pr = 0; pr < nf; pr += 2) {
            const bool has_b = pr + 1 < nf;
            fft_frame_pair<false>(x, tw_s, wave, p.S, (f0 + pr) * p.hop, has_b, (f0 + pr + 1) * p.hop);
            for (int k = threadIdx.x; k < NBINS; k += THREADS) {
                float2 xa, xb;
                split_bins(x, k, xa, xb);
                pw[0][k] = (xa.x * xa.x + xa.y * xa.y) * p.inv_wsum;
                pw[1][k] = (xb.x * xb.x + xb.y * xb.y) * p.inv_wsum;
            }
            __syncthreads();
            for (int o = threadIdx.x; o < 2 * p.n_mels; o += THREADS) {
                const int which = o / p.n_mels, m = o - which * p.n_mels;
                const int2 be = __ldg(&p.band[m]);
                const float* fb = p.fbT + static_cast<long long>(m) * NBINS;
                float acc = 0.f;
                for (int k = be.x; k < be.y; ++k) acc = fmaf(pw[which][k], __ldg(fb + k), acc);
                out_tile[m][pr + which] = acc;
            }
            __syncthreads();
        }
        for (int o = threadIdx.x; o < p.n_mels * FRAMES_PER_CTA; o += THREADS) {
            const int m = o / FRAMES_PER_CTA, fr = o - m * FRAMES_PER_CTA;
            if (fr < nf) {
                const long long at = (b * p.n_mels + m) * p.T + f0 + fr;
                const float v = out_tile[m][fr];
                p.mel[at] = v;
                if (p.logmel != nullptr) p.logmel[at] = 10.f * log10f(fmaxf(v, 1e-10f));
            }
        }
        __syncthreads();
    }
}

struct SpectralArgs {
    const float* wave;      // [S]
    long long S;
    long long T;
    int hop;
    float bin_hz;           // sample_rate / NFFT
    const float2* tw;
    float* out;             // [2, T]
};

__device__ __forceinline__ float block_sum_128(float v, float* sh, int tid128) {
    // two independent 128-thread halves (one per frame of the pair) reduce side by side
    v = warp_sum(v);
    const int half = threadIdx.x >> 7, w = (threadIdx.x >> 5) & 3;
    if ((threadIdx.x & 31) == 0) sh[half * 4 + w] = v;
    __syncthreads();
    const float r = sh[half * 4 + 0] + sh[half * 4 + 1] + sh[half * 4 + 2] + sh[half * 4 + 3];
    __syncthreads();
    (void)tid128;
    return r;
}

__global__ void __launch_bounds__(THREADS)
spectral_stats_kernel(SpectralArgs p) {
    __shared__ float2 x[NFFT];
    __shared__ float2 tw_s[NFFT / 2];
    __shared__ float mag[2][NBINS + 3];
    __shared__ float red[8];
    for (int i = threadIdx.x; i < NFFT / 2; i += THREADS) tw_s[i] = p.tw[i];
    __syncthreads();
    const long long pairs = (p.T + 1) / 2;
    for (long long pr = blockIdx.x; pr < pairs; pr += gridDim.x) {
        const long long fa = 2 * pr, fb = fa + 1;
        const bool has_b = fb < p.T;
        fft_frame_pair<true>(x, tw_s, p.wave, p.S, fa * p.hop, has_b, fb * p.hop);
        for (int k = threadIdx.x; k < NBINS; k += THREADS) {
            float2 xa, xb;
            split_bins(x, k, xa, xb);
            mag[0][k] = sqrtf(xa.x * xa.x + xa.y * xa.y) + 1e-12f;          // nat.py:2418
            mag[1][k] = sqrtf(xb.x * xb.x + xb.y * xb.y) + 1e-12f;
        }
        __syncthreads();
        const int which = threadIdx.x >> 7, t = threadIdx.x & 127;
        float sm = 0.f, smf = 0.f;
        for (int k = t; k < NBINS; k += 128) {
            const float m = mag[which][k];
            sm += m;
            smf = fmaf(m, static_cast<float>(k) * p.bin_hz, smf);
        }
        const float total = block_sum_128(sm, red, t) + 1e-8f;                // nat.py:2425
        const float centroid = block_sum_128(smf, red, t) / total;            // nat.py:2426
        float sv = 0.f;
        for (int k = t; k < NBINS; k += 128) {
            const float d = static_cast<float>(k) * p.bin_hz - centroid;
            sv = fmaf(mag[which][k], d * d, sv);
        }
        const float var = block_sum_128(sv, red, t) / total;                  // nat.py:2429-2430
        if (t == 0 && (which == 0 || has_b)) {
            const long long f = which == 0 ? fa : fb;
            p.out[f] = centroid;
            p.out[p.T + f] = sqrtf(var);
        }
        __syncthreads();
    }
}

// dense [NBINS, n_mels] filterbank -> band-major copy + per-band non-zero range (one CTA per band)
__global__ void __launch_bounds__(128)
fb_to_banded_kernel(const float* __restrict__ fb, int n_mels, float* __restrict__ fbT, int2* __restrict__ band) {
    __shared__ int lo_s, hi_s;
    const int m = blockIdx.x;
    if (threadIdx.x == 0) { lo_s = NBINS; hi_s = 0; }
    __syncthreads();
    int lo = NBINS, hi = 0;
    for (int k = threadIdx.x; k < NBINS; k += blockDim.x) {
        const float v = __ldg(fb + static_cast<long long>(k) * n_mels + m);
        fbT[static_cast<long long>(m) * NBINS + k] = v;
        if (v != 0.f) { lo = min(lo, k); hi = max(hi, k + 1); }
    }
    atomicMin(&lo_s, lo);
    atomicMax(&hi_s, hi);
    __syncthreads();
    if (threadIdx.x == 0) band[m] = make_int2(min(lo_s, hi_s), hi_s);
}

}  // namespace fe
}  // namespace nat
