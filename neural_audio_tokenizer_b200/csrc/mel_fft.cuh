// Front-end kernels: framed 2048-point real DFT with fused epilogues.
//
//   MEL mode       replaces torchaudio MelSpectrogram(n_fft=2048, hop, n_mels, normalized=True) as the reference
//                  builds it at nat.py:2281-2290: reflect padding, periodic Hann, /sum(w^2), |.|^2, banded
//                  (98.5 % sparse) HTK filterbank projection, optional 10*log10.
//   SPECTRAL mode  replaces the per-frame rfft loop of nat.py:2405-2430: |X| + 1e-12, centroid and bandwidth.
//
// Two real frames share one complex FFT (frame a in the real lane, frame b in the imaginary lane) and are separated
// afterwards with the conjugate-symmetry identities. The transform is a three-pass decimation-in-frequency
// 2048 = 16 x 16 x 8 by a team of 128 threads with 16 points per thread in registers: pass 1 reads the windowed
// samples straight from global memory, passes 2 and 3 exchange through a padded 17 KB shared buffer; the spectrum
// ends up in natural order with one pad per 16 elements (fft_pos).
// Every shared-memory access of the transform is an 8-byte complex element, served per half-warp: one pad per 16
// elements makes all five exchange patterns (pass-1 store, pass-2 load / store, pass-3 load / store) hit 16 distinct
// bank pairs, and the twiddles come from tables laid out by (k, thread) / (k, thread & 7), so that consecutive lanes
// read consecutive elements (indexing one e^{-2 pi i e / 2048} table by thread * k is a stride-k access: up to 8-way
// conflicts, which was half of the kernel's shared-memory wavefronts in round 1, profiles/r1_v5_mel_ncu_summary.csv).
// A CTA is two teams and owns 8 consecutive frames so that the [n_mels, T] output is written in 32-byte runs.
#pragma once

#include "nat_common.cuh"

namespace nat {
namespace fe {

constexpr int NFFT = 2048;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int TEAM = 128;                    // threads per FFT
constexpr int THREADS = 256;                 // two teams per CTA
constexpr int FRAMES_PER_CTA = 8;
constexpr int MAX_MELS = 256;
constexpr int FFT_BUF = NFFT + NFFT / 16;    // padded complex elements per transform
constexpr int TW1_ELEMS = 15 * TEAM;         // pass-1 twiddles e^{-2 pi i t k / 2048}, [k - 1][t]
constexpr int TW2_ELEMS = 15 * 8;            // pass-2 twiddles e^{-2 pi i n2 k / 128},  [k - 1][n2]

// ---- small in-register DFTs (forward, e^{-2 pi i / n}), natural order in and out
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
__host__ __device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);      // t1 - i t3
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);      // t1 + i t3
}
__host__ __device__ __forceinline__ void fft8(float2 (&a)[8]) {
    float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    o1 = make_float2((o1.x + o1.y) * h, (o1.y - o1.x) * h);      // * e^{-i pi/4}
    o2 = make_float2(o2.y, -o2.x);                               // * -i
    o3 = make_float2((o3.y - o3.x) * h, -(o3.x + o3.y) * h);     // * e^{-3 i pi/4}
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}
__host__ __device__ __forceinline__ void fft16(float2 (&a)[16]) {
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    fft8(e);
    fft8(o);
    // e^{-2 pi i k / 16}, k = 1..7
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    o[1] = cmul(o[1], make_float2(c1, -s1));
    o[2] = cmul(o[2], make_float2(h, -h));
    o[3] = cmul(o[3], make_float2(s1, -c1));
    o[4] = make_float2(o[4].y, -o[4].x);
    o[5] = cmul(o[5], make_float2(-s1, -c1));
    o[6] = cmul(o[6], make_float2(-h, -h));
    o[7] = cmul(o[7], make_float2(-c1, -s1));
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = cadd(e[k], o[k]); a[k + 8] = csub(e[k], o[k]); }
}

// Physical slot of logical element i, between the passes and afterwards: one pad per 16 elements. For a half-warp of
// consecutive threads every access pattern of the three passes then covers 16 distinct 8-byte bank pairs (checked
// exhaustively by tests/test_host_fft.py).
__host__ __device__ __forceinline__ int fft_phys(int i) { return i + (i >> 4); }
// Where X[k] lands after pass 3: the same map (natural order).
__host__ __device__ __forceinline__ int fft_pos(int k) { return k + (k >> 4); }
// Twiddle tables (built on the host in double precision, copied to shared memory once per CTA):
//   tw1[(k - 1) * 128 + t]  = e^{-2 pi i t k / 2048}   pass 1, thread t, output k = 1..15
//   tw2[(k - 1) * 8 + n2]   = e^{-2 pi i n2 k / 128}   pass 2, thread & 7 = n2, output k = 1..15
__host__ __device__ __forceinline__ float2 fft_twiddle_value(long long num, long long den) {
    const double a = -2.0 * 3.14159265358979323846 * static_cast<double>(num % den) / static_cast<double>(den);
    return make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
}

// The three passes for thread t of a team; a barrier over the team separates them. `in` holds x[128 n + t], n < 16.
__host__ __device__ __forceinline__ void fft_pass1(float2* S, const float2* tw1, int t, float2 (&a)[16]) {
    fft16(a);
#pragma unroll
    for (int k = 0; k < 16; ++k) S[fft_phys(k * 128 + t)] = k == 0 ? a[0] : cmul(a[k], tw1[(k - 1) * TEAM + t]);
}
__host__ __device__ __forceinline__ void fft_pass2(float2* S, const float2* tw2, int t) {
    const int s = t >> 3, n2 = t & 7;
    float2 a[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) a[n] = S[fft_phys(s * 128 + 8 * n + n2)];
    fft16(a);
#pragma unroll
    for (int k = 0; k < 16; ++k) S[fft_phys(s * 128 + k * 8 + n2)] = k == 0 ? a[0] : cmul(a[k], tw2[(k - 1) * 8 + n2]);
}
// Pass 3 reads two length-8 sub-transforms (u = t, t + 128), and, after a team barrier, writes them in natural order:
// sub-transform u = k1 * 16 + k2 holds X[k1 + 16 k2 + 256 k3], k3 < 8.
__host__ __device__ __forceinline__ void fft_pass3_load(const float2* S, int t, float2 (&a)[8], float2 (&b)[8]) {
#pragma unroll
    for (int n = 0; n < 8; ++n) { a[n] = S[fft_phys(t * 8 + n)]; b[n] = S[fft_phys((t + TEAM) * 8 + n)]; }
    fft8(a);
    fft8(b);
}
__host__ __device__ __forceinline__ void fft_pass3_store(float2* S, int t, const float2 (&a)[8], const float2 (&b)[8]) {
    const int ka = (t >> 4) + 16 * (t & 15), kb = ka + 8;          // u = t + 128 has k1 = (t >> 4) + 8
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) { S[fft_pos(ka + 256 * k3)] = a[k3]; S[fft_pos(kb + 256 * k3)] = b[k3]; }
}

__device__ __forceinline__ float hann_from_tw(const float2* tw, int n) {
    // periodic Hann: 0.5 - 0.5 cos(2 pi n / N); tw[k].x = cos(2 pi k / N) for k < N/2 (global table, read once per CTA)
    return n < NFFT / 2 ? 0.5f - 0.5f * __ldg(&tw[n]).x : 0.5f + 0.5f * __ldg(&tw[n - NFFT / 2]).x;
}
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(TEAM) : "memory"); }

struct MelArgs {
    const float* wave;      // [B, S]
    long long S;
    long long T;            // frames per clip
    int hop;
    int n_mels;
    const float2* tw;       // [NFFT/2] exp(-2 pi i k / N) (window), followed by the pass tables: [TW1_ELEMS], [TW2_ELEMS]
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

// Transform frames a and b (b may not exist) of one clip by one team: S ends up holding the spectrum of a + i b in
// natural order (read it through fft_pos). Ends with a team barrier.
template <bool SPECTRAL>
__device__ __forceinline__ void fft_frame_pair(float2* S, const float2* tw1_s, const float2* tw2_s, const float* win_s,
                                               const float* __restrict__ wave, long long len, long long start_a,
                                               bool has_b, long long start_b, int team, int t) {
    float2 a[16];
    const long long shift = SPECTRAL ? 0 : NFFT / 2;
    const long long last = (has_b ? start_b : start_a) - shift + NFFT - 1;
    if (has_b && start_a - shift >= 0 && last < len) {
        // both frames lie inside the clip (all but the first and last few): plain coalesced loads, all in flight
        const float* pa = wave + (start_a - shift) + t;
        const float* pb = wave + (start_b - shift) + t;
        float va[16], vb[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) { va[n] = __ldg(pa + n * TEAM); vb[n] = __ldg(pb + n * TEAM); }
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float wn = win_s[n * TEAM + t];
            a[n] = make_float2(va[n] * wn, vb[n] * wn);
        }
    } else {
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const int idx = n * TEAM + t;
            const float wn = win_s[idx];
            const float va = frame_sample<SPECTRAL>(wave, len, start_a + idx) * wn;
            const float vb = has_b ? frame_sample<SPECTRAL>(wave, len, start_b + idx) * wn : 0.f;
            a[n] = make_float2(va, vb);
        }
    }
    fft_pass1(S, tw1_s, t, a);
    team_sync(team);
    fft_pass2(S, tw2_s, t);
    team_sync(team);
    float2 ra[8], rb[8];
    fft_pass3_load(S, t, ra, rb);
    team_sync(team);
    fft_pass3_store(S, t, ra, rb);
    team_sync(team);
}

// X_a[k], X_b[k] from Z = FFT(a + i b):  X_a = (Z[k] + conj Z[N-k]) / 2,  X_b = (Z[k] - conj Z[N-k]) / (2i)
__device__ __forceinline__ void split_bins(const float2* S, int k, float2& xa, float2& xb) {
    const float2 z = S[fft_pos(k)];
    const float2 y = S[fft_pos((NFFT - k) & (NFFT - 1))];
    xa = make_float2(0.5f * (z.x + y.x), 0.5f * (z.y - y.y));
    xb = make_float2(0.5f * (z.y + y.y), 0.5f * (y.x - z.x));
}

constexpr int FB_PACK_CAP = 2304;            // non-zero filterbank weights staged in shared memory (HTK, 128 bands: ~2000)

__global__ void __launch_bounds__(THREADS, 3)
mel_power_kernel(MelArgs p, long long groups_per_clip, long long total_groups) {
    extern __shared__ __align__(16) unsigned char fe_smem[];
    float2* tw1_s = reinterpret_cast<float2*>(fe_smem);                               // [TW1_ELEMS]
    float2* tw2_s = tw1_s + TW1_ELEMS;                                                // [TW2_ELEMS]
    float2* S_all = tw2_s + TW2_ELEMS;                                                // [2][FFT_BUF]
    float* win_s = reinterpret_cast<float*>(S_all + 2 * FFT_BUF);                     // [NFFT]
    float* wpack = win_s + NFFT;                                                      // [FB_PACK_CAP]
    int* woff = reinterpret_cast<int*>(wpack + FB_PACK_CAP);                          // [MAX_MELS + 1]
    float* out_tile = reinterpret_cast<float*>(woff + MAX_MELS + 1);                  // [n_mels][FRAMES_PER_CTA + 1]
    const int team = threadIdx.x / TEAM, t = threadIdx.x % TEAM;
    float2* S = S_all + team * FFT_BUF;
    // the transform buffer is dead once the bins are split: the power spectra of the pair live in it afterwards,
    // interleaved {frame a, frame b} per bin, so that one 8-byte load serves both frames of a filterbank tap
    float2* pw = S;
    for (int i = threadIdx.x; i < TW1_ELEMS + TW2_ELEMS; i += THREADS) tw1_s[i] = p.tw[NFFT / 2 + i];
    for (int i = threadIdx.x; i < NFFT; i += THREADS) win_s[i] = hann_from_tw(p.tw, i);
    if (threadIdx.x == 0) {
        int off = 0;
        for (int m = 0; m < p.n_mels; ++m) { woff[m] = off; const int2 be = p.band[m]; off += max(0, be.y - be.x); }
        woff[p.n_mels] = off;
    }
    __syncthreads();
    const bool packed = woff[p.n_mels] <= FB_PACK_CAP;          // a dense user filterbank falls back to global reads
    if (packed) {
        for (int m = 0; m < p.n_mels; ++m) {
            const int2 be = p.band[m];
            for (int k = be.x + threadIdx.x; k < be.y; k += THREADS)
                wpack[woff[m] + k - be.x] = p.fbT[static_cast<long long>(m) * NBINS + k];
        }
    }
    __syncthreads();
    const int quad = t & 3;
    for (long long grp = blockIdx.x; grp < total_groups; grp += gridDim.x) {
        const long long b = grp / groups_per_clip;
        const long long f0 = (grp - b * groups_per_clip) * FRAMES_PER_CTA;
        const float* wave = p.wave + b * p.S;
        const int nf = static_cast<int>(min(static_cast<long long>(FRAMES_PER_CTA), p.T - f0));
        // team 0 takes frames 0..3 of the group, team 1 frames 4..7, two frames per transform
        for (int pr = team * 4; pr < team * 4 + 4; pr += 2) {
            if (pr >= nf) break;                                                      // team-uniform
            const bool has_b = pr + 1 < nf;
            fft_frame_pair<false>(S, tw1_s, tw2_s, win_s, wave, p.S, (f0 + pr) * p.hop, has_b, (f0 + pr + 1) * p.hop, team, t);
            float2 pab[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int k = t + q * TEAM;
                if (k < NBINS) {
                    float2 xa, xb;
                    split_bins(S, k, xa, xb);
                    pab[q] = make_float2((xa.x * xa.x + xa.y * xa.y) * p.inv_wsum, (xb.x * xb.x + xb.y * xb.y) * p.inv_wsum);
                }
            }
            team_sync(team);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int k = t + q * TEAM;
                if (k < NBINS) pw[k] = pab[q];
            }
            team_sync(team);
            // banded projection: four lanes per band, both frames of the pair per lane, taps interleaved over the
            // quad, fixed-order quad reduction
            for (int m = t >> 2; m < p.n_mels; m += TEAM / 4) {
                const int2 be = p.band[m];
                float acc0 = 0.f, acc1 = 0.f;
                if (packed) {
                    const float* wv = wpack + woff[m] - be.x;
                    for (int k = be.x + quad; k < be.y; k += 4) {
                        const float2 pv = pw[k];
                        const float w = wv[k];
                        acc0 = fmaf(pv.x, w, acc0);
                        acc1 = fmaf(pv.y, w, acc1);
                    }
                } else {
                    const float* fb = p.fbT + static_cast<long long>(m) * NBINS;
                    for (int k = be.x + quad; k < be.y; k += 4) {
                        const float2 pv = pw[k];
                        const float w = __ldg(fb + k);
                        acc0 = fmaf(pv.x, w, acc0);
                        acc1 = fmaf(pv.y, w, acc1);
                    }
                }
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1);
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
                if (quad == 0) {
                    out_tile[m * (FRAMES_PER_CTA + 1) + pr] = acc0;
                    out_tile[m * (FRAMES_PER_CTA + 1) + pr + 1] = acc1;       // column pr + 1 <= 8: inside the padded row
                }
            }
            team_sync(team);
        }
        __syncthreads();
        for (int o = threadIdx.x; o < p.n_mels * FRAMES_PER_CTA; o += THREADS) {
            const int m = o / FRAMES_PER_CTA, fr = o - m * FRAMES_PER_CTA;
            if (fr < nf) {
                const long long at = (b * p.n_mels + m) * p.T + f0 + fr;
                const float v = out_tile[m * (FRAMES_PER_CTA + 1) + fr];
                p.mel[at] = v;
                if (p.logmel != nullptr) p.logmel[at] = 10.f * log10f(fmaxf(v, 1e-10f));
            }
        }
        __syncthreads();
    }
}
constexpr size_t mel_smem_bytes(int n_mels) {
    return sizeof(float2) * (TW1_ELEMS + TW2_ELEMS + 2 * FFT_BUF) + sizeof(float) * (NFFT + FB_PACK_CAP) +
           sizeof(int) * (MAX_MELS + 1) + sizeof(float) * n_mels * (FRAMES_PER_CTA + 1);
}

struct SpectralArgs {
    const float* wave;      // [S]
    long long S;
    long long T;
    int hop;
    float bin_hz;           // sample_rate / NFFT
    const float2* tw;       // as MelArgs::tw
    float* out;             // [2, T]
};

// Sum over the 128 threads of a team (one team per frame pair; each half-team... see caller). Ends with team barriers.
__device__ __forceinline__ float team_sum(float v, float* sh, int team, int t) {
    v = warp_sum(v);
    if ((t & 31) == 0) sh[team * 4 + (t >> 5)] = v;
    team_sync(team);
    const float r = sh[team * 4 + 0] + sh[team * 4 + 1] + sh[team * 4 + 2] + sh[team * 4 + 3];
    team_sync(team);
    return r;
}

__global__ void __launch_bounds__(THREADS, 3)
spectral_stats_kernel(SpectralArgs p) {
    extern __shared__ __align__(16) unsigned char fe_smem[];
    float2* tw1_s = reinterpret_cast<float2*>(fe_smem);                               // [TW1_ELEMS]
    float2* tw2_s = tw1_s + TW1_ELEMS;                                                // [TW2_ELEMS]
    float2* S_all = tw2_s + TW2_ELEMS;                                                // [2][FFT_BUF]
    float* win_s = reinterpret_cast<float*>(S_all + 2 * FFT_BUF);                     // [NFFT]
    __shared__ float red[8];
    const int team = threadIdx.x / TEAM, t = threadIdx.x % TEAM;
    float2* S = S_all + team * FFT_BUF;
    float* mag0 = reinterpret_cast<float*>(S);         // the transform buffer is dead once the bins are split
    float* mag1 = mag0 + NBINS + 3;
    for (int i = threadIdx.x; i < TW1_ELEMS + TW2_ELEMS; i += THREADS) tw1_s[i] = p.tw[NFFT / 2 + i];
    for (int i = threadIdx.x; i < NFFT; i += THREADS) win_s[i] = hann_from_tw(p.tw, i);
    __syncthreads();
    const long long pairs = (p.T + 1) / 2;
    // one frame pair per team and iteration
    for (long long pr = static_cast<long long>(blockIdx.x) * 2 + team; pr < pairs; pr += static_cast<long long>(gridDim.x) * 2) {
        const long long fa = 2 * pr, fb = fa + 1;
        const bool has_b = fb < p.T;
        fft_frame_pair<true>(S, tw1_s, tw2_s, win_s, p.wave, p.S, fa * p.hop, has_b, fb * p.hop, team, t);
        float ma[9], mb[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int k = t + q * TEAM;
            if (k < NBINS) {
                float2 xa, xb;
                split_bins(S, k, xa, xb);
                ma[q] = sqrtf(xa.x * xa.x + xa.y * xa.y) + 1e-12f;          // nat.py:2418
                mb[q] = sqrtf(xb.x * xb.x + xb.y * xb.y) + 1e-12f;
            }
        }
        team_sync(team);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int k = t + q * TEAM;
            if (k < NBINS) { mag0[k] = ma[q]; mag1[k] = mb[q]; }
        }
        team_sync(team);
        for (int which = 0; which < 2; ++which) {
            if (which == 1 && !has_b) break;                               // team-uniform
            const float* mag = which ? mag1 : mag0;
            float sm = 0.f, smf = 0.f;
            for (int k = t; k < NBINS; k += TEAM) {
                const float m = mag[k];
                sm += m;
                smf = fmaf(m, static_cast<float>(k) * p.bin_hz, smf);
            }
            const float total = team_sum(sm, red, team, t) + 1e-8f;        // nat.py:2425
            const float centroid = team_sum(smf, red, team, t) / total;    // nat.py:2426
            float sv = 0.f;
            for (int k = t; k < NBINS; k += TEAM) {
                const float d = static_cast<float>(k) * p.bin_hz - centroid;
                sv = fmaf(mag[k], d * d, sv);
            }
            const float var = team_sum(sv, red, team, t) / total;          // nat.py:2429-2430
            if (t == 0) {
                const long long f = which == 0 ? fa : fb;
                p.out[f] = centroid;
                p.out[p.T + f] = sqrtf(var);
            }
        }
        team_sync(team);
    }
}
constexpr size_t spectral_smem_bytes() { return sizeof(float2) * (TW1_ELEMS + TW2_ELEMS + 2 * FFT_BUF) + sizeof(float) * NFFT; }

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
