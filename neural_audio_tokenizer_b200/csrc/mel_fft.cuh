// Front-end kernels: framed 2048-point real DFT with fused epilogues.
//
//   MEL mode       replaces torchaudio MelSpectrogram(n_fft=2048, hop, n_mels, normalized=True) as the reference
//                  builds it at nat.py:2281-2290: reflect padding, periodic Hann, /sum(w^2), |.|^2, banded
//                  (98.5 % sparse) HTK filterbank projection, optional 10*log10.
//   SPECTRAL mode  replaces the per-frame rfft loop of nat.py:2405-2430: |X| + 1e-12, centroid and bandwidth.
//
// Two real frames share one complex FFT (frame a in the real lane, frame b in the imaginary lane) and are separated
// afterwards with the conjugate-symmetry identities. The transform is a three-pass decimation-in-frequency
// 2048 = 16 x 16 x 8 by a team of 128 threads with 16 points per thread in registers: pass 1 reads the samples
// straight from global memory and windows them with a Hann value computed in registers, passes 2 and 3 exchange
// through a padded 17 KB shared buffer. Pass 3 gives every thread the two length-8 sub-transforms whose outputs are
// each other's mirror bins (k and N - k), so the two real spectra are separated in registers and the complex
// spectrum is never written back: only the 1025 power (or magnitude) values of each frame go to shared memory.
// Every shared-memory access of the transform is an 8-byte complex element, served per half-warp: one pad per 16
// elements makes all exchange patterns (pass-1 store, pass-2 load / store, pass-3 load, spectrum store) hit 16 distinct
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
// Every index below is fft_phys(.) of the logical element, written as a per-thread base plus a compile-time offset
// (fft_phys(128 k + t) = 136 k + t + (t >> 4), and so on), so that the accesses need no address arithmetic.
__host__ __device__ __forceinline__ void fft_pass1(float2* S, const float2* tw1, int t, float2 (&a)[16]) {
    fft16(a);
    float2* p = S + t + (t >> 4);
    const float2* w = tw1 + t;
#pragma unroll
    for (int k = 0; k < 16; ++k) p[k * 136] = k == 0 ? a[0] : cmul(a[k], w[(k - 1) * TEAM]);
}
__host__ __device__ __forceinline__ void fft_pass2(float2* S, const float2* tw2, int t) {
    const int s = t >> 3, n2 = t & 7;
    float2* p = S + s * 136 + n2;                  // fft_phys(128 s + 8 n + n2) = 136 s + n2 + 8 n + (n >> 1)
    const float2* w = tw2 + n2;
    float2 a[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) a[n] = p[8 * n + (n >> 1)];
    fft16(a);
#pragma unroll
    for (int k = 0; k < 16; ++k) p[8 * k + (k >> 1)] = k == 0 ? a[0] : cmul(a[k], w[(k - 1) * 8]);
}
// Pass 3: sub-transform u = k1 * 16 + k2 (8 consecutive elements at 8 u) yields X[base(u) + 256 k3], k3 < 8, with
// base(u) = k1 + 16 k2. A thread takes the two sub-transforms whose bases add up to 256: the mirror of bin
// base_a + 256 j is then base_b + 256 (7 - j), held by the same thread. Thread 0 takes the two self-mirrored ones
// (bases 0 and 128). For a half-warp both loads cover 16 distinct bank pairs (tests/host/fft_check.cu).
__host__ __device__ __forceinline__ void fft_pass3_units(int t, int& ua, int& ub) {
    if (t >= 16) { ua = t; ub = (16 - (t >> 4)) * 16 + 15 - (t & 15); }
    else if (t >= 8) { ua = 128 + t; ub = 128 + 15 - t; }
    else if (t > 0) { ua = t; ub = 16 - t; }
    else { ua = 0; ub = 8; }
}
__host__ __device__ __forceinline__ int fft_sub_base(int u) { return (u >> 4) + 16 * (u & 15); }
__host__ __device__ __forceinline__ void fft_pass3(const float2* S, int t, float2 (&a)[8], float2 (&b)[8]) {
    int ua, ub;
    fft_pass3_units(t, ua, ub);
    const float2* pa = S + 8 * ua + (ua >> 1);     // fft_phys(8 u + n) = 8 u + (u >> 1) + n
    const float2* pb = S + 8 * ub + (ub >> 1);
#pragma unroll
    for (int n = 0; n < 8; ++n) { a[n] = pa[n]; b[n] = pb[n]; }
    fft8(a);
    fft8(b);
}
constexpr int BINS_PER_THREAD = 9;           // slots 0..7 for every thread, slot 8 (bin 1024) for thread 0 only
// Bin (<= 1024) of a thread's slot, and the pair {X[bin], X[N - bin]} out of the pass-3 registers.
__host__ __device__ __forceinline__ int fft_pass3_bin(int t, int slot) {
    if (t == 0) return slot < 4 ? 256 * slot : (slot < 8 ? 128 + 256 * (slot - 4) : NFFT / 2);
    int ua, ub;
    fft_pass3_units(t, ua, ub);
    const int ka = fft_sub_base(ua);
    return slot < 4 ? ka + 256 * slot : 256 - ka + 256 * (slot - 4);
}
// fft_pos(fft_pass3_bin(t, slot)) without per-slot arithmetic: slots 0..3 and 4..7 are 272 apart (256 bins + 16 pads)
__host__ __device__ __forceinline__ void fft_pass3_slot_bases(int t, int& pos_a, int& pos_b) {
    pos_a = fft_pos(fft_pass3_bin(t, 0));
    pos_b = fft_pos(fft_pass3_bin(t, 4));
}
__host__ __device__ __forceinline__ void fft_pass3_pair(bool first, int slot, const float2 (&a)[8], const float2 (&b)[8],
                                                        float2& z, float2& y) {
    if (slot < 4) { z = a[slot]; y = first ? a[(8 - slot) & 7] : b[7 - slot]; }
    else if (slot < 8) { z = b[slot - 4]; y = first ? b[11 - slot] : a[11 - slot]; }
    else { z = a[4]; y = a[4]; }
}
// X_a[k], X_b[k] from Z = FFT(a + i b), z = Z[k], y = Z[N - k]:  X_a = (z + conj y) / 2,  X_b = (z - conj y) / (2i)
__host__ __device__ __forceinline__ void split_pair(float2 z, float2 y, float2& xa, float2& xb) {
    xa = make_float2(0.5f * (z.x + y.x), 0.5f * (z.y - y.y));
    xb = make_float2(0.5f * (z.y + y.y), 0.5f * (y.x - z.x));
}

// Periodic Hann value of sample 128 n + t: sin^2(pi (128 n + t) / 2048) = (sin(pi n / 16) cos b + cos(pi n / 16) sin b)^2
// with b = pi t / 2048 computed once per thread; n is a compile-time constant wherever this is called.
__host__ __device__ __forceinline__ float hann_sample(int n, float sin_b, float cos_b) {
    float sa = 0.f, ca = 1.f;
    switch (n) {
        case 0: sa = 0.f; ca = 1.f; break;
        case 1: sa = 0.19509032201612826785f; ca = 0.98078528040323044913f; break;
        case 2: sa = 0.38268343236508977173f; ca = 0.92387953251128675613f; break;
        case 3: sa = 0.55557023301960222474f; ca = 0.83146961230254523708f; break;
        case 4: sa = 0.70710678118654752440f; ca = 0.70710678118654752440f; break;
        case 5: sa = 0.83146961230254523708f; ca = 0.55557023301960222474f; break;
        case 6: sa = 0.92387953251128675613f; ca = 0.38268343236508977173f; break;
        case 7: sa = 0.98078528040323044913f; ca = 0.19509032201612826785f; break;
        case 8: sa = 1.f; ca = 0.f; break;
        case 9: sa = 0.98078528040323044913f; ca = -0.19509032201612826785f; break;
        case 10: sa = 0.92387953251128675613f; ca = -0.38268343236508977173f; break;
        case 11: sa = 0.83146961230254523708f; ca = -0.55557023301960222474f; break;
        case 12: sa = 0.70710678118654752440f; ca = -0.70710678118654752440f; break;
        case 13: sa = 0.55557023301960222474f; ca = -0.83146961230254523708f; break;
        case 14: sa = 0.38268343236508977173f; ca = -0.92387953251128675613f; break;
        default: sa = 0.19509032201612826785f; ca = -0.98078528040323044913f; break;
    }
    const float s = fmaf(sa, cos_b, ca * sin_b);
    return s * s;
}
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(TEAM) : "memory"); }

struct MelArgs {
    const float* wave;      // [B, S]
    long long S;
    long long T;            // frames per clip
    int hop;
    int n_mels;
    const float2* tw;       // [NFFT/2] exp(-2 pi i k / N), followed by the pass tables: [TW1_ELEMS], [TW2_ELEMS]
    const int* layout;      // [LAYOUT_INTS] projection layout of the filterbank (fb_layout_kernel)
    const float* wpack;     // its weights, [layout total]
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

// Transform frames a and b (b may not exist) of one clip by one team: on return every thread holds its two pass-3
// sub-transforms of Z = FFT(a + i b) in registers (fft_pass3_bin / fft_pass3_pair say which bins they are). The
// caller must pass a team barrier before it reuses S.
template <bool SPECTRAL>
__device__ __forceinline__ void fft_frame_pair(float2* S, const float2* tw1_s, const float2* tw2_s, float sin_b, float cos_b,
                                               const float* __restrict__ wave, long long len, long long start_a,
                                               bool has_b, long long start_b, int team, int t,
                                               float2 (&ra)[8], float2 (&rb)[8]) {
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
            const float wn = hann_sample(n, sin_b, cos_b);
            a[n] = make_float2(va[n] * wn, vb[n] * wn);
        }
    } else {
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const int idx = n * TEAM + t;
            const float wn = hann_sample(n, sin_b, cos_b);
            const float va = frame_sample<SPECTRAL>(wave, len, start_a + idx) * wn;
            const float vb = has_b ? frame_sample<SPECTRAL>(wave, len, start_b + idx) * wn : 0.f;
            a[n] = make_float2(va, vb);
        }
    }
    fft_pass1(S, tw1_s, t, a);
    team_sync(team);
    fft_pass2(S, tw2_s, t);
    team_sync(team);
    fft_pass3(S, t, ra, rb);
}

// Filterbank projection layout (built once per filterbank by fb_layout_kernel, read by mel_power_kernel).
//
// The projection of a frame pair is split into one task per lane: task j walks a run of consecutive bins of the power
// spectrum and accumulates it against one or two weight columns. The tasks of 32 consecutive lanes form a group that
// one warp executes in lock step; a group's weights are tap-major, wpack[group offset + 32 (c i + w) + lane] for step
// i and weight column w < c, zero outside the task's own bins, so weight reads are conflict free and the trip count
// is uniform over the warp.
//   single  (any filterbank, c = 1)   task j = band j over its non-zero bins [lo_j, hi_j)
//   dual    (c = 2)                   for filterbanks in which every bin belongs to at most two bands, and those
//           adjacent, as in the triangular banks torchaudio builds (nat.py:2281-2290): the bins are cut at the band
//           starts, task 0 = [lo_0, lo_2), task j = [lo_{j+1}, lo_{j+2}); column 0 is band j (its falling half; all
//           of band 0), column 1 the rising half of band j + 1. Every power value is then read once instead of twice
//           and out[m] = column 0 of task m + column 1 of task m - 1.
// The spectrum sits in shared memory with one pad per 16 bins, and a task may start up to 15 bins early (its extra
// leading taps have zero weights) where the group's trip count leaves room: the starts of every half-warp are chosen
// greedily so that its 16 lanes begin in different bank pairs (16 random starts collide three deep on average).
constexpr int FB_PACK_CAP = 3072;            // floats of weights staged in shared memory (HTK 128 bands, dual: 2944)
constexpr int MAX_GROUPS = MAX_MELS / 32;
constexpr int LAYOUT_GOFF = 4;               // ints: [0] groups, [1] weight columns (1 | 2), [2] total weights, [3] 0
constexpr int LAYOUT_START = LAYOUT_GOFF + MAX_GROUPS + 1;
constexpr int LAYOUT_INTS = LAYOUT_START + MAX_MELS;
// worst case (single, a dense filterbank): every group walks all NBINS bins
__host__ __device__ constexpr long long fb_wpack_capacity(int n_mels) { return static_cast<long long>((n_mels + 31) / 32) * 32 * NBINS; }

__global__ void __launch_bounds__(MAX_MELS)
fb_layout_kernel(const float* __restrict__ fbT, const int2* __restrict__ band, int n_mels, int* __restrict__ layout,
                 float* __restrict__ wpack) {
    __shared__ int bd[MAX_MELS + 2], hi_s[MAX_MELS], ts_s[MAX_MELS], te_s[MAX_MELS], slack_s[MAX_MELS], start_s[MAX_MELS];
    __shared__ int goff_s[MAX_GROUPS + 1], nstep_s[MAX_GROUPS];
    const int m = threadIdx.x, lane = m & 31, g = m >> 5;
    const int n_groups = (n_mels + 31) / 32;
    int lo = 0, hi = 0;
    if (m < n_mels) { const int2 be = band[m]; lo = be.x; hi = max(be.x, be.y); }
    bd[m] = lo; hi_s[m] = hi;
    __syncthreads();
    bool ok = true;
    if (m < n_mels) {
        ok = hi > lo;
        if (m + 1 < n_mels && bd[m + 1] < lo) ok = false;
        if (m + 2 < n_mels && hi > bd[m + 2]) ok = false;
    }
    int cols = (__syncthreads_and(ok) && n_mels >= 3) ? 2 : 1;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (m == 0 && cols == 2) {
            bd[n_mels] = max(hi_s[n_mels - 2], bd[n_mels - 1]);
            bd[n_mels + 1] = max(hi_s[n_mels - 1], bd[n_mels]);
        }
        __syncthreads();
        int ts = 0, te = 0;
        if (m < n_mels) {
            if (cols == 2) { ts = m == 0 ? bd[0] : bd[m + 1]; te = bd[m + 2]; }
            else { ts = lo; te = hi; }
        }
        int n = te - ts;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
        ts_s[m] = ts; te_s[m] = te;
        slack_s[m] = m < n_mels ? min(min(n - (te - ts), ts), 15) : -1;
        if (lane == 0 && g < n_groups) nstep_s[g] = n;
        __syncthreads();
        if (m == 0) {
            goff_s[0] = 0;
            for (int q = 0; q < n_groups; ++q) goff_s[q + 1] = goff_s[q] + nstep_s[q] * 32 * cols;
        }
        __syncthreads();
        // the dual form of a wide filterbank can need more room than the single form has been given: fall back
        if (cols == 2 && goff_s[n_groups] > fb_wpack_capacity(n_mels)) { cols = 1; __syncthreads(); continue; }
        break;
    }
    // early starts, one half-warp at a time: the lanes with the least freedom choose first
    if ((lane & 15) == 0 && g < n_groups) {
        int used[16];
        unsigned done = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) used[q] = 0;
        for (int round = 0; round < 16; ++round) {
            int pick = -1;
            for (int q = 0; q < 16; ++q)
                if (!((done >> q) & 1) && slack_s[m + q] >= 0 && (pick < 0 || slack_s[m + q] < slack_s[m + pick])) pick = q;
            if (pick < 0) break;
            done |= 1u << pick;
            int best_x = 0, best_c = 1 << 30;
            for (int x = 0; x <= slack_s[m + pick]; ++x) {
                const int c = used[fft_pos(ts_s[m + pick] - x) & 15];
                if (c < best_c) { best_c = c; best_x = x; }
            }
            // (the last step of the group must stay inside the 2048-slot buffer whatever the other bands are)
            start_s[m + pick] = min(ts_s[m + pick] - best_x, NFFT - 1 - nstep_s[g]);
            ++used[fft_pos(start_s[m + pick]) & 15];
        }
    }
    __syncthreads();
    // lanes without a band follow the first lane of their group (same address: a broadcast, never a conflict)
    if (m >= n_mels) start_s[m] = g < n_groups ? start_s[g * 32] : 0;
    __syncthreads();
    if (m == 0) { layout[0] = n_groups; layout[1] = cols; layout[2] = goff_s[n_groups]; layout[3] = 0; }
    if (m <= MAX_GROUPS) layout[LAYOUT_GOFF + m] = m <= n_groups ? goff_s[m] : goff_s[n_groups];
    layout[LAYOUT_START + m] = start_s[m];
    for (int q = 0; q < n_groups; ++q) {
        const int base = goff_s[q], cnt = goff_s[q + 1] - base;
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
            const int task = q * 32 + (j & 31), col = (j >> 5) % cols, i = (j >> 5) / cols;
            const int k = start_s[task] + i, target = task + col;
            float v = 0.f;
            if (task < n_mels && target < n_mels && k >= ts_s[task] && k < te_s[task])
                v = fbT[static_cast<long long>(target) * NBINS + k];
            wpack[base + j] = v;
        }
    }
}

// One task: `n` steps over the power spectrum from padded bin `k0`, weights tap-major with stride 32 * COLS.
template <int COLS>
__device__ __forceinline__ void project_task(const float2* pw, int k0, const float* wv, int n, float (&acc)[4]) {
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const float2 pv = pw[fft_pos(k0 + i)];
        const float w0 = wv[32 * COLS * i];
        acc[0] = fmaf(pv.x, w0, acc[0]);
        acc[1] = fmaf(pv.y, w0, acc[1]);
        if (COLS == 2) {
            const float w1 = wv[32 * COLS * i + 32];
            acc[2] = fmaf(pv.x, w1, acc[2]);
            acc[3] = fmaf(pv.y, w1, acc[3]);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 3)
mel_power_kernel(MelArgs p, long long groups_per_clip, long long total_groups) {
    extern __shared__ __align__(16) unsigned char fe_smem[];
    float2* tw1_s = reinterpret_cast<float2*>(fe_smem);                               // [TW1_ELEMS]
    float2* tw2_s = tw1_s + TW1_ELEMS;                                                // [TW2_ELEMS]
    float2* S_all = tw2_s + TW2_ELEMS;                                                // [2][FFT_BUF]
    float* wpack_s = reinterpret_cast<float*>(S_all + 2 * FFT_BUF);                   // [FB_PACK_CAP]
    int* goff = reinterpret_cast<int*>(wpack_s + FB_PACK_CAP);                        // [MAX_GROUPS + 1] offsets into wpack
    float* tile_a = reinterpret_cast<float*>(goff + MAX_GROUPS + 1);                  // [n_mels][FRAMES_PER_CTA + 1] column 0
    float* tile_b = tile_a + p.n_mels * (FRAMES_PER_CTA + 1);                         // same, column 1 (dual layouts)
    const int team = threadIdx.x / TEAM, t = threadIdx.x % TEAM;
    float2* S = S_all + team * FFT_BUF;
    // the transform buffer is dead once pass 3 has read it: the power spectra of the pair live in it afterwards,
    // interleaved {frame a, frame b} per bin (one 8-byte load serves both frames of a filterbank tap), one pad per 16
    // bins (fft_pos) so that the strided stores out of pass 3 are conflict free
    float2* pw = S;
    float sin_b, cos_b;
    sincospif(static_cast<float>(t) * (1.f / NFFT), &sin_b, &cos_b);
    int pos_a, pos_b;
    fft_pass3_slot_bases(t, pos_a, pos_b);
    const int w_total = p.layout[2];
    const bool staged = w_total <= FB_PACK_CAP;                 // wide filterbanks read their weights through L1 instead
    for (int i = threadIdx.x; i < TW1_ELEMS + TW2_ELEMS; i += THREADS) tw1_s[i] = p.tw[NFFT / 2 + i];
    if (threadIdx.x <= MAX_GROUPS) goff[threadIdx.x] = p.layout[LAYOUT_GOFF + threadIdx.x];
    if (staged) for (int i = threadIdx.x; i < w_total; i += THREADS) wpack_s[i] = p.wpack[i];
    for (int i = threadIdx.x; i < FRAMES_PER_CTA + 1; i += THREADS) tile_b[i] = 0.f;  // band 0 has no rising-half task
    __syncthreads();
    for (long long grp = blockIdx.x; grp < total_groups; grp += gridDim.x) {
        const long long b = grp / groups_per_clip;
        const long long f0 = (grp - b * groups_per_clip) * FRAMES_PER_CTA;
        const float* wave = p.wave + b * p.S;
        const int nf = static_cast<int>(min(static_cast<long long>(FRAMES_PER_CTA), p.T - f0));
        // team 0 takes frames 0..3 of the group, team 1 frames 4..7, two frames per transform
        for (int pr = team * 4; pr < team * 4 + 4; pr += 2) {
            if (pr >= nf) break;                                                      // team-uniform
            const bool has_b = pr + 1 < nf;
            float2 ra[8], rb[8];
            fft_frame_pair<false>(S, tw1_s, tw2_s, sin_b, cos_b, wave, p.S, (f0 + pr) * p.hop, has_b,
                                  (f0 + pr + 1) * p.hop, team, t, ra, rb);
            float2 pab[BINS_PER_THREAD];
#pragma unroll
            for (int q = 0; q < BINS_PER_THREAD; ++q) {
                float2 z, y, xa, xb;
                fft_pass3_pair(t == 0, q, ra, rb, z, y);
                split_pair(z, y, xa, xb);
                pab[q] = make_float2((xa.x * xa.x + xa.y * xa.y) * p.inv_wsum, (xb.x * xb.x + xb.y * xb.y) * p.inv_wsum);
            }
            team_sync(team);                                                          // every pass-3 load is done
#pragma unroll
            for (int q = 0; q < 4; ++q) { pw[pos_a + 272 * q] = pab[q]; pw[pos_b + 272 * q] = pab[q + 4]; }
            if (t == 0) pw[fft_pos(NFFT / 2)] = pab[8];
            team_sync(team);
            // filterbank projection: one task per lane, both frames of the pair per lane. Taps outside the task's own
            // bins carry zero weights and read finite values of this transform (k < 2048 keeps fft_pos(k) in the buffer).
            const int n_groups = p.layout[0], cols = p.layout[1];
            for (int m = t; m < n_groups * 32; m += TEAM) {                           // warp-uniform trip count
                const int g = m >> 5;
                const int k0 = p.layout[LAYOUT_START + m];
                const int n = (goff[g + 1] - goff[g]) / (32 * cols);
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if (staged) {
                    const float* wv = wpack_s + goff[g] + (m & 31);
                    if (cols == 2) project_task<2>(pw, k0, wv, n, acc); else project_task<1>(pw, k0, wv, n, acc);
                } else {
                    // same loops, weights through the read-only path
                    const float* wv = p.wpack + goff[g] + (m & 31);
                    for (int i = 0; i < n; ++i) {
                        const float2 pv = pw[fft_pos(k0 + i)];
                        const float w0 = __ldg(wv + 32 * cols * i);
                        acc[0] = fmaf(pv.x, w0, acc[0]);
                        acc[1] = fmaf(pv.y, w0, acc[1]);
                        if (cols == 2) {
                            const float w1 = __ldg(wv + 32 * cols * i + 32);
                            acc[2] = fmaf(pv.x, w1, acc[2]);
                            acc[3] = fmaf(pv.y, w1, acc[3]);
                        }
                    }
                }
                if (m < p.n_mels) {
                    tile_a[m * (FRAMES_PER_CTA + 1) + pr] = acc[0];
                    tile_a[m * (FRAMES_PER_CTA + 1) + pr + 1] = acc[1];       // column pr + 1 <= 8: inside the padded row
                    if (cols == 2 && m + 1 < p.n_mels) {
                        tile_b[(m + 1) * (FRAMES_PER_CTA + 1) + pr] = acc[2];
                        tile_b[(m + 1) * (FRAMES_PER_CTA + 1) + pr + 1] = acc[3];
                    }
                }
            }
            team_sync(team);
        }
        __syncthreads();
        for (int o = threadIdx.x; o < p.n_mels * FRAMES_PER_CTA; o += THREADS) {
            const int m = o / FRAMES_PER_CTA, fr = o - m * FRAMES_PER_CTA;
            if (fr < nf) {
                const long long at = (b * p.n_mels + m) * p.T + f0 + fr;
                float v = tile_a[m * (FRAMES_PER_CTA + 1) + fr];
                if (p.layout[1] == 2) v += tile_b[m * (FRAMES_PER_CTA + 1) + fr];
                p.mel[at] = v;
                if (p.logmel != nullptr) p.logmel[at] = 10.f * log10f(fmaxf(v, 1e-10f));
            }
        }
        __syncthreads();
    }
}
constexpr size_t mel_smem_bytes(int n_mels) {
    return sizeof(float2) * (TW1_ELEMS + TW2_ELEMS + 2 * FFT_BUF) + sizeof(float) * FB_PACK_CAP +
           sizeof(int) * (MAX_GROUPS + 1) + 2 * sizeof(float) * n_mels * (FRAMES_PER_CTA + 1);
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

// Sums of four values over the 128 threads of a team, in a fixed order. `sh` is this team's scratch for this reduction
// (the caller alternates between two so that one barrier per reduction is enough).
__device__ __forceinline__ void team_sum4(float (&v)[4], float4* sh, int team, int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = warp_sum(v[i]);
    if ((t & 31) == 0) sh[t >> 5] = make_float4(v[0], v[1], v[2], v[3]);
    team_sync(team);
    const float4 w0 = sh[0], w1 = sh[1], w2 = sh[2], w3 = sh[3];
    v[0] = (w0.x + w1.x) + (w2.x + w3.x);
    v[1] = (w0.y + w1.y) + (w2.y + w3.y);
    v[2] = (w0.z + w1.z) + (w2.z + w3.z);
    v[3] = (w0.w + w1.w) + (w2.w + w3.w);
}

// Centroid and bandwidth of both frames of a pair straight from the pass-3 registers: every thread owns 8 (thread 0: 9)
// of the 1025 bins of each frame, so the magnitudes never touch shared memory; two team reductions per pair.
__global__ void __launch_bounds__(THREADS, 3)
spectral_stats_kernel(SpectralArgs p) {
    extern __shared__ __align__(16) unsigned char fe_smem[];
    float2* tw1_s = reinterpret_cast<float2*>(fe_smem);                               // [TW1_ELEMS]
    float2* tw2_s = tw1_s + TW1_ELEMS;                                                // [TW2_ELEMS]
    float2* S_all = tw2_s + TW2_ELEMS;                                                // [2][FFT_BUF]
    __shared__ float4 red[2][2][4];                                                   // [team][reduction][warp]
    const int team = threadIdx.x / TEAM, t = threadIdx.x % TEAM;
    float2* S = S_all + team * FFT_BUF;
    float sin_b, cos_b;
    sincospif(static_cast<float>(t) * (1.f / NFFT), &sin_b, &cos_b);
    const float hz_a = static_cast<float>(fft_pass3_bin(t, 0)), hz_b = static_cast<float>(fft_pass3_bin(t, 4));
    for (int i = threadIdx.x; i < TW1_ELEMS + TW2_ELEMS; i += THREADS) tw1_s[i] = p.tw[NFFT / 2 + i];
    __syncthreads();
    const long long pairs = (p.T + 1) / 2;
    // one frame pair per team and iteration
    for (long long pr = static_cast<long long>(blockIdx.x) * 2 + team; pr < pairs; pr += static_cast<long long>(gridDim.x) * 2) {
        const long long fa = 2 * pr, fb = fa + 1;
        const bool has_b = fb < p.T;
        float2 ra[8], rb[8];
        fft_frame_pair<true>(S, tw1_s, tw2_s, sin_b, cos_b, p.wave, p.S, fa * p.hop, has_b, fb * p.hop, team, t, ra, rb);
        float ma[BINS_PER_THREAD], mb[BINS_PER_THREAD], hz[BINS_PER_THREAD];
#pragma unroll
        for (int q = 0; q < BINS_PER_THREAD; ++q) {
            float2 z, y, xa, xb;
            fft_pass3_pair(t == 0, q, ra, rb, z, y);
            split_pair(z, y, xa, xb);
            const bool mine = q < 8 || t == 0;                               // slot 8 (bin 1024) exists for thread 0 only
            ma[q] = mine ? sqrtf(xa.x * xa.x + xa.y * xa.y) + 1e-12f : 0.f;  // nat.py:2418
            mb[q] = mine ? sqrtf(xb.x * xb.x + xb.y * xb.y) + 1e-12f : 0.f;
            hz[q] = (q < 4 ? hz_a + 256.f * q : (q < 8 ? hz_b + 256.f * (q - 4) : static_cast<float>(NFFT / 2))) * p.bin_hz;
        }
        float v[4] = {0.f, 0.f, 0.f, 0.f};                                   // sum |X_a|, sum f |X_a|, sum |X_b|, sum f |X_b|
#pragma unroll
        for (int q = 0; q < BINS_PER_THREAD; ++q) {
            v[0] += ma[q]; v[1] = fmaf(ma[q], hz[q], v[1]);
            v[2] += mb[q]; v[3] = fmaf(mb[q], hz[q], v[3]);
        }
        team_sum4(v, red[team][0], team, t);                                 // (also: every pass-3 load of S is done)
        const float total_a = v[0] + 1e-8f, total_b = v[2] + 1e-8f;          // nat.py:2425
        const float cen_a = v[1] / total_a, cen_b = v[3] / total_b;          // nat.py:2426
        float u[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < BINS_PER_THREAD; ++q) {
            const float da = hz[q] - cen_a, db = hz[q] - cen_b;
            u[0] = fmaf(ma[q], da * da, u[0]);
            u[1] = fmaf(mb[q], db * db, u[1]);
        }
        team_sum4(u, red[team][1], team, t);
        if (t == 0) {                                                        // nat.py:2429-2430
            p.out[fa] = cen_a;
            p.out[p.T + fa] = sqrtf(u[0] / total_a);
            if (has_b) { p.out[fb] = cen_b; p.out[p.T + fb] = sqrtf(u[1] / total_b); }
        }
    }
}
constexpr size_t spectral_smem_bytes() { return sizeof(float2) * (TW1_ELEMS + TW2_ELEMS + 2 * FFT_BUF); }

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
