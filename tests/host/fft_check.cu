// Host-side check of the 16 x 16 x 8 transform in csrc/mel_fft.cuh: the passes are replayed serially over the
// 128 "threads" of a team and compared with a direct double-precision DFT. Built and run by tests/test_host_fft.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neural_audio_tokenizer_b200/csrc/mel_fft.cuh"

int main() {
    using namespace nat::fe;
    std::vector<float2> tw1(TW1_ELEMS), tw2(TW2_ELEMS);
    for (int k = 1; k < 16; ++k) {
        for (int t = 0; t < TEAM; ++t) tw1[(k - 1) * TEAM + t] = fft_twiddle_value((long long)t * k, NFFT);
        for (int n2 = 0; n2 < 8; ++n2) tw2[(k - 1) * 8 + n2] = fft_twiddle_value((long long)n2 * k, 128);
    }
    std::vector<float2> x(NFFT);
    srand(7);
    for (auto& v : x) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    std::vector<float2> S(FFT_BUF, make_float2(0.f, 0.f));
    for (int t = 0; t < TEAM; ++t) {
        float2 a[16];
        for (int n = 0; n < 16; ++n) a[n] = x[n * TEAM + t];
        fft_pass1(S.data(), tw1.data(), t, a);
    }
    for (int t = 0; t < TEAM; ++t) fft_pass2(S.data(), tw2.data(), t);
    // direct DFT in double precision
    std::vector<double> Xre(NFFT), Xim(NFFT);
    double max_ref = 0.0;
    for (int k = 0; k < NFFT; ++k) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n < NFFT; ++n) {
            const double a = -2.0 * M_PI * ((long long)k * n % NFFT) / NFFT;
            re += x[n].x * std::cos(a) - x[n].y * std::sin(a);
            im += x[n].x * std::sin(a) + x[n].y * std::cos(a);
        }
        Xre[k] = re; Xim[k] = im;
        max_ref = std::fmax(max_ref, std::hypot(re, im));
    }
    // pass 3 leaves every bin k <= 1024 together with its mirror N - k in the registers of exactly one thread
    double max_err = 0.0;
    std::vector<char> used(NBINS, 0);
    for (int t = 0; t < TEAM; ++t) {
        float2 a[8], b[8];
        fft_pass3(S.data(), t, a, b);
        for (int slot = 0; slot < BINS_PER_THREAD; ++slot) {
            if (slot == 8 && t != 0) continue;
            const int bin = fft_pass3_bin(t, slot);
            if (bin < 0 || bin >= NBINS || used[bin]) { printf("FAIL: bad or repeated bin %d (t=%d slot=%d)\n", bin, t, slot); return 1; }
            used[bin] = 1;
            float2 z, y;
            fft_pass3_pair(t == 0, slot, a, b, z, y);
            const int mk = (NFFT - bin) & (NFFT - 1);
            max_err = std::fmax(max_err, std::hypot(z.x - Xre[bin], z.y - Xim[bin]));
            max_err = std::fmax(max_err, std::hypot(y.x - Xre[mk], y.y - Xim[mk]));
        }
    }
    for (int k = 0; k < NBINS; ++k) if (!used[k]) { printf("FAIL: bin %d not produced\n", k); return 1; }
    // the in-register Hann window against the closed form
    double max_win = 0.0;
    for (int t = 0; t < TEAM; ++t) {
        const float sb = (float)std::sin(M_PI * t / NFFT), cb = (float)std::cos(M_PI * t / NFFT);
        for (int n = 0; n < 16; ++n) {
            const double w = 0.5 - 0.5 * std::cos(2.0 * M_PI * (n * TEAM + t) / NFFT);
            max_win = std::fmax(max_win, std::fabs(hann_sample(n, sb, cb) - w));
        }
    }
    if (!(max_win <= 2.5e-7)) { printf("FAIL: window error %.3e\n", max_win); return 1; }
    // every 8-byte exchange pattern of a half-warp (16 consecutive threads) covers 16 distinct bank pairs
    auto conflict_free = [](int (*addr)(int t, int j), int n_j, const char* what, int first_t0 = 0) {
        for (int j = 0; j < n_j; ++j)
            for (int t0 = first_t0; t0 < TEAM; t0 += 16) {
                int seen = 0;
                for (int t = t0; t < t0 + 16; ++t) {
                    const int bank = addr(t, j) & 15;
                    if (seen & (1 << bank)) { printf("FAIL: %s bank conflict (j=%d, t0=%d)\n", what, j, t0); return false; }
                    seen |= 1 << bank;
                }
            }
        return true;
    };
    bool ok = true;
    ok &= conflict_free([](int t, int k) { return fft_phys(k * 128 + t); }, 16, "pass-1 store");
    ok &= conflict_free([](int t, int n) { return fft_phys((t >> 3) * 128 + 8 * n + (t & 7)); }, 16, "pass-2 load/store");
    ok &= conflict_free([](int t, int n) { int ua, ub; fft_pass3_units(t, ua, ub); return fft_phys(ua * 8 + n); }, 8, "pass-3 load a");
    ok &= conflict_free([](int t, int n) { int ua, ub; fft_pass3_units(t, ua, ub); return fft_phys(ub * 8 + n); }, 8, "pass-3 load b");
    ok &= conflict_free([](int t, int slot) { return fft_pos(fft_pass3_bin(t, slot)); }, 8, "power-spectrum store", 16);
    // (threads 0..15 hold the self-mirrored sub-transform groups k1 = 0 and k1 = 8: their stores pair up two-way)
    ok &= conflict_free([](int t, int k) { return k * TEAM + t; }, 15, "pass-1 twiddles");
    if (!ok) return 1;
    printf("max_err %.3e max_ref %.3e rel %.3e\n", max_err, max_ref, max_err / max_ref);
    if (!(max_err <= 2e-6 * max_ref)) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
