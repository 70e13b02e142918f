// Host-side check of the 16 x 16 x 8 transform in csrc/mel_fft.cuh: the passes are replayed serially over the
// 128 "threads" of a team and compared with a direct double-precision DFT. Built and run by tests/test_host_fft.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neural_audio_tokenizer_b200/csrc/mel_fft.cuh"

int main() {
    using namespace nat::fe;
    std::vector<float2> tw(NFFT / 2);
    for (int k = 0; k < NFFT / 2; ++k) {
        const double a = -2.0 * M_PI * k / NFFT;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<float2> x(NFFT);
    srand(7);
    for (auto& v : x) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    std::vector<float2> S(FFT_BUF, make_float2(0.f, 0.f));
    for (int t = 0; t < TEAM; ++t) {
        float2 a[16];
        for (int n = 0; n < 16; ++n) a[n] = x[n * TEAM + t];
        fft_pass1(S.data(), tw.data(), t, a);
    }
    for (int t = 0; t < TEAM; ++t) fft_pass2(S.data(), tw.data(), t);
    std::vector<float2> ra(TEAM * 8), rb(TEAM * 8);
    for (int t = 0; t < TEAM; ++t) {
        float2 a[8], b[8];
        fft_pass3_load(S.data(), t, a, b);
        for (int i = 0; i < 8; ++i) { ra[t * 8 + i] = a[i]; rb[t * 8 + i] = b[i]; }
    }
    for (int t = 0; t < TEAM; ++t) {               // (a team barrier separates the loads from the stores on the device)
        float2 a[8], b[8];
        for (int i = 0; i < 8; ++i) { a[i] = ra[t * 8 + i]; b[i] = rb[t * 8 + i]; }
        fft_pass3_store(S.data(), t, a, b);
    }
    double max_err = 0.0, max_ref = 0.0;
    std::vector<char> used(FFT_BUF, 0);
    for (int k = 0; k < NFFT; ++k) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n < NFFT; ++n) {
            const double a = -2.0 * M_PI * ((long long)k * n % NFFT) / NFFT;
            re += x[n].x * std::cos(a) - x[n].y * std::sin(a);
            im += x[n].x * std::sin(a) + x[n].y * std::cos(a);
        }
        const int pos = fft_pos(k);
        if (pos < 0 || pos >= FFT_BUF || used[pos]) { printf("FAIL: bad or repeated position for k=%d\n", k); return 1; }
        used[pos] = 1;
        max_err = std::fmax(max_err, std::hypot(S[pos].x - re, S[pos].y - im));
        max_ref = std::fmax(max_ref, std::hypot(re, im));
    }
    printf("max_err %.3e max_ref %.3e rel %.3e\n", max_err, max_ref, max_err / max_ref);
    if (!(max_err <= 2e-6 * max_ref)) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
