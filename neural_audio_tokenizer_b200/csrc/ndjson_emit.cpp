// Host-side emitter of the frame / token events of the reference's NDJSON stream, byte for byte.
//
// Replaces the per-frame Python loop of StreamingProtocol.create_ndjson_stream (nat.py:4482-4513) and
// NDJSONStreamer.create_frame (nat.py:2722-2836), which reads the index streams one scalar at a time
// (`int(codes[0, i])`, nat.py:4484-4487) and tops out near 20 k frames/s. The header and end
// events stay with the reference's own json.dumps calls; this function produces every line in between, including
// the final flush of a buffered RLE event that create_end_marker (nat.py:2838-2853) prepends to the end event.
//
// What "byte for byte" needs (Python semantics restated):
//   * json.dumps(..., separators=(',', ':')) of a dict keeps insertion order; ints print in decimal;
//   * round(x, 3) is the correctly rounded 3-decimal value of the binary double (same as glibc "%.3f"), and repr()
//     of that float is the decimal with trailing zeros removed but at least one fractional digit;
//   * a buffered RLE event's "dur" is extended with `dur += frames_elapsed * frame_duration_ms` in double
//     arithmetic and printed unrounded with repr(): the shortest decimal that round-trips (std::to_chars).
//
// Every event line depends only on its own frame and the frames after it (see emit_range), so frame ranges are
// formatted by a pool of host threads (NAT_NDJSON_THREADS, default: all hardware threads) and concatenated.
#include "../../include/nat_b200.h"

#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace {

// The writers below append at a raw cursor: the caller guarantees room for the longest possible line.
struct Cursor {
    char* p;
    inline void ch(char c) { *p++ = c; }
    template <size_t N>
    inline void lit(const char (&t)[N]) { memcpy(p, t, N - 1); p += N - 1; }
    inline void str(const std::string& t) { memcpy(p, t.data(), t.size()); p += t.size(); }
    inline void integer(long long v) {
        if (v >= 0 && v < 100000) {                       // token ids and small frame indices: no division loop
            const unsigned u = static_cast<unsigned>(v);
            if (u >= 10000) *p++ = static_cast<char>('0' + u / 10000);
            if (u >= 1000) *p++ = static_cast<char>('0' + u / 1000 % 10);
            if (u >= 100) *p++ = static_cast<char>('0' + u / 100 % 10);
            if (u >= 10) *p++ = static_cast<char>('0' + u / 10 % 10);
            *p++ = static_cast<char>('0' + u % 10);
            return;
        }
        p = std::to_chars(p, p + 24, v).ptr;
    }
    // repr(round(x, 3)) for a non-negative finite x: the exactly rounded 3-decimal value of the binary double
    // (std::to_chars with a precision rounds the exact value, like glibc "%.3f"), trailing zeros removed down to
    // one fractional digit. Fast path: x * 1000 rounded to an integer is the same answer whenever x * 1000 is not
    // within 1e-3 of a tie (the product's own rounding error is below 2^-17 for x < 2^35).
    inline void round3(double x) {
        if (x >= 0.0 && x < 34359738368.0) {
            const double v = x * 1000.0;
            const double fl = __builtin_floor(v);
            const double d = v - fl;
            if (d < 0.499 || d > 0.501) {
                unsigned long long q = static_cast<unsigned long long>(d > 0.5 ? fl + 1.0 : fl);
                const unsigned frac = static_cast<unsigned>(q % 1000ull);
                integer(static_cast<long long>(q / 1000ull));
                *p++ = '.';
                *p++ = static_cast<char>('0' + frac / 100);
                if (frac % 100) {
                    *p++ = static_cast<char>('0' + frac / 10 % 10);
                    if (frac % 10) *p++ = static_cast<char>('0' + frac % 10);
                }
                return;
            }
        }
        char* b = p;
        char* e = std::to_chars(p, p + 48, x, std::chars_format::fixed, 3).ptr;
        while (e - b >= 3 && e[-1] == '0' && e[-2] != '.') --e;
        p = e;
    }
    // repr(x) for a finite double with decimal exponent in [-4, 16): shortest round-trip digits, fixed notation,
    // ".0" appended to integral values. Outside that range Python switches to exponent form; durations never get
    // there, but the general shortest form is used as a fallback rather than printing something wrong.
    inline void repr(double x) {
        const double ax = x < 0 ? -x : x;
        if (ax != 0.0 && (ax < 1e-4 || ax >= 1e16)) {
            p = std::to_chars(p, p + 48, x, std::chars_format::scientific).ptr;
            return;
        }
        char* b = p;
        p = std::to_chars(p, p + 48, x, std::chars_format::fixed).ptr;
        bool has_point = false;
        for (char* q = b; q != p; ++q) has_point |= (*q == '.');
        if (!has_point) { *p++ = '.'; *p++ = '0'; }
    }
};

template <typename T>
inline long long code_at(const void* base, long long ld, int layer, long long frame) {
    return static_cast<long long>(static_cast<const T*>(base)[layer * ld + frame]);
}

struct Job {
    const void* sem; const void* ac;
    int code_dtype, n_sem, n_ac;
    long long ld, n_frames;
    bool rle;
    const unsigned char* layer_is_rle;
    double frame_ms;
    const unsigned char* keyframe;       // [n_frames] in RLE mode, else nullptr
    std::string dur3;                    // repr(round(frame_ms, 3))
    double dur3_value;

    long long code(const void* base, int layer, long long f) const {
        switch (code_dtype) {
            case NAT_CODES_I64: return code_at<long long>(base, ld, layer, f);
            case NAT_CODES_I32: return code_at<int>(base, ld, layer, f);
            default: return code_at<short>(base, ld, layer, f);
        }
    }
    // _detect_changed_layers (nat.py:4413-4440): the previous tokens are simply those of frame f - 1, because the
    // reference runs the detection on every frame of an RLE stream (keyframes included, nat.py:4507-4508)
    bool layer_changed(int slot, long long f) const {
        if (f == 0) return true;
        return slot < n_sem ? code(sem, slot, f) != code(sem, slot, f - 1)
                            : code(ac, slot - n_sem, f) != code(ac, slot - n_sem, f - 1);
    }
    bool any_changed(long long f) const {
        for (int s = 0; s < n_sem + n_ac; ++s) if (layer_changed(s, f)) return true;
        return false;
    }
};

// The lines of every event that STARTS in [f0, f1). The reference emits a buffered RLE event only when the next event
// arrives, but always before that event's own line, so the stream is ordered by starting frame and each line depends
// on its own frame and the frames after it only (the duration bookkeeping of nat.py:2772-2831):
//   dur = round(frame_ms, 3); then every later frame without a change adds frame_ms (one addition per frame, in
//   that order); the frame that ends the run adds (f - last) * frame_ms if it is a change, nothing if it is a
//   keyframe or the end of the stream (create_end_marker flushes without extending).
// Longest possible line: fixed text < 96 bytes, fi / ts / dur < 64 each, per layer a key (<= 8) and a value (<= 20).
inline size_t max_line_bytes(int n_layers) { return 320 + static_cast<size_t>(n_layers) * 32; }

// malloc'ed, realloc-grown text (no zero fill, large blocks grow by remapping): one per worker.
struct Part {
    char* data = nullptr;
    size_t cap = 0, size = 0;
    bool failed = false;
    char& operator[](size_t i) { return data[i]; }
};

void emit_range(const Job& j, long long f0, long long f1, Part& out) {
    const int n_layers = j.n_sem + j.n_ac;
    const size_t line_cap = max_line_bytes(n_layers);
    size_t used = 0;
    bool first = true;
    auto room = [&](size_t lines_left) {               // never less than one worst-case line ahead of the cursor
        if (out.cap - used < line_cap + 1) {
            const size_t want = used + line_cap + 1 + std::min<size_t>(lines_left, 1 << 16) * 128 + out.cap / 2;
            char* q = static_cast<char*>(realloc(out.data, want));
            if (q == nullptr) { out.failed = true; return false; }
            out.data = q;
            out.cap = want;
        }
        return true;
    };
    for (long long f = f0; f < f1; ++f) {
        const double time_ms = static_cast<double>(f) * j.frame_ms;                   // nat.py:4490
        const bool key = j.rle && j.keyframe[f];
        if (!j.rle || key) {                                                        // nat.py:2746-2769
            if (!room(static_cast<size_t>(f1 - f))) return;
            Cursor c{&out[used]};
            if (!first) c.ch('\n');
            first = false;
            c.lit("{\"event\":\"frame\",\"fi\":");
            c.integer(f);
            c.lit(",\"ts\":");
            c.round3(time_ms);
            c.lit(",\"dur\":");
            c.str(j.dur3);
            c.lit(",\"S\":[");
            for (int i = 0; i < j.n_sem; ++i) { if (i) c.ch(','); c.integer(j.code(j.sem, i, f)); }
            c.lit("],\"A\":[");
            for (int i = 0; i < j.n_ac; ++i) { if (i) c.ch(','); c.integer(j.code(j.ac, i, f)); }
            c.ch(']');
            if (key) c.lit(",\"is_keyframe\":true");
            c.ch('}');
            used = static_cast<size_t>(c.p - out.data);
            continue;
        }
        if (!j.any_changed(f)) continue;                                            // extends an earlier event only
        double dur = j.dur3_value;                                                  // nat.py:2789
        long long last = f;
        for (long long g = f + 1; g < j.n_frames; ++g) {
            if (j.keyframe[g]) break;
            dur += static_cast<double>(g - last) * j.frame_ms;                      // nat.py:2776-2778 / 2826-2828
            last = g;
            if (j.any_changed(g)) break;
        }
        if (!room(static_cast<size_t>(f1 - f))) return;
        Cursor c{&out[used]};
        if (!first) c.ch('\n');
        first = false;
        c.lit("{\"event\":\"tokens\",\"fi\":");
        c.integer(f);
        c.lit(",\"ts\":");
        c.round3(time_ms);
        c.lit(",\"dur\":");
        c.repr(dur);
        for (int s = 0; s < n_layers; ++s) {                                        // nat.py:2793-2802
            if (!j.layer_is_rle[s] || !j.layer_changed(s, f)) continue;
            c.lit(",\"");
            if (s < j.n_sem) { c.ch('S'); c.integer(s); c.lit("\":"); c.integer(j.code(j.sem, s, f)); }
            else { c.ch('A'); c.integer(s - j.n_sem); c.lit("\":"); c.integer(j.code(j.ac, s - j.n_sem, f)); }
        }
        bool any = false;                                                           // nat.py:2805-2813
        for (int i = 0; i < j.n_sem; ++i) {
            if (j.layer_is_rle[i]) continue;
            if (any) c.ch(','); else c.lit(",\"S_dense\":[");
            any = true;
            c.integer(j.code(j.sem, i, f));
        }
        if (any) c.ch(']');
        any = false;
        for (int i = 0; i < j.n_ac; ++i) {
            if (j.layer_is_rle[j.n_sem + i]) continue;
            if (any) c.ch(','); else c.lit(",\"A_dense\":[");
            any = true;
            c.integer(j.code(j.ac, i, f));
        }
        if (any) c.ch(']');
        c.ch('}');
        used = static_cast<size_t>(c.p - out.data);
        out.size = used;
    }
    out.size = used;
}

int worker_count(long long n_frames) {
    const char* e = getenv("NAT_NDJSON_THREADS");
    long long want = e ? atoll(e) : static_cast<long long>(std::thread::hardware_concurrency());
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    const long long by_size = n_frames / 8192;                // a thread is not worth less than ~8 k frames
    return static_cast<int>(std::max<long long>(1, std::min(want, by_size)));
}

}  // namespace

extern "C" {

void nat_free_host(void* p) { free(p); }

int nat_ndjson_emit_frames(const void* sem_codes_host, const void* ac_codes_host, int code_dtype, int n_sem, int n_ac,
                           int64_t ld_frames, int64_t num_frames, int sample_rate, int hop_length, int rle_mode,
                           const unsigned char* layer_is_rle, double keyframe_interval_seconds, char** text_out,
                           size_t* len_out) {
    if (text_out == nullptr || len_out == nullptr) return NAT_ERR_INVALID_ARGUMENT;
    *text_out = nullptr;
    *len_out = 0;
    if (n_sem < 0 || n_ac < 0 || n_sem > 64 || n_ac > 64 || num_frames < 0 || ld_frames < num_frames ||
        sample_rate <= 0 || hop_length <= 0 || code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16)
        return NAT_ERR_INVALID_ARGUMENT;
    if (num_frames > 0 && ((n_sem > 0 && sem_codes_host == nullptr) || (n_ac > 0 && ac_codes_host == nullptr)))
        return NAT_ERR_INVALID_ARGUMENT;
    if (rle_mode && layer_is_rle == nullptr) return NAT_ERR_INVALID_ARGUMENT;

    Job j;
    j.sem = sem_codes_host; j.ac = ac_codes_host; j.code_dtype = code_dtype; j.n_sem = n_sem; j.n_ac = n_ac;
    j.ld = ld_frames; j.n_frames = num_frames; j.rle = rle_mode != 0; j.layer_is_rle = layer_is_rle;
    // NDJSONStreamer.__init__ (nat.py:2624-2627), in the same double operations
    const double frames_per_second = static_cast<double>(sample_rate) / static_cast<double>(hop_length);
    j.frame_ms = 1000.0 / frames_per_second;
    {
        char tmp[64];
        Cursor c{tmp};
        c.round3(j.frame_ms);
        j.dur3.assign(tmp, c.p);
    }
    j.dur3_value = strtod(j.dur3.c_str(), nullptr);          // round(frame_duration_ms, 3) as a double

    // the keyframe clock is the one sequential dependency (nat.py:4442-4450): a cheap pass of its own
    std::vector<unsigned char> keyframe;
    if (j.rle) {
        keyframe.assign(static_cast<size_t>(num_frames), 0);
        double last_keyframe_time = 0.0;
        for (long long f = 0; f < num_frames; ++f) {
            const double time_seconds = static_cast<double>(f) * j.frame_ms / 1000.0;
            if (time_seconds - last_keyframe_time >= keyframe_interval_seconds) {
                last_keyframe_time = time_seconds;
                keyframe[static_cast<size_t>(f)] = 1;
            }
        }
    }
    j.keyframe = j.rle ? keyframe.data() : nullptr;

    const int n_workers = worker_count(num_frames);
    std::vector<Part> parts(static_cast<size_t>(n_workers));
    auto run = [&](int w) {
        emit_range(j, num_frames * w / n_workers, num_frames * (w + 1) / n_workers, parts[w]);
    };
    if (n_workers == 1) {
        run(0);
    } else {
        std::vector<std::thread> pool;
        pool.reserve(n_workers - 1);
        for (int w = 1; w < n_workers; ++w) pool.emplace_back(run, w);
        run(0);
        for (auto& t : pool) t.join();
    }
    auto release = [&]() { for (auto& p : parts) free(p.data); };
    for (auto& p : parts) if (p.failed) { release(); return NAT_ERR_INVALID_ARGUMENT; }

    if (n_workers == 1) {                                     // hand the worker's own buffer to the caller
        Part& p = parts[0];
        if (p.data == nullptr) p.data = static_cast<char*>(malloc(1));
        if (p.data == nullptr) return NAT_ERR_INVALID_ARGUMENT;
        p.data[p.size] = 0;                                   // room() keeps at least one spare byte
        *text_out = p.data;
        *len_out = p.size;
        return NAT_OK;
    }
    size_t total = 0, non_empty = 0;
    for (auto& p : parts) if (p.size != 0) { total += p.size; ++non_empty; }
    if (non_empty > 1) total += non_empty - 1;
    char* mem = static_cast<char*>(malloc(total + 1));
    if (mem == nullptr) { release(); return NAT_ERR_INVALID_ARGUMENT; }
    size_t off = 0;
    std::vector<size_t> offs(parts.size(), 0);
    for (size_t w = 0; w < parts.size(); ++w) {
        if (parts[w].size == 0) continue;
        if (off != 0) mem[off++] = '\n';
        offs[w] = off;
        off += parts[w].size;
    }
    {
        auto copy = [&](int w) { if (parts[w].size != 0) memcpy(mem + offs[w], parts[w].data, parts[w].size); };
        std::vector<std::thread> pool;
        for (int w = 1; w < n_workers; ++w) pool.emplace_back(copy, w);
        copy(0);
        for (auto& t : pool) t.join();
    }
    release();
    mem[total] = 0;
    *text_out = mem;
    *len_out = total;
    return NAT_OK;
}

}  // extern "C"
