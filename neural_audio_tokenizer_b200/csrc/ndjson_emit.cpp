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
#include "../../include/nat_b200.h"

#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

inline void put_int(std::string& s, long long v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    s.append(buf, r.ptr);
}

// repr(round(x, 3))
inline void put_round3(std::string& s, double x) {
    char buf[64];
    int n = snprintf(buf, sizeof buf, "%.3f", x);
    while (n > 0 && buf[n - 1] == '0' && buf[n - 2] != '.') --n;
    if (n == 2 + 0) {}                                              // (never: "%.3f" always has a '.')
    // "-0.0" can only arise from a negative input; times and durations here are non-negative
    s.append(buf, static_cast<size_t>(n));
}

// repr(x) for a finite double with decimal exponent in [-4, 16): shortest round-trip digits, fixed notation, ".0"
// appended to integral values. Outside that range Python switches to exponent form; durations never get there, but
// the general shortest form is used as a fallback rather than printing something wrong.
inline void put_repr(std::string& s, double x) {
    char buf[64];
    const double ax = x < 0 ? -x : x;
    if (ax != 0.0 && (ax < 1e-4 || ax >= 1e16)) {
        auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
        // Python: 1e+16 / 1e-05 (two-digit exponent, no trailing ".0" in the mantissa)
        s.append(buf, r.ptr);
        return;
    }
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
    bool has_point = false;
    for (char* p = buf; p != r.ptr; ++p) has_point |= (*p == '.');
    s.append(buf, r.ptr);
    if (!has_point) s.append(".0");
}

struct Buffered {
    bool live = false;
    long long fi = 0;
    double ts_ms = 0.0, dur = 0.0;
    std::vector<std::pair<int, long long>> rle;       // (layer slot: 0..n_sem-1 semantic, n_sem.. acoustic, token)
    std::vector<long long> s_dense, a_dense;
};

template <typename T>
inline long long code_at(const void* base, long long ld, int layer, long long frame) {
    return static_cast<long long>(static_cast<const T*>(base)[layer * ld + frame]);
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

    auto code = [&](const void* base, int layer, long long f) -> long long {
        switch (code_dtype) {
            case NAT_CODES_I64: return code_at<long long>(base, ld_frames, layer, f);
            case NAT_CODES_I32: return code_at<int>(base, ld_frames, layer, f);
            default: return code_at<short>(base, ld_frames, layer, f);
        }
    };

    // NDJSONStreamer.__init__ (nat.py:2624-2627), in the same double operations
    const double frames_per_second = static_cast<double>(sample_rate) / static_cast<double>(hop_length);
    const double frame_duration_ms = 1000.0 / frames_per_second;
    std::string dur3;
    put_round3(dur3, frame_duration_ms);

    std::string out;
    out.reserve(static_cast<size_t>(num_frames) * (rle_mode ? 48 : 64 + 6 * (n_sem + n_ac)) + 256);
    bool first_line = true;
    auto newline = [&]() { if (!first_line) out.push_back('\n'); first_line = false; };

    Buffered buf;
    long long last_frame_index = -1;
    double last_keyframe_time = 0.0;
    std::vector<long long> sem(n_sem), ac(n_ac), prev_sem(n_sem), prev_ac(n_ac);
    bool have_prev = false;
    std::vector<int> changed;
    changed.reserve(n_sem + n_ac);

    auto flush = [&]() {                       // _flush_buffered_event (nat.py:2713-2720) + the caller's line join
        if (!buf.live) return;
        newline();
        out.append("{\"event\":\"tokens\",\"fi\":");
        put_int(out, buf.fi);
        out.append(",\"ts\":");
        put_round3(out, buf.ts_ms);
        out.append(",\"dur\":");
        put_repr(out, buf.dur);
        for (auto& kv : buf.rle) {
            out.append(",\"");
            if (kv.first < n_sem) { out.push_back('S'); put_int(out, kv.first); }
            else { out.push_back('A'); put_int(out, kv.first - n_sem); }
            out.append("\":");
            put_int(out, kv.second);
        }
        if (!buf.s_dense.empty()) {
            out.append(",\"S_dense\":[");
            for (size_t i = 0; i < buf.s_dense.size(); ++i) { if (i) out.push_back(','); put_int(out, buf.s_dense[i]); }
            out.push_back(']');
        }
        if (!buf.a_dense.empty()) {
            out.append(",\"A_dense\":[");
            for (size_t i = 0; i < buf.a_dense.size(); ++i) { if (i) out.push_back(','); put_int(out, buf.a_dense[i]); }
            out.push_back(']');
        }
        out.push_back('}');
        buf.live = false;
    };
    auto detect_changed = [&]() {              // _detect_changed_layers (nat.py:4413-4440)
        changed.clear();
        for (int i = 0; i < n_sem; ++i) if (!have_prev || sem[i] != prev_sem[i]) changed.push_back(i);
        for (int i = 0; i < n_ac; ++i) if (!have_prev || ac[i] != prev_ac[i]) changed.push_back(n_sem + i);
        prev_sem = sem;
        prev_ac = ac;
        have_prev = true;
    };

    for (long long f = 0; f < num_frames; ++f) {
        for (int i = 0; i < n_sem; ++i) sem[i] = code(sem_codes_host, i, f);
        for (int i = 0; i < n_ac; ++i) ac[i] = code(ac_codes_host, i, f);
        const double time_ms = static_cast<double>(f) * frame_duration_ms;          // nat.py:4490
        const double time_seconds = time_ms / 1000.0;
        bool is_keyframe = false;                                                   // nat.py:4442-4450
        if (rle_mode && time_seconds - last_keyframe_time >= keyframe_interval_seconds) {
            last_keyframe_time = time_seconds;
            is_keyframe = true;
        }
        if (rle_mode && !is_keyframe) {
            detect_changed();
            if (!changed.empty()) {                                                 // nat.py:2772-2822
                if (buf.live) {
                    buf.dur += static_cast<double>(f - last_frame_index) * frame_duration_ms;
                    flush();
                }
                buf.live = true;
                buf.fi = f;
                buf.ts_ms = time_ms;
                buf.dur = strtod(dur3.c_str(), nullptr);                            // round(frame_duration_ms, 3)
                buf.rle.clear();
                buf.s_dense.clear();
                buf.a_dense.clear();
                for (int slot : changed)
                    if (layer_is_rle[slot]) buf.rle.emplace_back(slot, slot < n_sem ? sem[slot] : ac[slot - n_sem]);
                for (int i = 0; i < n_sem; ++i) if (!layer_is_rle[i]) buf.s_dense.push_back(sem[i]);
                for (int i = 0; i < n_ac; ++i) if (!layer_is_rle[n_sem + i]) buf.a_dense.push_back(ac[i]);
                last_frame_index = f;
            } else if (buf.live) {                                                  // nat.py:2823-2831
                buf.dur += static_cast<double>(f - last_frame_index) * frame_duration_ms;
                last_frame_index = f;
            }
        } else {                                                                    // nat.py:2746-2769
            flush();
            newline();
            out.append("{\"event\":\"frame\",\"fi\":");
            put_int(out, f);
            out.append(",\"ts\":");
            put_round3(out, time_ms);
            out.append(",\"dur\":");
            out.append(dur3);
            out.append(",\"S\":[");
            for (int i = 0; i < n_sem; ++i) { if (i) out.push_back(','); put_int(out, sem[i]); }
            out.append("],\"A\":[");
            for (int i = 0; i < n_ac; ++i) { if (i) out.push_back(','); put_int(out, ac[i]); }
            out.push_back(']');
            if (is_keyframe) out.append(",\"is_keyframe\":true");
            out.push_back('}');
            if (rle_mode) detect_changed();                                         // nat.py:4507-4508
        }
    }
    flush();                                                                        // create_end_marker, nat.py:2843-2845

    char* mem = static_cast<char*>(malloc(out.size() + 1));
    if (mem == nullptr) return NAT_ERR_INVALID_ARGUMENT;
    memcpy(mem, out.data(), out.size());
    mem[out.size()] = 0;
    *text_out = mem;
    *len_out = out.size();
    return NAT_OK;
}

}  // extern "C"
