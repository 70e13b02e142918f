/*
 * nat_b200.h -- C ABI of the B200-native RVQ + mel/spectral hot path for defcron/neural-audio-tokenizer.
 *
 * The reference (one Python file, cited as nat.py = /root/reference/neural_audio_tokenizer.py) has no FFI or plugin
 * interface for this path; its boundary is a Python class surface.  Each entry point below
 * names the reference interface it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add; `neural_audio_tokenizer_b200/_lib.py` is that binding as shipped here.
 *
 * Conventions
 *   - every pointer whose name ends in _dev is a DEVICE pointer on the current CUDA device; *_host are host pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is stream-ordered, no call
 *     synchronises the device, none allocates device memory except the *_create functions;
 *   - return value: 0 = NAT_OK, anything else is an error code; `nat_last_error()` returns a thread-local message;
 *   - there is NO CPU implementation behind these symbols: without a CUDA device they fail with NAT_ERR_CUDA.
 */
#ifndef NAT_B200_H_
#define NAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAT_B200_ABI_VERSION 2

enum nat_status {
    NAT_OK = 0,
    NAT_ERR_INVALID_ARGUMENT = 1,   /* bad rank / dims / null pointer: the Python mirror raises ValueError first   */
    NAT_ERR_CUDA = 2,               /* a CUDA runtime or driver call failed (message has the CUDA error string)     */
    NAT_ERR_WORKSPACE = 3,          /* workspace too small for even one tile of frames                              */
    NAT_ERR_UNSUPPORTED = 4         /* shape outside the supported envelope (D > 2048, K > 65536, n_fft != 2048 ..) */
};

/* Layout of the feature tensor handed to the quantiser. */
enum nat_layout {
    NAT_LAYOUT_BCT = 0,             /* [B, D, T] contiguous, time fastest: what nat.py:3239-3240 passes             */
    NAT_LAYOUT_ROWS = 1             /* [B*T, D] contiguous, feature fastest: nat.py:2141-2142 after its transpose   */
};

/* Integer width of the emitted index streams. */
enum nat_code_dtype {
    NAT_CODES_I64 = 0,              /* torch.long, what nat.py:2157 / 2172 return                                    */
    NAT_CODES_I32 = 1,
    NAT_CODES_I16 = 2               /* K <= 32768; the width used for the multi-GPU all-gather      */
};

enum nat_rvq_flags {
    NAT_RVQ_DEFAULT = 0,
    NAT_RVQ_EXACT_SCAN = 1,         /* skip the tensor-core pass: every frame takes the exact fp64 full scan (slow;
                                       used by tests as an on-device cross-check and for tiny latency-bound calls)   */
    NAT_RVQ_SINGLE_STREAM = 2       /* keep every kernel on `stream`: by default long inputs are split over two
                                       internal streams (forked from / joined to `stream`) so the HBM-bound row
                                       kernels of one half overlap the tensor-bound GEMM of the other              */
};

const char* nat_last_error(void);
int nat_abi_version(void);

/* Number of SMs / name of the current device; for the bench's grid sizing report. */
int nat_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, size_t name_len);

/* ---------------------------------------------------------------------------------------------------------------
 * Residual vector quantiser (argmin contract; the sampling form follows below)
 * Replaces: ResidualVectorQuantizer.forward / .encode (nat.py:1358-1426) and the VectorQuantizer.forward it loops
 * over (nat.py:2119-2183), argmin branch (nat.py:2155-2157).
 * ------------------------------------------------------------------------------------------------------------- */

typedef struct nat_rvq_codebooks nat_rvq_codebooks;      /* opaque: fp32 copy, scaled fp16 copy, norms, bounds  */

/* Snapshot L codebooks [K, D] fp32 (the `codebook` buffers of nat.py:2115) into device-side derived state.
 * The caller must re-create (or call nat_rvq_codebooks_update) after mutating a codebook in place
 * (nat.py:593, 1527, 2221 all use copy_). */
int nat_rvq_codebooks_create(const float* const* codebooks_dev, int L, int K, int D, void* stream,
                             nat_rvq_codebooks** out);
int nat_rvq_codebooks_update(nat_rvq_codebooks* cb, const float* const* codebooks_dev, void* stream);
int nat_rvq_codebooks_destroy(nat_rvq_codebooks* cb);
int nat_rvq_codebooks_dims(const nat_rvq_codebooks* cb, int* L, int* K, int* D);

/* Bytes of device workspace wanted for `n_frames` frames (bounded: long inputs are processed in chunks). */
size_t nat_rvq_workspace_bytes(const nat_rvq_codebooks* cb, int64_t n_frames);

/* Per-layer counters written by nat_rvq_encode_f32 when stats_dev != NULL: uint64 [L][NAT_RVQ_STAT_FIELDS]. */
#define NAT_RVQ_STAT_FIELDS 4
enum nat_rvq_stat { NAT_STAT_CERTIFIED = 0,   /* frames decided by the tensor-core pass alone                  */
                    NAT_STAT_RERANKED = 1,    /* frames whose top candidates were re-ranked exactly            */
                    NAT_STAT_FULL_SCAN = 2,   /* frames that needed the exact full scan                        */
                    NAT_STAT_RERANK_FP64 = 3  /* re-ranked frames (counted in RERANKED too) whose fp32 score
                                                 intervals overlapped: settled in fp64 (fused stack kernel)   */ };

/* Encode B*T frames through all L layers.
 *   x_dev            fp32 features in `layout`
 *   codes_out_dev    [L, B*T] integers of `code_dtype` (frame n = b*T + t); required
 *   quantized_out_dev  optional, same layout/shape as x: sum over layers of the straight-through quantised vectors
 *                    (nat.py:2167, 1408), bit-identical op order
 *   loss_out_dev     optional, float [L]: per-layer q_latent + commitment_weight * e_latent (nat.py:2162-2164)
 *   stats_dev        optional, uint64 [L][NAT_RVQ_STAT_FIELDS], ACCUMULATED (caller zeroes)
 */
int nat_rvq_encode_f32(const nat_rvq_codebooks* cb, const float* x_dev, int layout, int64_t B, int64_t T,
                       void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                       float commitment_weight, unsigned long long* stats_dev,
                       void* workspace_dev, size_t workspace_bytes, int flags, void* stream);

/* Sampling form of the same call: the reference's DEFAULT selection mode (nat.py:2150-2154, taken whenever
 * `self.training or self.use_stochastic`): probs = softmax(-cdist / temperature); codes = multinomial(probs, 1), which
 * ATen evaluates as argmax_k probs_k / q_k with q ~ Exp(1) per (frame, code). Every code is scored exactly (fp64
 * accumulation); no tensor-core pass, so this path is for the reference's own clip sizes, not for bulk throughput.
 *   temperatures_host  float [L] on the HOST; a layer with temperature <= 0 takes the exact argmin instead
 *                      (per-layer `use_stochastic`, nat.py:2105)
 *   noise_dev          float [L, B*T, K] on the device: the Exp(1) draws of each sampling layer in the order the
 *                      reference makes them (one `empty_like(probs).exponential_(1)` per layer from torch's CPU
 *                      generator) -- codes then equal the reference's except at near-ties of probs / q; or NULL:
 *                      q comes from Philox4x32-10 keyed by philox_seed (counter: frame, code, philox_draw + layer),
 *                      equal to the reference in distribution only. In this mode the distances come from the
 *                      tensor-core pass (bulk throughput; the score matrix, chunk frames x K floats, is carved off
 *                      the tail of the workspace and the chunk shrinks until both fit: nothing is allocated)
 *                      unless flags has NAT_RVQ_EXACT_SCAN, which keeps the exact per-frame scan
 * Outputs as nat_rvq_encode_f32. */
int nat_rvq_sample_f32(const nat_rvq_codebooks* cb, const float* x_dev, int layout, int64_t B, int64_t T,
                       void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                       float commitment_weight, const float* temperatures_host, const float* noise_dev,
                       unsigned long long philox_seed, unsigned long long philox_draw,
                       void* workspace_dev, size_t workspace_bytes, int flags, void* stream);

/* Same call, timed: CUDA events around every kernel launch on `stream`, then a stream synchronise, and the summed
 * device milliseconds per kernel class in prof_ms_host[NAT_PROF_FIELDS] (bench.py's roofline leg; not a hot path). */
#define NAT_PROF_FIELDS 8
enum nat_prof { NAT_PROF_PREP = 0,        /* layout transpose + fp16 operand / bound preparation of layer 0        */
                NAT_PROF_GEMM = 1,        /* tcgen05 distance GEMM + top-4 epilogue (the dominant kernel)           */
                NAT_PROF_DECIDE = 2,      /* exact decision + residual update + next-layer operand                  */
                NAT_PROF_SCAN = 3,        /* exact full-scan path                                                   */
                NAT_PROF_LOSS = 4,
                NAT_PROF_OUTPUT = 5,      /* quantised-sum reconstruction / layout out                              */
                NAT_PROF_GEMM_LAUNCHES = 6, /* count, not ms                                                        */
                NAT_PROF_WALL = 7 };      /* first event to last event                                              */
int nat_rvq_encode_profile_f32(const nat_rvq_codebooks* cb, const float* x_dev, int layout, int64_t B, int64_t T,
                               void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                               float commitment_weight, unsigned long long* stats_dev,
                               void* workspace_dev, size_t workspace_bytes, int flags, void* stream,
                               float* prof_ms_host);

/* Kernel launches issued by this library in this process so far (bench.py's gpu_launches). */
unsigned long long nat_launch_count(void);

/* Sum of per-layer code-vector gathers. Replaces ResidualVectorQuantizer.decode / VectorQuantizer.decode
 * (nat.py:1428-1446, 2185-2203).  codes_dev: [n_code_layers, B*T] of `code_dtype`; only the first
 * min(n_code_layers, L) lists are used (nat.py:1442).  out_dev: [B, D, T] (or rows) fp32. */
int nat_rvq_decode_f32(const nat_rvq_codebooks* cb, const void* codes_dev, int code_dtype, int n_code_layers,
                       int64_t B, int64_t T, int layout, float* out_dev, void* stream);

/* Host-buffer form of nat_rvq_encode_f32 for one stack (H2D of features and D2H of indices inside).
 * x_host / codes_out_host are host pointers (pinned memory overlaps copies with compute; pageable works). Uses a
 * context private to the handle, one call at a time (serialised); nat_tokenize_host_f32 below is the general form. */
int nat_rvq_encode_host_f32(const nat_rvq_codebooks* cb, const float* x_host, int layout, int64_t B, int64_t T,
                            void* codes_out_host, int code_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * The stacks of one tokenizer in one call
 * Replaces the pair of calls `self.semantic_quantizer(...)`, `self.acoustic_quantizer(...)` at nat.py:3239-3240
 * (codes only: the form `encode` / the tokenise path needs, nat.py:1422-1426).
 * ------------------------------------------------------------------------------------------------------------- */

/* Encode B*T frames through n_stacks stacks (n_stacks <= 2: S0-S3 and A0-A3).
 *   x_dev[i]         features of stack i in `layout`; stacks given the SAME pointer share one layer-0 preparation
 *   codes_out_dev    [sum_i L_i, B*T] integers of `code_dtype`: stack 0's streams, then stack 1's
 * Stacks with equal codebook_size and input_dim (<= 1024) run in ONE persistent launch per chunk of frames; other
 * shapes, inputs of a few tiles and NAT_RVQ_EXACT_SCAN run stack after stack as nat_rvq_encode_f32 would. */
size_t nat_rvq_stacks_workspace_bytes(const nat_rvq_codebooks* const* stacks, int n_stacks, int64_t n_frames);
int nat_rvq_encode_stacks_f32(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                              int layout, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                              void* workspace_dev, size_t workspace_bytes, int flags, void* stream);
/* 1 when the stacks take the shared-preparation + persistent-kernel form for this many frames. */
int nat_rvq_stacks_fused(const nat_rvq_codebooks* const* stacks, int n_stacks, int64_t n_frames);
/* The same with the reference's time-base alignment folded into the preparation's loads. Replaces
 * `F.interpolate(features, size=T_target, mode='linear', align_corners=False)` at nat.py:3225-3236 followed by the two
 * quantiser calls: x_dev[i] is [B, D, t_in[i]] (time fastest), every stack is quantised on the common T-frame time
 * base (the caller passes T = min_i t_in[i], nat.py:3227), with the floating-point steps of ATen's CPU kernel
 * (bit-identical to the reference's CPU call). Only where nat_rvq_stacks_fused() is 1; elsewhere NAT_ERR_UNSUPPORTED
 * (align with nat_interp_linear_f32 first). */
int nat_rvq_encode_stacks_aligned_f32(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                                      const int64_t* t_in, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                                      void* workspace_dev, size_t workspace_bytes, int flags, void* stream);
/* Same call, timed per kernel class like nat_rvq_encode_profile_f32 (bench.py's roofline leg). */
int nat_rvq_encode_stacks_profile_f32(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                                      int layout, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                                      void* workspace_dev, size_t workspace_bytes, int flags, void* stream,
                                      float* prof_ms_host);

/* Host-buffer form (the end-to-end call): H2D of the features and D2H of the index streams inside.
 * A nat_host_ctx owns the device staging arena, the copy stream and the events of such calls; it serves one call at
 * a time, so concurrent callers (one per stream / thread) each create their own. Nothing is kept on the codebook
 * handles, which stay shareable across streams.
 *   x_host           features, host memory (pinned memory overlaps the copies with compute; pageable works); every
 *                    chunk is uploaded once and all stacks run on it
 *   codes_out_host   [sum_i L_i, B*T] of `code_dtype`; has landed when the call returns */
typedef struct nat_host_ctx nat_host_ctx;
int nat_host_ctx_create(nat_host_ctx** out);
int nat_host_ctx_destroy(nat_host_ctx* ctx);
int nat_tokenize_host_f32(nat_host_ctx* ctx, const nat_rvq_codebooks* const* stacks, int n_stacks, const float* x_host,
                          int layout, int64_t B, int64_t T, void* codes_out_host, int code_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Front-end
 * ------------------------------------------------------------------------------------------------------------- */

/* Mel power spectrogram. Replaces self.mel_transform(waveform) at nat.py:2290, i.e.
 * torchaudio.transforms.MelSpectrogram(sample_rate, n_fft, hop_length, n_mels, normalized=True) (nat.py:2281-2287):
 * reflect-pad n_fft/2, periodic Hann, one-sided DFT, /sum(w^2), |.|^2, HTK filterbank (f_min 0, f_max sr/2, no norm).
 *   wave_dev     [B, S] fp32, S > n_fft/2
 *   fb_dev       optional [n_fft/2+1, n_mels] fp32 filterbank to use instead of the built-in one
 *   mel_out_dev  [B, n_mels, 1 + S/hop] fp32
 *   logmel_out_dev optional, same shape: 10*log10(max(mel, 1e-10)) (an addition; no tokenise-path reference)
 * n_fft must be 2048 (the value the reference hard-codes, nat.py:2233, 2396). */
int nat_mel_power_f32(const float* wave_dev, int64_t B, int64_t S, int sample_rate, int n_fft, int hop, int n_mels,
                      const float* fb_dev, float* mel_out_dev, float* logmel_out_dev, void* stream);

/* The same transform with the filterbank prepared once (what the MelSpectrogram drop-in does: its filterbank is a
 * constant of the object): nat_mel_filterbank_prepare turns a dense [n_fft/2+1, n_mels] filterbank into the banded
 * form the kernel reads (nat_mel_filterbank_bytes(n_mels) bytes of device memory owned by the caller). The dense
 * fb_dev of nat_mel_power_f32 is converted on every call into stream-ordered scratch instead. */
size_t nat_mel_filterbank_bytes(int n_mels);
int nat_mel_filterbank_prepare(const float* fb_dev, int n_mels, void* banded_out_dev, void* stream);
int nat_mel_power_banded_f32(const float* wave_dev, int64_t B, int64_t S, int sample_rate, int n_fft, int hop,
                             int n_mels, const void* fb_banded_dev, float* mel_out_dev, float* logmel_out_dev,
                             void* stream);

/* Spectral centroid and bandwidth per frame. Replaces the STFT loop of SemanticAudioEncoder._spectral_fallback
 * (nat.py:2395-2433): frames = 1 + (S - n_fft)/hop (1 when S < n_fft), no centring, zero-padded tail.
 *   wave_dev [S] fp32;  out_dev [2, T] fp32 (row 0 centroid, row 1 bandwidth). */
int nat_spectral_stats_f32(const float* wave_dev, int64_t S, int sample_rate, int n_fft, int hop, float* out_dev,
                           void* stream);

int64_t nat_mel_num_frames(int64_t S, int hop);                       /* 1 + S/hop          (center=True)      */
int64_t nat_spectral_num_frames(int64_t S, int n_fft, int hop);       /* nat.py:2400-2403                      */

/* ---------------------------------------------------------------------------------------------------------------
 * NDJSON emission (host code)
 * ------------------------------------------------------------------------------------------------------------- */

/* Every line between the header event and the end event of the reference's NDJSON stream, byte for byte:
 * replaces the per-frame loop of StreamingProtocol.create_ndjson_stream (nat.py:4482-4513) and
 * NDJSONStreamer.create_frame (nat.py:2722-2836), plus the final flush of create_end_marker (nat.py:2843-2845).
 *   sem_codes_host / ac_codes_host   HOST index streams [n_sem][ld_frames] / [n_ac][ld_frames] of `code_dtype`
 *   rle_mode                          StreamingProtocol.rle_mode
 *   layer_is_rle                      [n_sem + n_ac] bytes, 1 where NDJSONStreamer._should_use_rle_for_layer(name)
 *                                     (nat.py:2707-2711) is true; required when rle_mode != 0
 *   text_out / len_out                malloc'ed text (lines joined by '\n', no trailing newline; empty when there
 *                                     is nothing to emit); release with nat_free_host() */
int nat_ndjson_emit_frames(const void* sem_codes_host, const void* ac_codes_host, int code_dtype, int n_sem, int n_ac,
                           int64_t ld_frames, int64_t num_frames, int sample_rate, int hop_length, int rle_mode,
                           const unsigned char* layer_is_rle, double keyframe_interval_seconds, char** text_out,
                           size_t* len_out);
void nat_free_host(void* p);

/* ---------------------------------------------------------------------------------------------------------------
 * Time-base alignment before quantisation
 * ------------------------------------------------------------------------------------------------------------- */

/* out[row, i] = linear interpolation of x[row, :] at T_out points: F.interpolate(x, size=T_out, mode='linear',
 * align_corners=False) on [B, C, T] features (nat.py:3225-3236), rows = B * C, with the floating-point steps of the
 * reference's CPU call (bit-identical to torch 2.11 CPU). */
int nat_interp_linear_f32(const float* x_dev, int64_t rows, int64_t t_in, int64_t t_out, float* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Token statistics over the index streams
 * ------------------------------------------------------------------------------------------------------------- */

/* counts_dev[v] += #{tokens == v} for 0 <= v < vocab (uint64 [vocab], ACCUMULATED: call once per stream to pool
 * layers, caller zeroes); tokens outside [0, vocab) are added to *outliers_dev. Replaces the `torch.unique` passes of
 * the diversity and entropy figures (nat.py:4913-4917, 3442-3447, 3577-3584): unique tokens = non-zero bins, counts =
 * the bins themselves. */
int nat_token_histogram(const void* codes_dev, int code_dtype, int64_t n_tokens, int vocab,
                        unsigned long long* counts_dev, unsigned long long* outliers_dev, void* stream);

/* hist_dev[i * bins + j] += #{n : a[n] in bin i, b[n] in bin j} with numpy.histogram2d's binning over the given
 * float64 edges (bins + 1 each, device memory), bins <= 64: the integer part of _calculate_mutual_information
 * (nat.py:3586-3637). uint64 [bins, bins], ACCUMULATED. */
int nat_token_joint_histogram(const void* a_dev, const void* b_dev, int code_dtype, int64_t n, const double* edges_a_dev,
                              const double* edges_b_dev, int bins, unsigned long long* hist_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Debug / validation hooks (used by tests/; not part of the drop-in surface)
 * ------------------------------------------------------------------------------------------------------------- */

/* Run only the tensor-core pass of layer `layer` on rows [N, D] and dump the raw fp32 accumulators
 * acc[n, k] = sum_d fp16(x*sx)[n,d] * fp16(c*sc)[k,d] into scores_out_dev [N, Kpad] (Kpad = K rounded up to 256),
 * plus the per-row scale sx [N] and the layer's codebook scale sc [1]. */
int nat_debug_rvq_scores(const nat_rvq_codebooks* cb, int layer, const float* rows_dev, int64_t N,
                         float* scores_out_dev, float* row_scale_out_dev, float* cb_scale_out_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* Per-CTA cycle counters of the fused stack kernel (where each warp role waits; rvq_stack_sm100.cuh DBG_*).
 * enable != 0 switches recording on for later encode calls on this handle; out_host (optional) receives the counters
 * of the launches since the last read, [min(max_ctas, n_ctas)][n_slots] uint64, and resets them. Synchronises. */
int nat_debug_stack_counters(nat_rvq_codebooks* cb, int enable, unsigned long long* out_host, int max_ctas,
                             int* n_ctas, int* n_slots);

/* ---- all-gather of the index streams over NVLink peer memory (one process per GPU; csrc_host/peer_exchange.cpp) ----------
 * Replaces the `dist.all_gather` a sharded run would issue on the [L, frames] index streams (the
 * reference itself is single-device). Every rank's [rows, col_bytes] block is written by the copy engines straight
 * into its column range of every rank's [rows, world * col_bytes] output; no kernel, no staging on the receiver.
 *   nat_peer_create      allocates this rank's three output buffers (step s uses buffer s % 3) and step counters on the
 *                        current device
 *   nat_peer_export      writes two CUDA IPC handles (128 bytes) to hand to the other ranks (any transport)
 *   nat_peer_connect     handles_all = the world's 128-byte records in rank order
 *   nat_peer_all_gather  stream-ordered: pushes `block_dev` (row pitch `block_pitch` bytes) to every rank, then makes
 *                        `stream` wait until every rank's block of this step has landed here; *gathered_out = the
 *                        output buffer of this step (valid, in stream order, until the call after next)
 * Collective: every rank must make the same sequence of all_gather calls. */
typedef struct nat_peer_ctx nat_peer_ctx;
int nat_peer_create(int world, int rank, size_t rows, size_t col_bytes, nat_peer_ctx** out);
int nat_peer_export(const nat_peer_ctx* ctx, void* handles_out_128_bytes);
int nat_peer_connect(nat_peer_ctx* ctx, const void* handles_all);
int nat_peer_all_gather(nat_peer_ctx* ctx, const void* block_dev, size_t block_pitch, void* stream, void** gathered_out);
void* nat_peer_buffer(const nat_peer_ctx* ctx, int which);
/* Teardown, in this order on every rank: nat_peer_disconnect (unmaps the peers' buffers), a barrier of the caller's,
 * nat_peer_destroy (frees this rank's buffers: nobody maps them any more). */
int nat_peer_disconnect(nat_peer_ctx* ctx);
void nat_peer_destroy(nat_peer_ctx* ctx);
/* The nat_peer_* functions keep their own thread-local message (they are a translation unit of their own). */
const char* nat_peer_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NAT_B200_H_ */
