// Per-frame (row) kernels of the RVQ path: operand preparation, the exact decision, the residual update, the
// exact full-scan path, reconstruction of the quantised sum, decode, and layout transposes.
//
// Reference semantics restated here (nat.py = /root/reference/neural_audio_tokenizer.py):
//   q   = codebook[j]                      nat.py:2159
//   t   = q - r ;  q_ste = r + t           nat.py:2167   (straight-through form, kept as three fp32 ops)
//   r'  = r - q_ste                        nat.py:1405
//   mse = mean(t^2); loss = mse + w * mse  nat.py:2162-2164
//   out = sum_l q_ste_l  (left to right)   nat.py:1408
//
// Exactness contract: the tensor-core pass yields, per frame, the six smallest packed scores. A frame
// is decided from them only if the proven error window excludes every code that is not a candidate; candidates
// inside the window are re-ranked with fp64 dot products on the fp32 data; frames whose last candidate is still
// inside the window take the exact full scan. Ties go to the lower index, as torch.argmin does (nat.py:2157).
#pragma once

#include "nat_common.cuh"
#include "rvq_gemm_sm100.cuh"

namespace nat {
namespace rows {

// Per-layer constants derived from a codebook (device resident, written by rvq_prepare.cuh).
struct LayerConst {
    float sc;          // power-of-two scale applied before the fp16 cast: c_hat = c * sc
    float chat_max;    // max_k ||c_hat_k||            (upper bound, rounded up)
    float clo_max;     // max_k ||c_hat_k - fp16(c_hat_k)||
    float ctil_max;    // max_k ||fp16(c_hat_k)||
    float cmax2;       // max_k ||c_k||^2  (unscaled)
    float cabs;        // max |c_kd|       (unscaled): bounds the growth of a residual row's largest element
    float pad[2];
};

constexpr float kGammaPerK = 2.384185791015625e-07f;   // 2^-22 per accumulated product: tensor-core fp32 accumulation

__device__ __forceinline__ float pow2_scale_for(float amax) {
    // sx = 2^-e with e = floor(log2(amax)), clamped so that sx stays a normal float; amax == 0 -> 1.
    if (!(amax > 0.f)) return 1.f;
    int e = static_cast<int>((__float_as_uint(amax) >> 23) & 0xFF) - 127;
    e = max(-126, min(126, e));
    return __uint_as_float(static_cast<uint32_t>(127 - e) << 23);
}

// {alpha, bias, window, sx} of one frame from its scale and the three sums over the scaled row:
//   lo2 = ||x_hat - fp16(x_hat)||^2,  xt2 = ||fp16(x_hat)||^2,  xh2 = ||x_hat||^2        (x_hat = r * sx)
// Error budget of the coarse score (DESIGN.md "Exactness"): |dot error| <= lo*chat + xt*clo + gamma*xt*ctil.
__device__ __forceinline__ float4 make_rowinfo(float sx, float lo2, float xt2, float xh2,
                                               const LayerConst* __restrict__ lc, int d_pad) {
    const float up = 1.0005f;                                  // covers the fp32 rounding of the sums
    const float inv = 1.f / (sx * lc->sc);                     // exact: both are powers of two
    const float rn2 = xh2 * up / (sx * sx);
    const float xt = sqrtf(xt2 * up), lo = sqrtf(lo2 * up);
    const float gamma = kGammaPerK * static_cast<float>(d_pad);
    const float e_dot = lo * lc->chat_max + xt * lc->clo_max + gamma * xt * lc->ctil_max;
    const float e_abs = 9.5367431640625e-07f * (rn2 + lc->cmax2);   // 2^-20: fp32 epilogue + ||c||^2 rounding
    const float E = 2.f * inv * e_dot * up + e_abs + 1e-30f;
    float4 ri;
    ri.x = -2.f * inv;                                         // alpha: score = acc * alpha + ||c||^2 + bias
    ri.y = (rn2 + E) * 1.001f;                                 // bias: keeps every shifted score >= 0
    ri.z = 2.f * E;                                            // decision window
    ri.w = sx;
    return ri;
}

// Same triple when only ||x_hat||^2 and ||x_hat - fp16(x_hat)||^2 were summed: ||fp16(x_hat)|| <= ||x_hat|| + ||lo||.
__device__ __forceinline__ float4 make_rowinfo_bound(float sx, float lo2, float xh2, const LayerConst* __restrict__ lc,
                                                     int d_pad) {
    const float up = 1.0005f;
    const float xt = (sqrtf(xh2 * up) + sqrtf(lo2 * up)) * 1.000001f;
    return make_rowinfo(sx, lo2, xt * xt, xh2, lc, d_pad);
}

// Second half of every row producer: given the fp32 row already in global memory, emit the fp16 operand row and the
// {alpha, bias, window} triple the coarse pass and the decision need. Warp-collective.
__device__ __forceinline__ void finalize_row(const float4* r4, int dp4, float amax_lane,
                                             uint2* __restrict__ a_row, float4* __restrict__ rowinfo_out,
                                             const LayerConst* __restrict__ lc, int d_pad,
                                             float* __restrict__ rowamax_out,
                                             float4* __restrict__ rowinfo_out_b = nullptr,
                                             const LayerConst* __restrict__ lc_b = nullptr) {
    const int lane = threadIdx.x & 31;
    const float amax = warp_max(amax_lane);
    if (lane == 0 && rowamax_out != nullptr) *rowamax_out = amax;
    const float sx = pow2_scale_for(amax);
    float lo2 = 0.f, xt2 = 0.f, xh2 = 0.f;
    for (int i = lane; i < dp4; i += 32) {
        const float4 v = r4[i];
        const float xs[4] = {v.x * sx, v.y * sx, v.z * sx, v.w * sx};
        __half h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            h[e] = __float2half_rn(xs[e]);
            const float hf = __half2float(h[e]);
            const float lo = xs[e] - hf;
            lo2 = fmaf(lo, lo, lo2);
            xt2 = fmaf(hf, hf, xt2);
            xh2 = fmaf(xs[e], xs[e], xh2);
        }
        uint2 packed;
        packed.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
        packed.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
        a_row[i] = packed;
    }
    lo2 = warp_sum(lo2);
    xt2 = warp_sum(xt2);
    xh2 = warp_sum(xh2);
    if (lane == 0) {
        *rowinfo_out = make_rowinfo(sx, lo2, xt2, xh2, lc, d_pad);
        if (rowinfo_out_b != nullptr) *rowinfo_out_b = make_rowinfo(sx, lo2, xt2, xh2, lc_b, d_pad);
    }
}

// ---------------------------------------------------------------------------------------------------- layer 0 prep
// x rows [n, D] (any alignment) -> r [n, Dp] fp32 (zero padded), A [n, Dp] fp16, rowinfo. One warp per frame.
__global__ void __launch_bounds__(256)
prep_rows_kernel(const float* __restrict__ x, long long x_ld, int n, int D, int dp, float* __restrict__ r,
                 __half* __restrict__ a, float4* __restrict__ rowinfo, float* __restrict__ rowamax,
                 const LayerConst* __restrict__ lc, bool in_place,
                 float4* __restrict__ rowinfo_b = nullptr, const LayerConst* __restrict__ lc_b = nullptr) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
        float* rr = r + static_cast<long long>(row) * dp;
        float amax = 0.f;
        if (in_place) {
            const float4* r4 = reinterpret_cast<const float4*>(rr);
            for (int i = lane; i < dp / 4; i += 32) {
                const float4 v = r4[i];
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        } else {
            const float* xr = x + static_cast<long long>(row) * x_ld;
            for (int i = lane; i < dp; i += 32) {
                const float v = i < D ? __ldg(xr + i) : 0.f;
                rr[i] = v;
                amax = fmaxf(amax, fabsf(v));
            }
            __syncwarp();
        }
        finalize_row(reinterpret_cast<const float4*>(rr), dp / 4, amax,
                     reinterpret_cast<uint2*>(a + static_cast<long long>(row) * dp), rowinfo + row, lc, dp,
                     rowamax != nullptr ? rowamax + row : nullptr,
                     rowinfo_b != nullptr ? rowinfo_b + row : nullptr, lc_b);
    }
}

// Fused layer-0 preparation for the [B, D, T] layout: 32 frames x all features go through shared memory once
// (coalesced 128-byte reads along time), then each warp turns 4 frames into fp32 rows, fp16 operand rows and the
// {alpha, bias, window} triple. One HBM read of x, one write of r and A. Dynamic smem: 32 * (dp + 1) floats.
constexpr int kPrepFrames = 32;
constexpr int kPrepThreads = 512;                       // 16 warps: 2 CTAs of 98 KB per SM keep 32 warps of loads in flight
__global__ void __launch_bounds__(kPrepThreads)
prep_bct_fused_kernel(const float* __restrict__ x, long long T, int D, long long n0, int n, int dp,
                      float* __restrict__ r, __half* __restrict__ a, float4* __restrict__ rowinfo,
                      float* __restrict__ rowamax, const LayerConst* __restrict__ lc,
                      float4* __restrict__ rowinfo_b = nullptr, const LayerConst* __restrict__ lc_b = nullptr,
                      long long T_in = 0, float scale = 1.f) {
    extern __shared__ float s_tile[];                   // [32][dp + 1]
    constexpr int kWarps = kPrepThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = dp + 1;
    const int f0 = blockIdx.x * kPrepFrames;
    const int f = f0 + lane;
    // T_in > 0: x is [B, D, T_in] and frame t of the T-frame time base is the two-tap linear interpolation of
    // F.interpolate(x, size=T, mode='linear', align_corners=False) (nat.py:3225-3236), in the floating-point steps of
    // interp.cuh (bit-identical to the reference's CPU call): the alignment costs no pass of its own.
    long long src = -1, src1 = -1, pitch = T;
    float l0 = 1.f, l1 = 0.f;
    if (f < n) {
        const long long g = n0 + f, b = g / T, t = g - b * T;
        if (T_in > 0) {
            const float real = fmaxf(__fmaf_rn(scale, static_cast<float>(static_cast<int>(t)) + 0.5f, -0.5f), 0.f);
            const int i0 = min(static_cast<int>(real), static_cast<int>(T_in) - 1);
            const int i1 = i0 + (i0 < T_in - 1 ? 1 : 0);
            l1 = fminf(fmaxf(__fsub_rn(real, static_cast<float>(i0)), 0.f), 1.f);
            l0 = __fsub_rn(1.f, l1);
            src = b * D * T_in + i0;
            src1 = b * D * T_in + i1;
            pitch = T_in;
        } else {
            src = b * D * T + t;
        }
    }
    // lane = frame (consecutive t: one 128-byte line per feature), warps stride over features, 16 loads in flight
    // per thread (the kernel is latency-bound otherwise: HBM needs ~30 KB in flight per SM)
    constexpr int U = 16;
    for (int d0 = warp * U; d0 < dp; d0 += kWarps * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int d = d0 + u;
            v[u] = (src >= 0 && d < D) ? __ldcs(x + src + static_cast<long long>(d) * pitch) : 0.f;
        }
        if (T_in > 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int d = d0 + u;
                if (src >= 0 && d < D) v[u] = __fmaf_rn(l0, v[u], __fmul_rn(l1, __ldcs(x + src1 + static_cast<long long>(d) * pitch)));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (d0 + u < dp) s_tile[lane * ld + d0 + u] = v[u];
    }
    __syncthreads();
    for (int fr = warp; fr < kPrepFrames; fr += kWarps) {
        const int row = f0 + fr;
        if (row >= n) break;                            // warp-uniform
        const float* src_row = s_tile + fr * ld;
        float amax = 0.f;
        for (int i = lane; i < dp; i += 32) amax = fmaxf(amax, fabsf(src_row[i]));
        amax = warp_max(amax);
        const float sx = pow2_scale_for(amax);
        float* rr = r + static_cast<long long>(row) * dp;
        __half* ar = a + static_cast<long long>(row) * dp;
        float lo2 = 0.f, xt2 = 0.f, xh2 = 0.f;
        for (int i = lane; i < dp; i += 32) {
            const float v = src_row[i];
            const float xs = v * sx;
            const __half h = __float2half_rn(xs);
            const float hf = __half2float(h);
            const float lo = xs - hf;
            lo2 = fmaf(lo, lo, lo2);
            xt2 = fmaf(hf, hf, xt2);
            xh2 = fmaf(xs, xs, xh2);
            rr[i] = v;
            ar[i] = h;
        }
        lo2 = warp_sum(lo2); xt2 = warp_sum(xt2); xh2 = warp_sum(xh2);
        if (lane == 0) {
            rowinfo[row] = make_rowinfo(sx, lo2, xt2, xh2, lc, dp);
            if (rowinfo_b != nullptr) rowinfo_b[row] = make_rowinfo(sx, lo2, xt2, xh2, lc_b, dp);   // second stack, same frames
            if (rowamax != nullptr) rowamax[row] = amax;
        }
    }
}

// [B, D, T] (time fastest) -> rows [n, Dp] for frames n0..n0+n of the flattened (b, t) index; and back.
// 32 frames x 32 features per tile through shared memory so both sides are coalesced.
__global__ void __launch_bounds__(256)
bct_to_rows_kernel(const float* __restrict__ x, long long T, int D, long long n0, int n, int dp,
                   float* __restrict__ r) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int f0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int dy = ty; dy < 32; dy += 8) {
        const int d = d0 + dy, f = f0 + tx;
        float v = 0.f;
        if (d < D && f < n) {
            const long long g = n0 + f, b = g / T, t = g - b * T;
            v = __ldg(x + (b * D + d) * T + t);
        }
        tile[dy][tx] = v;
    }
    __syncthreads();
    for (int fy = ty; fy < 32; fy += 8) {
        const int f = f0 + fy, d = d0 + tx;
        if (f < n && d < dp) r[static_cast<long long>(f) * dp + d] = tile[tx][fy];
    }
}
__global__ void __launch_bounds__(256)
rows_to_bct_kernel(const float* __restrict__ r, int dp, long long T, int D, long long n0, int n,
                   float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int f0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int fy = ty; fy < 32; fy += 8) {
        const int f = f0 + fy, d = d0 + tx;
        tile[fy][tx] = (f < n && d < D) ? r[static_cast<long long>(f) * dp + d] : 0.f;
    }
    __syncthreads();
    for (int dy = ty; dy < 32; dy += 8) {
        const int d = d0 + dy, f = f0 + tx;
        if (d < D && f < n) {
            const long long g = n0 + f, b = g / T, t = g - b * T;
            out[(b * D + d) * T + t] = tile[tx][dy];
        }
    }
}

// ---------------------------------------------------------------------------------------------------- decision
__device__ __forceinline__ void store_code(void* codes, int dtype, long long pos, int j) {
    if (dtype == 0) reinterpret_cast<long long*>(codes)[pos] = j;
    else if (dtype == 1) reinterpret_cast<int*>(codes)[pos] = j;
    else reinterpret_cast<short*>(codes)[pos] = static_cast<short>(j);
}
__device__ __forceinline__ int load_code(const void* codes, int dtype, long long pos) {
    if (dtype == 0) return static_cast<int>(reinterpret_cast<const long long*>(codes)[pos]);
    if (dtype == 1) return reinterpret_cast<const int*>(codes)[pos];
    return static_cast<int>(reinterpret_cast<const unsigned short*>(codes)[pos]);
}

// Exact score of code k for the fp32 row r: ||c_k||^2 - 2 r.c_k with fp64 accumulation. Warp-collective.
__device__ __forceinline__ double exact_score(const float4* r4, const float4* __restrict__ c4, int dp4,
                                              double cn64) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int i = lane; i < dp4; i += 32) {
        const float4 a = r4[i];
        const float4 b = __ldg(c4 + i);
        acc = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc);
        acc = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc);
        acc = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc);
        acc = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc);
    }
    acc = warp_sum(acc);
    return cn64 - 2.0 * acc;
}

struct UpdateArgs {
    float* r;                 // [n, Dp] residual, updated in place
    __half* a;                // [n, Dp] fp16 operand of the NEXT layer
    float4* rowinfo;          // per frame, read (this layer's window) then overwritten (next layer)
    float* rowamax;           // per frame max |r| (kept current for the fused kernel's scale bound), or nullptr
    const float* cb;          // this layer's codebook [K, Dp] fp32 (zero padded)
    const double* cn64;       // [K] ||c_k||^2 in fp64
    const LayerConst* lc_next;   // constants of the next layer, nullptr on the last
    void* codes;              // this layer's index stream, offset to the chunk
    double* row_loss;         // [n] sum_d t^2 per frame, or nullptr
    unsigned long long* stats;   // this layer's counters or nullptr
    int n, K, dp, code_dtype;
};

// Apply code j to frame `row`: residual update in the reference's op order, loss term, next operand. Warp-collective.
__device__ __forceinline__ void apply_code(const UpdateArgs& p, int row, int j) {
    const int lane = threadIdx.x & 31;
    const int dp4 = p.dp >> 2;
    if (p.lc_next == nullptr && p.row_loss == nullptr) {      // last layer, codes only: the residual is dead
        if (lane == 0) store_code(p.codes, p.code_dtype, row, j);
        return;
    }
    float4* r4 = reinterpret_cast<float4*>(p.r + static_cast<long long>(row) * p.dp);
    const float4* c4 = reinterpret_cast<const float4*>(p.cb + static_cast<long long>(j) * p.dp);
    float amax = 0.f;
    double loss = 0.0;
    for (int i = lane; i < dp4; i += 32) {
        const float4 rv = r4[i];
        const float4 cv = __ldg(c4 + i);
        float4 nr;
        float t, q;
        t = __fsub_rn(cv.x, rv.x); q = __fadd_rn(rv.x, t); nr.x = __fsub_rn(rv.x, q); loss += static_cast<double>(__fmul_rn(t, t));
        t = __fsub_rn(cv.y, rv.y); q = __fadd_rn(rv.y, t); nr.y = __fsub_rn(rv.y, q); loss += static_cast<double>(__fmul_rn(t, t));
        t = __fsub_rn(cv.z, rv.z); q = __fadd_rn(rv.z, t); nr.z = __fsub_rn(rv.z, q); loss += static_cast<double>(__fmul_rn(t, t));
        t = __fsub_rn(cv.w, rv.w); q = __fadd_rn(rv.w, t); nr.w = __fsub_rn(rv.w, q); loss += static_cast<double>(__fmul_rn(t, t));
        r4[i] = nr;
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(nr.x), fabsf(nr.y)), fmaxf(fabsf(nr.z), fabsf(nr.w))));
    }
    if (p.row_loss != nullptr) {
        loss = warp_sum(loss);
        if (lane == 0) p.row_loss[row] = loss;
    }
    if (lane == 0) store_code(p.codes, p.code_dtype, row, j);
    if (p.lc_next != nullptr) {
        __syncwarp();
        finalize_row(r4, dp4, amax, reinterpret_cast<uint2*>(p.a + static_cast<long long>(row) * p.dp),
                     p.rowinfo + row, p.lc_next, p.dp, p.rowamax != nullptr ? p.rowamax + row : nullptr);
    }
}

// One warp per frame: decide from the coarse candidates, or defer the frame to the full-scan list.
__global__ void __launch_bounds__(256)
decide_update_kernel(UpdateArgs p, const gemm::Cand* __restrict__ cand, int* __restrict__ scan_list,
                     int* __restrict__ scan_count) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int dp4 = p.dp >> 2;
    unsigned long long n_cert = 0, n_rerank = 0, n_scan = 0;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.n; row += warps) {
        const gemm::Cand c = cand[row];
        const float window = p.rowinfo[row].z;
        constexpr int NC = gemm::NCAND;
        float f[NC];
        bool valid[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            f[i] = __int_as_float(c.key[i] & 0xFFFFFF00);
            valid[i] = c.key[i] != gemm::KEY_INVALID && c.idx[i] < p.K && fabsf(f[i]) < 3.0e38f;
        }
        // 8 low bits of every key were replaced by the column: true shifted score lies in [f, f + |f| 2^-15].
        const float thr = f[0] + fabsf(f[0]) * 6.103515625e-05f + window;
        int nwin = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) nwin += (valid[i] && f[i] <= thr) ? 1 : 0;
        int j = c.idx[0];
        if (!valid[0] || (nwin == NC && p.K > NC)) {
            // the last kept candidate is still inside the window: codes we did not keep may matter -> exact full scan
            if (lane == 0) scan_list[atomicAdd(scan_count, 1)] = row;
            ++n_scan;
            continue;
        }
        if (nwin > 1) {
            const float4* r4 = reinterpret_cast<const float4*>(p.r + static_cast<long long>(row) * p.dp);
            double best = 0.0;
            int bestj = -1;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                if (!(valid[i] && f[i] <= thr)) continue;          // warp-uniform
                const int k = c.idx[i];
                const double s = exact_score(r4, reinterpret_cast<const float4*>(p.cb + static_cast<long long>(k) * p.dp),
                                             dp4, p.cn64[k]);
                if (bestj < 0 || s < best || (s == best && k < bestj)) { best = s; bestj = k; }
            }
            j = bestj;
            ++n_rerank;
        } else {
            ++n_cert;
        }
        apply_code(p, row, j);
    }
    if (p.stats != nullptr && lane == 0) {
        if (n_cert) atomicAdd(p.stats + 0, n_cert);
        if (n_rerank) atomicAdd(p.stats + 1, n_rerank);
        if (n_scan) atomicAdd(p.stats + 2, n_scan);
    }
}

// Exact full scan: one CTA of 16 warps per listed frame; scan_list == nullptr means "every frame 0..count".
// The frame sits in shared memory; each warp scores four codes per sweep so 24+ independent 16-byte loads are in
// flight per lane (the scan is L2-latency bound, not flop bound).
constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanCodes = 4;

__global__ void __launch_bounds__(kScanThreads)
full_scan_kernel(UpdateArgs p, const int* __restrict__ scan_list, const int* __restrict__ scan_count,
                 int count_if_all, bool count_stats) {
    extern __shared__ float4 s_row[];                   // [dp / 4]
    __shared__ double s_best[kScanWarps];
    __shared__ int s_idx[kScanWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dp4 = p.dp >> 2;
    const int count = scan_list != nullptr ? *scan_count : count_if_all;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int row = scan_list != nullptr ? scan_list[e] : e;
        const float4* r4 = reinterpret_cast<const float4*>(p.r + static_cast<long long>(row) * p.dp);
        for (int i = threadIdx.x; i < dp4; i += kScanThreads) s_row[i] = r4[i];
        __syncthreads();
        double best = 0.0;
        int bestj = -1;
        for (int k0 = warp * kScanCodes; k0 < p.K; k0 += kScanWarps * kScanCodes) {
            double acc[kScanCodes];
            const float4* c4[kScanCodes];
#pragma unroll
            for (int c = 0; c < kScanCodes; ++c) {
                acc[c] = 0.0;
                c4[c] = reinterpret_cast<const float4*>(p.cb + static_cast<long long>(min(k0 + c, p.K - 1)) * p.dp);
            }
            for (int i = lane; i < dp4; i += 32) {
                const float4 a = s_row[i];
#pragma unroll
                for (int c = 0; c < kScanCodes; ++c) {
                    const float4 b = __ldg(c4[c] + i);
                    acc[c] = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc[c]);
                    acc[c] = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc[c]);
                    acc[c] = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc[c]);
                    acc[c] = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < kScanCodes; ++c) {
                const int k = k0 + c;
                const double dot = warp_sum(acc[c]);
                if (k < p.K) {
                    const double sc = p.cn64[k] - 2.0 * dot;
                    if (bestj < 0 || sc < best) { best = sc; bestj = k; }   // k ascending: first minimum kept
                }
            }
        }
        if (lane == 0) { s_best[warp] = best; s_idx[warp] = bestj; }
        __syncthreads();
        if (warp == 0) {
            double b = 0.0;
            int bj = -1;
            for (int w = 0; w < kScanWarps; ++w) {
                const int wj = s_idx[w];
                if (wj < 0) continue;
                if (bj < 0 || s_best[w] < b || (s_best[w] == b && wj < bj)) { b = s_best[w]; bj = wj; }
            }
            apply_code(p, row, bj);
            if (count_stats && p.stats != nullptr && lane == 0 && scan_list == nullptr) atomicAdd(p.stats + 2, 1ULL);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------- sampling
// The reference's default selection (nat.py:2150-2154): probs = softmax(-cdist / temperature), then
// torch.multinomial(probs, 1), which ATen evaluates as  argmax_k probs_k / q_k  with q ~ Exp(1) drawn for every
// (frame, code) pair (aten/src/ATen/native/Distributions.cpp, multinomial, the n_sample == 1 path).
//
// One CTA per frame, every code scored exactly: d_k = sqrt(max(0, ||r||^2 - 2 r.c_k + ||c_k||^2)) from fp64
// accumulation rounded to fp32 (the reference's fp32 sgemm expansion carries its own rounding), then the
// same fp32 steps as ATen: z = -d / T, e = exp(z - max z), p = e / sum e, v = p / q, first maximum.
//   noise != nullptr: q is read from noise[row * K + k] -- the host drew it from torch's CPU generator in the
//     reference's order, so codes equal the reference's except where v's two best values nearly tie;
//   noise == nullptr: q = -log(u), u in (0, 1] from Philox4x32-10 keyed by `seed`, counter (global row, k / 4, draw),
//     a different random stream from torch's: equal in distribution only.
struct SampleArgs {
    const float* noise;            // [n, K] for this chunk and layer, or nullptr
    unsigned long long seed;       // Philox key
    unsigned long long row0;       // global index of the chunk's first frame (Philox counter)
    unsigned draw;                 // distinguishes layers / calls in the Philox counter
    float temperature;
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n2 = hi0 ^ ctr[3] ^ k1;
        ctr[0] = n0; ctr[1] = lo1; ctr[2] = n2; ctr[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// Exp(1) from 32 random bits: u = (bits + 1) / 2^32 in (0, 1], q = -log(u) >= 0; q == 0 only for u == 1, replaced by
// the smallest positive value the formula can produce so that p / q stays finite.
__device__ __forceinline__ float exp1_from_bits(uint32_t bits) {
    const float u = (static_cast<float>(bits >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float q = -__logf(u);           // MUFU.LG2 * ln 2: the draw is noise, two ulps of it do not matter
    return q > 0.f ? q : 5.9604645e-8f;
}

template <int BLOCK>
__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sh[0];
    for (int w = 1; w < BLOCK / 32; ++w) r = fmaxf(r, sh[w]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads)
sample_scan_kernel(UpdateArgs p, SampleArgs sa) {
    extern __shared__ float4 s_dyn[];                   // [dp / 4] frame, then [K] scores
    __shared__ float s_red[kScanWarps];
    __shared__ double s_redd[kScanWarps];
    __shared__ float s_bv[kScanWarps];
    __shared__ int s_bi[kScanWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dp4 = p.dp >> 2;
    float4* s_row = s_dyn;
    float* s_z = reinterpret_cast<float*>(s_dyn + dp4);
    for (int row = blockIdx.x; row < p.n; row += gridDim.x) {
        const float4* r4 = reinterpret_cast<const float4*>(p.r + static_cast<long long>(row) * p.dp);
        double xn = 0.0;
        for (int i = threadIdx.x; i < dp4; i += kScanThreads) {
            const float4 v = r4[i];
            s_row[i] = v;
            xn += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
                  static_cast<double>(v.w) * v.w;
        }
        xn = warp_sum(xn);
        if (lane == 0) s_redd[warp] = xn;
        __syncthreads();
        xn = 0.0;
        for (int w = 0; w < kScanWarps; ++w) xn += s_redd[w];
        // z_k = -d_k / T for every code
        for (int k0 = warp * kScanCodes; k0 < p.K; k0 += kScanWarps * kScanCodes) {
            double acc[kScanCodes];
            const float4* c4[kScanCodes];
#pragma unroll
            for (int c = 0; c < kScanCodes; ++c) {
                acc[c] = 0.0;
                c4[c] = reinterpret_cast<const float4*>(p.cb + static_cast<long long>(min(k0 + c, p.K - 1)) * p.dp);
            }
            for (int i = lane; i < dp4; i += 32) {
                const float4 a = s_row[i];
#pragma unroll
                for (int c = 0; c < kScanCodes; ++c) {
                    const float4 b = __ldg(c4[c] + i);
                    acc[c] = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc[c]);
                    acc[c] = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc[c]);
                    acc[c] = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc[c]);
                    acc[c] = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < kScanCodes; ++c) {
                const int k = k0 + c;
                const double dot = warp_sum(acc[c]);
                if (k < p.K && lane == 0) {
                    const float d2 = static_cast<float>(fmax(xn + p.cn64[k] - 2.0 * dot, 0.0));
                    s_z[k] = __fdiv_rn(-__fsqrt_rn(d2), sa.temperature);
                }
            }
        }
        __syncthreads();
        float zmax = -__int_as_float(0x7F800000);
        for (int k = threadIdx.x; k < p.K; k += kScanThreads) zmax = fmaxf(zmax, s_z[k]);
        zmax = block_reduce_max<kScanThreads>(zmax, s_red);
        float esum = 0.f;
        for (int k = threadIdx.x; k < p.K; k += kScanThreads) {
            const float e = expf(s_z[k] - zmax);
            s_z[k] = e;
            esum += e;
        }
        esum = warp_sum(esum);
        if (lane == 0) s_red[warp] = esum;
        __syncthreads();
        esum = 0.f;
        for (int w = 0; w < kScanWarps; ++w) esum += s_red[w];
        // v_k = (e_k / sum) / q_k, first maximum
        float bv = -1.f;
        int bi = 0x7FFFFFFF;
        for (int k = threadIdx.x; k < p.K; k += kScanThreads) {
            float q;
            if (sa.noise != nullptr) {
                q = __ldg(sa.noise + static_cast<long long>(row) * p.K + k);
            } else {
                const unsigned long long g = sa.row0 + static_cast<unsigned long long>(row);
                uint32_t ctr[4] = {static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(k >> 2), sa.draw};
                philox4x32_10(ctr, static_cast<uint32_t>(sa.seed), static_cast<uint32_t>(sa.seed >> 32));
                q = exp1_from_bits(ctr[k & 3]);
            }
            const float v = __fdiv_rn(__fdiv_rn(s_z[k], esum), q);
            if (v > bv) { bv = v; bi = k; }              // k ascending per thread: first maximum kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_bv[warp] = bv; s_bi[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            float b = s_bv[0];
            int bj = s_bi[0];
            for (int w = 1; w < kScanWarps; ++w)
                if (s_bv[w] > b || (s_bv[w] == b && s_bi[w] < bj)) { b = s_bv[w]; bj = s_bi[w]; }
            bj = min(max(bj, 0), p.K - 1);               // all-NaN rows (non-finite input) still yield a valid index
            apply_code(p, row, bj);
        }
        __syncthreads();
    }
}

// Philox sampling from tensor-core scores: one warp per frame over the raw accumulators the coarse GEMM dumped
// (acc[row, k] = x_tilde . c_tilde_k, fp16 operands, fp32 accumulate). Distances carry the coarse pass's rounding
// (~1e-5 relative), which is immaterial for a path whose contract is distributional; the noise is the SAME Philox
// stream as sample_scan_kernel's (counter: frame, k / 4, draw), so the two paths pick the same code except where the
// two best values of  -d_k / T - log q_k  nearly tie (tests compare them frame by frame).
__global__ void __launch_bounds__(256)
sample_from_acc_kernel(UpdateArgs p, SampleArgs sa, const float* __restrict__ acc, int acc_ld,
                       const float* __restrict__ cn32) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int dp4 = p.dp >> 2;
    const float inv_t = 1.f / sa.temperature;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.n; row += warps) {
        const float4* r4 = reinterpret_cast<const float4*>(p.r + static_cast<long long>(row) * p.dp);
        double xn = 0.0;
        for (int i = lane; i < dp4; i += 32) {
            const float4 v = r4[i];
            xn += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
                  static_cast<double>(v.w) * v.w;
        }
        const float xnf = static_cast<float>(warp_sum(xn));
        const float alpha = p.rowinfo[row].x;
        const float4* a4 = reinterpret_cast<const float4*>(acc + static_cast<long long>(row) * acc_ld);
        const unsigned long long g = sa.row0 + static_cast<unsigned long long>(row);
        float bv = -__int_as_float(0x7F800000);
        int bi = 0x7FFFFFFF;
        for (int k4 = lane; k4 * 4 < p.K; k4 += 32) {           // four consecutive codes per lane: one Philox block
            const float4 a = a4[k4];
            const float4 c = __ldg(reinterpret_cast<const float4*>(cn32) + k4);
            uint32_t ctr[4] = {static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(k4), sa.draw};
            philox4x32_10(ctr, static_cast<uint32_t>(sa.seed), static_cast<uint32_t>(sa.seed >> 32));
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k4 * 4 + e;
                if (k < p.K) {
                    float d;                                        // sqrt.approx: 1 ulp, one MUFU instead of a Newton step
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(d) : "f"(fmaxf(fmaf(av[e], alpha, cv[e]) + xnf, 0.f)));
                    const float v = -d * inv_t - __logf(exp1_from_bits(ctr[e]));
                    if (v > bv) { bv = v; bi = k; }                 // k ascending per lane: first maximum kept
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        bi = min(max(bi, 0), p.K - 1);
        apply_code(p, row, bi);
    }
}

// ---------------------------------------------------------------------------------------------------- small inputs
// Exact argmin from a dumped accumulator matrix: the low-latency form for inputs of a few tiles, where the persistent
// fused kernel would run a whole stack on a handful of SMs. The coarse GEMM is dealt over tiles x codebook chunks
// (rvq_gemm_topk_kernel<DUMP>, n_splits) and this kernel, one warp per frame, applies the same certificate as the
// fused kernel to the same coarse scores  s_k = fma(acc_k, alpha, ||c_k||^2): every code within the frame's proven
// window of the minimum is a candidate; one candidate is certified, several are re-ranked exactly in fp64
// (ties -> lowest index), then the code is applied (residual update, next operand).
__global__ void __launch_bounds__(256)
argmin_from_acc_kernel(UpdateArgs p, const float* __restrict__ acc, int acc_ld, const float* __restrict__ cn32) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int dp4 = p.dp >> 2;
    const int k4n = (p.K + 3) >> 2;
    unsigned long long n_cert = 0, n_rerank = 0;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.n; row += warps) {
        const float4 ri = p.rowinfo[row];
        const float alpha = ri.x, window = ri.z;
        const float4* a4 = reinterpret_cast<const float4*>(acc + static_cast<long long>(row) * acc_ld);
        const float4* c4n = reinterpret_cast<const float4*>(cn32);
        const float inf = __int_as_float(0x7F800000);
        float m = inf;
        for (int k4 = lane; k4 < k4n; k4 += 32) {              // padding columns carry ||c||^2 = +inf
            const float4 a = a4[k4];
            const float4 c = __ldg(c4n + k4);
            m = fminf(fminf(m, fminf(fmaf(a.x, alpha, c.x), fmaf(a.y, alpha, c.y))),
                      fminf(fmaf(a.z, alpha, c.z), fmaf(a.w, alpha, c.w)));
        }
        m = -warp_max(-m);
        const float thr = fmaf(fabsf(m) + window, 2.4e-7f, m + window);     // m + W rounded up (superset of the window)
        const float4* r4 = reinterpret_cast<const float4*>(p.r + static_cast<long long>(row) * p.dp);
        double best = 0.0;
        int bestj = -1, n_cand = 0, only = 0;
        for (int k40 = 0; k40 < k4n; k40 += 32) {
            const int k4 = k40 + lane;
            float sc[4] = {inf, inf, inf, inf};
            if (k4 < k4n) {
                const float4 a = a4[k4];
                const float4 c = __ldg(c4n + k4);
                sc[0] = fmaf(a.x, alpha, c.x); sc[1] = fmaf(a.y, alpha, c.y);
                sc[2] = fmaf(a.z, alpha, c.z); sc[3] = fmaf(a.w, alpha, c.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                unsigned hits = __ballot_sync(0xffffffffu, sc[e] <= thr && k4 * 4 + e < p.K);
                while (hits != 0) {                               // warp-uniform
                    const int src = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const int k = (k40 + src) * 4 + e;
                    if (n_cand == 0) {
                        only = k;                                 // scored only if a second candidate turns up
                    } else {
                        if (n_cand == 1) {
                            best = exact_score(r4, reinterpret_cast<const float4*>(p.cb + static_cast<long long>(only) * p.dp),
                                               dp4, p.cn64[only]);
                            bestj = only;
                        }
                        const double s = exact_score(r4, reinterpret_cast<const float4*>(p.cb + static_cast<long long>(k) * p.dp),
                                                     dp4, p.cn64[k]);
                        if (s < best || (s == best && k < bestj)) { best = s; bestj = k; }
                    }
                    ++n_cand;
                }
            }
        }
        int j = n_cand <= 1 ? only : bestj;
        if (n_cand == 0) j = 0;                                   // cannot happen for finite input (the minimum is a hit)
        if (n_cand <= 1) ++n_cert; else ++n_rerank;
        apply_code(p, row, j);
    }
    if (p.stats != nullptr && lane == 0) {
        if (n_cert) atomicAdd(p.stats + 0, n_cert);
        if (n_rerank) atomicAdd(p.stats + 1, n_rerank);
    }
}

// ---------------------------------------------------------------------------------------------------- loss
// Deterministic fixed-order sum of row_loss[0..n) added to *acc (one block).
// Block b sums row_loss[b * ld ..] into acc[b]: all layers of a stack in one launch.
__global__ void __launch_bounds__(1024)
reduce_loss_kernel(const double* __restrict__ row_loss, int n, double* __restrict__ acc, long long ld = 0) {
    __shared__ double sh[1024];
    row_loss += blockIdx.x * ld;
    acc += blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += row_loss[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *acc += sh[0];
}
__global__ void finish_loss_kernel(const double* __restrict__ acc, const double* __restrict__ acc2, int L, double count,
                                   float w, float* __restrict__ loss_out) {
    const int l = threadIdx.x;
    if (l < L) {
        const double total = acc2 != nullptr ? acc[l] + acc2[l] : acc[l];     // the two lanes, fixed order
        const float mse = static_cast<float>(total / count);
        loss_out[l] = __fadd_rn(mse, __fmul_rn(w, mse));           // q_latent + w * e_latent, nat.py:2164
    }
}

// ---------------------------------------------------------------------------------------------------- outputs
// rows [n, Dp] holding x -> rows holding sum_l q_ste_l, replaying the chain from the emitted codes; optionally the
// per-frame loss sums of every layer (sum_d t^2 with t = q - r, nat.py:2162-2163) from the same replay. One warp/frame.
// `src` may equal `dst` (in place); `dst` may be null (losses only).
__global__ void __launch_bounds__(256)
reconstruct_rows_kernel(const float* __restrict__ src, float* dst, int n, int dp, const float* __restrict__ cb_all,
                        long long cb_layer_ld, int L, const void* __restrict__ codes, int code_dtype,
                        long long codes_ld, long long code_off, double* __restrict__ row_loss = nullptr,
                        long long loss_ld = 0) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int dp4 = dp >> 2;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
        const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(row) * dp);
        float4* d4 = dst != nullptr ? reinterpret_cast<float4*>(dst + static_cast<long long>(row) * dp) : nullptr;
        int js[16];
        double loss[16];
        for (int l = 0; l < L; ++l) { js[l] = load_code(codes, code_dtype, l * codes_ld + code_off + row); loss[l] = 0.0; }
        for (int i = lane; i < dp4; i += 32) {
            float4 rv = s4[i];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = 0; l < L; ++l) {
                const float4 cv = __ldg(reinterpret_cast<const float4*>(cb_all + l * cb_layer_ld +
                                                                         static_cast<long long>(js[l]) * dp) + i);
                float t, q;
                double ls = 0.0;
                t = __fsub_rn(cv.x, rv.x); q = __fadd_rn(rv.x, t); rv.x = __fsub_rn(rv.x, q); acc.x = l ? __fadd_rn(acc.x, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                t = __fsub_rn(cv.y, rv.y); q = __fadd_rn(rv.y, t); rv.y = __fsub_rn(rv.y, q); acc.y = l ? __fadd_rn(acc.y, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                t = __fsub_rn(cv.z, rv.z); q = __fadd_rn(rv.z, t); rv.z = __fsub_rn(rv.z, q); acc.z = l ? __fadd_rn(acc.z, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                t = __fsub_rn(cv.w, rv.w); q = __fadd_rn(rv.w, t); rv.w = __fsub_rn(rv.w, q); acc.w = l ? __fadd_rn(acc.w, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                if (row_loss != nullptr) loss[l] += ls;
            }
            if (d4 != nullptr) d4[i] = acc;
        }
        if (row_loss != nullptr) {
            for (int l = 0; l < L; ++l) {
                const double v = warp_sum(loss[l]);
                if (lane == 0) row_loss[l * loss_ld + row] = v;
            }
        }
    }
}
// The same replay for the [B, D, T] form of the quantised sum (what ResidualVectorQuantizer.forward returns,
// nat.py:1408-1415): a CTA replays 32 consecutive frames into a shared-memory tile and writes the tile transposed, so
// the [frames, Dp] intermediate of reconstruct_rows_kernel + rows_to_bct_kernel (one write and one read of every
// element) never exists. Tile: [32][dp] floats in 16-byte chunks, chunk c of frame f stored at chunk c ^ (f & 7):
// the row-wise float4 stores (lane = chunk) and the column-wise float4 loads (lane = frame) are both conflict free.
// Same arithmetic, in the same order, as reconstruct_rows_kernel. Dynamic smem: kReplayFrames * dp floats.
constexpr int kReplayFrames = 32;
constexpr int kReplayThreads = 512;         // two CTAs of 16 warps per SM next to a 96 KB tile each
// LT > 0: the layer count as a compile-time constant (all code vectors of a chunk in flight together); 0: run-time L.
template <int LT>
__global__ void __launch_bounds__(kReplayThreads)
reconstruct_bct_kernel(const float* __restrict__ src, int n, int dp, int D, const float* __restrict__ cb_all,
                       long long cb_layer_ld, int L_rt, const void* __restrict__ codes, int code_dtype, long long codes_ld,
                       long long code_off, double* __restrict__ row_loss, long long loss_ld, long long T, long long n0,
                       float* __restrict__ out) {
    extern __shared__ __align__(16) float replay_tile[];
    float4* tile4 = reinterpret_cast<float4*>(replay_tile);
    constexpr int LMAX = LT > 0 ? LT : 16;
    const int L = LT > 0 ? LT : L_rt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dp4 = dp >> 2;
    const int n_tiles = (n + kReplayFrames - 1) / kReplayFrames;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int f0 = tile * kReplayFrames;
        for (int fr = warp; fr < kReplayFrames; fr += kReplayThreads / 32) {
            const int row = f0 + fr;
            if (row >= n) break;
            const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(row) * dp);
            const float4* c4[LMAX];
            double loss[LMAX];
#pragma unroll
            for (int l = 0; l < LMAX; ++l) {
                if (l < L) {
                    const int j = load_code(codes, code_dtype, l * codes_ld + code_off + row);
                    c4[l] = reinterpret_cast<const float4*>(cb_all + l * cb_layer_ld + static_cast<long long>(j) * dp);
                    loss[l] = 0.0;
                }
            }
            for (int i = lane; i < dp4; i += 32) {
                float4 rv = s4[i];
                float4 cvs[LMAX];
#pragma unroll
                for (int l = 0; l < LMAX; ++l)
                    if (LT > 0) cvs[l] = __ldg(c4[l] + i);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int l = 0; l < LMAX; ++l) {
                    if (l < L) {
                        const float4 cv = LT > 0 ? cvs[l] : __ldg(c4[l] + i);
                        float t, q;
                        double ls = 0.0;
                        t = __fsub_rn(cv.x, rv.x); q = __fadd_rn(rv.x, t); rv.x = __fsub_rn(rv.x, q); acc.x = l ? __fadd_rn(acc.x, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                        t = __fsub_rn(cv.y, rv.y); q = __fadd_rn(rv.y, t); rv.y = __fsub_rn(rv.y, q); acc.y = l ? __fadd_rn(acc.y, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                        t = __fsub_rn(cv.z, rv.z); q = __fadd_rn(rv.z, t); rv.z = __fsub_rn(rv.z, q); acc.z = l ? __fadd_rn(acc.z, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                        t = __fsub_rn(cv.w, rv.w); q = __fadd_rn(rv.w, t); rv.w = __fsub_rn(rv.w, q); acc.w = l ? __fadd_rn(acc.w, q) : q; ls += static_cast<double>(__fmul_rn(t, t));
                        if (row_loss != nullptr) loss[l] += ls;
                    }
                }
                tile4[fr * dp4 + (i ^ (fr & 7))] = acc;
            }
            if (row_loss != nullptr) {
#pragma unroll
                for (int l = 0; l < LMAX; ++l) {
                    if (l < L) {
                        const double v = warp_sum(loss[l]);
                        if (lane == 0) row_loss[l * loss_ld + row] = v;
                    }
                }
            }
        }
        __syncthreads();
        // lane = frame: 32 consecutive t of one feature are one 128-byte run of the output
        const int row = f0 + lane;
        const long long g = n0 + row, b = g / T, t = g - b * T;
        float* o = out + (b * D) * T + t;
        for (int c = warp; c < dp4; c += kReplayThreads / 32) {
            if (row < n) {
                const float4 v = tile4[lane * dp4 + (c ^ (lane & 7))];
                const int d = c * 4;
                if (d + 0 < D) o[static_cast<long long>(d + 0) * T] = v.x;
                if (d + 1 < D) o[static_cast<long long>(d + 1) * T] = v.y;
                if (d + 2 < D) o[static_cast<long long>(d + 2) * T] = v.z;
                if (d + 3 < D) o[static_cast<long long>(d + 3) * T] = v.w;
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
copy_rows_out_kernel(const float* __restrict__ r, int n, int dp, int D, float* __restrict__ out) {
    const long long total = static_cast<long long>(n) * D;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / D;
        out[i] = r[row * dp + (i - row * D)];
    }
}

// decode into [B, D, T]: a CTA sums the code vectors of 32 consecutive frames row-wise (coalesced 16-byte gathers) into
// the swizzled tile of reconstruct_bct_kernel and writes it transposed in 128-byte runs. The element-per-thread kernel
// below reads a 32-byte sector per 4-byte element in this layout (3.1 ms for 270 000 x 768 x 4 layers; this: see
// profiles/). Same sum order: ((0 + c_0) + c_1) + ...
template <int LT>
__global__ void __launch_bounds__(kReplayThreads)
decode_bct_kernel(const float* __restrict__ cb_all, long long cb_layer_ld, int dp, int D, int L_rt,
                  const void* __restrict__ codes, int code_dtype, long long N, long long T, float* __restrict__ out) {
    extern __shared__ __align__(16) float replay_tile[];
    float4* tile4 = reinterpret_cast<float4*>(replay_tile);
    constexpr int LMAX = LT > 0 ? LT : 16;
    const int L = LT > 0 ? LT : L_rt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dp4 = dp >> 2;
    const long long n_tiles = (N + kReplayFrames - 1) / kReplayFrames;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long f0 = tile * kReplayFrames;
        for (int fr = warp; fr < kReplayFrames; fr += kReplayThreads / 32) {
            const long long row = f0 + fr;
            if (row >= N) break;
            const float4* c4[LMAX];
#pragma unroll
            for (int l = 0; l < LMAX; ++l)
                if (l < L) c4[l] = reinterpret_cast<const float4*>(cb_all + l * cb_layer_ld +
                                                                    static_cast<long long>(load_code(codes, code_dtype, l * N + row)) * dp);
            for (int i = lane; i < dp4; i += 32) {
                float4 cvs[LMAX];
#pragma unroll
                for (int l = 0; l < LMAX; ++l)
                    if (LT > 0) cvs[l] = __ldg(c4[l] + i);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int l = 0; l < LMAX; ++l) {
                    if (l < L) {
                        const float4 cv = LT > 0 ? cvs[l] : __ldg(c4[l] + i);
                        acc.x = __fadd_rn(acc.x, cv.x); acc.y = __fadd_rn(acc.y, cv.y);
                        acc.z = __fadd_rn(acc.z, cv.z); acc.w = __fadd_rn(acc.w, cv.w);
                    }
                }
                tile4[fr * dp4 + (i ^ (fr & 7))] = acc;
            }
        }
        __syncthreads();
        const long long row = f0 + lane, b = row / T, t = row - b * T;
        float* o = out + (b * D) * T + t;
        for (int c = warp; c < dp4; c += kReplayThreads / 32) {
            if (row < N) {
                const float4 v = tile4[lane * dp4 + (c ^ (lane & 7))];
                const int d = c * 4;
                if (d + 0 < D) o[static_cast<long long>(d + 0) * T] = v.x;
                if (d + 1 < D) o[static_cast<long long>(d + 1) * T] = v.y;
                if (d + 2 < D) o[static_cast<long long>(d + 2) * T] = v.z;
                if (d + 3 < D) o[static_cast<long long>(d + 3) * T] = v.w;
            }
        }
        __syncthreads();
    }
}

// decode: out = ((0 + cb_0[code_0]) + cb_1[code_1]) + ...   (nat.py:1438-1444). One thread per output element.
__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ cb_all, long long cb_layer_ld, int dp, int D, int L_used,
              const void* __restrict__ codes, int code_dtype, long long N, long long T, int layout,
              float* __restrict__ out) {
    const long long total = N * D;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long n;
        int d;
        if (layout == 0) {          // [B, D, T]: i = (b*D + d)*T + t
            const long long t = i % T, bd = i / T;
            d = static_cast<int>(bd % D);
            n = (bd / D) * T + t;
        } else {
            n = i / D;
            d = static_cast<int>(i - n * D);
        }
        float acc = 0.f;
        for (int l = 0; l < L_used; ++l) {
            const int j = load_code(codes, code_dtype, l * N + n);
            acc = __fadd_rn(acc, __ldg(cb_all + l * cb_layer_ld + static_cast<long long>(j) * dp + d));
        }
        out[i] = acc;
    }
}

}  // namespace rows
}  // namespace nat
