// C ABI of the B200-native RVQ + front-end path (declared in include/nat_b200.h). Host-side orchestration only:
// argument checks, workspace carving, TMA descriptors, launches on the caller's stream. No CPU compute path exists.
#include "../../include/nat_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "mel_fft.cuh"
#include "nat_common.cuh"
#include "rvq_gemm_sm100.cuh"
#include "rvq_prepare.cuh"
#include "rvq_rows.cuh"
#include "rvq_stack_sm100.cuh"
#include "token_stats.cuh"
#include "interp.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define NAT_CUDA(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(NAT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

std::atomic<unsigned long long> g_launches{0};

// Optional per-kernel-class timing (nat_rvq_encode_profile_f32): CUDA events around every launch, on its stream.
struct Profiler {
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
};
thread_local Profiler* g_prof = nullptr;

struct LaunchScope {
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    LaunchScope(int cls, cudaStream_t s) : st(s) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (g_prof != nullptr) {
            cudaEvent_t a;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
            g_prof->spans.push_back({cls, a, b});
        }
    }
    ~LaunchScope() { if (b != nullptr) cudaEventRecord(b, st); }
};
#define NAT_LAUNCH(cls, st, ...) do { LaunchScope scope__((cls), (st)); __VA_ARGS__; } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// fp16 row-major [rows, cols] -> 2-D map with a (64 x box_rows) box and 128-byte swizzle. The descriptor is a pure
// function of its arguments: the last few are remembered per thread, so a caller that reuses its workspace (the 1 s
// chunk path) does not re-encode the same operand map on every call.
int make_map_f16(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
    struct Entry { const void* base; long long rows, cols; int box_rows; CUtensorMap map; };
    thread_local Entry cache[8];
    thread_local int next = 0, used = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].base == base && cache[i].rows == rows && cache[i].cols == cols && cache[i].box_rows == box_rows) {
            *map = cache[i].map;
            return NAT_OK;
        }
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(NAT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NAT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    cache[next] = Entry{base, rows, cols, box_rows, *map};
    next = (next + 1) % 8;
    used = std::min(used + 1, 8);
    return NAT_OK;
}

int code_bytes(int dtype) { return dtype == NAT_CODES_I64 ? 8 : dtype == NAT_CODES_I32 ? 4 : 2; }

long long chunk_cap_rows() {
    static long long cap = [] {
        const char* e = getenv("NAT_RVQ_CHUNK_ROWS");
        long long v = e ? atoll(e) : 0;
        if (v < 128) v = 1 << 19;
        return round_up(v, 128);
    }();
    return cap;
}

struct Workspace {
    float* r;
    float* r_work;        // a second set of residual rows: lets the fused kernel leave `r` (the prepared x rows) intact
    __half* a;
    float4* rowinfo;
    float* rowamax;
    nat::gemm::Cand* cand;
    double* row_loss;     // [L][rows]
    int* scan_list;
    int* scan_count;      // [L]
    double* loss_acc;     // [L]
    long long rows;       // capacity
};

constexpr size_t kWsFixed = 4096;
constexpr int kMaxStacks = 2;             // S0-S3 and A0-A3 of one tokenizer per call

size_t ws_per_row(int dp, int L) { return static_cast<size_t>(dp) * 10 + 16 + 4 + sizeof(nat::gemm::Cand) + 8 * L + 4; }

bool carve(void* base, size_t bytes, int dp, int L, long long want_rows, Workspace* ws) {
    if (bytes <= kWsFixed + 256 * 10) return false;
    long long rows = static_cast<long long>((bytes - kWsFixed - 256 * 10) / ws_per_row(dp, L));
    rows = std::min(rows, round_up(want_rows, 128));
    rows = rows / 128 * 128;
    if (rows < 128) return false;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t n) { char* q = p; p += round_up((long long)n, 256); return q; };
    ws->scan_count = reinterpret_cast<int*>(take(kWsFixed / 2));
    ws->loss_acc = reinterpret_cast<double*>(take(kWsFixed / 2));
    ws->r = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * dp * 4));
    ws->r_work = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * dp * 4));
    ws->a = reinterpret_cast<__half*>(take(static_cast<size_t>(rows) * dp * 2));
    ws->rowinfo = reinterpret_cast<float4*>(take(static_cast<size_t>(rows) * 16));
    ws->rowamax = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * 4));
    ws->cand = reinterpret_cast<nat::gemm::Cand*>(take(static_cast<size_t>(rows) * sizeof(nat::gemm::Cand)));
    ws->row_loss = reinterpret_cast<double*>(take(static_cast<size_t>(rows) * 8 * L));
    ws->scan_list = reinterpret_cast<int*>(take(static_cast<size_t>(rows) * 4));
    ws->rows = rows;
    return static_cast<size_t>(p - static_cast<char*>(base)) <= bytes;
}

}  // namespace

template <int NV>
static cudaError_t stack_kernel_attrs() {
    const void* fns[4] = {(const void*)nat::stack::rvq_stack_kernel<NV, 1, false>, (const void*)nat::stack::rvq_stack_kernel<NV, 2, false>,
                          (const void*)nat::stack::rvq_stack_kernel<NV, 1, true>, (const void*)nat::stack::rvq_stack_kernel<NV, 2, true>};
    for (const void* f : fns) {
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, nat::stack::SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// One persistent launch of the fused stack kernel: plain grid (pair == 1) or clusters of two CTAs (pair == 2).
struct StackMaps { CUtensorMap a0, aw, b; };

template <int NV>
static cudaError_t launch_stack(int pair, int grid, cudaStream_t st, const StackMaps& m,
                                const nat::stack::StackArgs& sa, bool programmatic) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(nat::stack::NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = nat::stack::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    // a stack that shares nothing written by the launch before it (the other stack of the same tokenizer) may start
    // on every SM that launch has already left: programmatic dependent launch, no wait inside the kernel
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = programmatic ? 2 : 1;
    // the instrumented instantiation runs only while counters or timing experiments are switched on
    const bool dbg = sa.dbg != nullptr || sa.dbg_mode != 0;
    if (pair == 2)
        return dbg ? cudaLaunchKernelEx(&cfg, nat::stack::rvq_stack_kernel<NV, 2, true>, m.a0, m.aw, m.b, sa)
                   : cudaLaunchKernelEx(&cfg, nat::stack::rvq_stack_kernel<NV, 2, false>, m.a0, m.aw, m.b, sa);
    return dbg ? cudaLaunchKernelEx(&cfg, nat::stack::rvq_stack_kernel<NV, 1, true>, m.a0, m.aw, m.b, sa)
               : cudaLaunchKernelEx(&cfg, nat::stack::rvq_stack_kernel<NV, 1, false>, m.a0, m.aw, m.b, sa);
}

// Caller-owned state of the host-buffer entry points: a two-slot device staging arena (grown on demand), the copy
// stream and the events that order H2D / compute / D2H. One context serves one call at a time; concurrent calls on
// different streams use different contexts (nothing of this lives on the shared codebook handles).
struct nat_host_ctx {
    int device = -1;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct nat_rvq_codebooks {
    int L, K, D, dp, kp, device, sm_count;
    float* cbf;                       // [L, K, dp]
    __half* cbh;                      // [L, kp, dp]
    float* cn32;                      // [L, kp]
    float* cn32m;                     // [L, kp] the same with later exact duplicates masked (+inf): the argmin paths
    unsigned long long* row_hash;     // [K] scratch of the duplicate search
    double* cn64;                     // [L, K]
    nat::rows::LayerConst* lc;        // [L]
    int* scratch;                     // [L, kScratchPerLayer]
    CUtensorMap map_b;                // box 64 x 256 codes
    CUtensorMap map_b_half;           // box 64 x 128 codes: the half of a B tile one CTA of a pair stages
    bool pair_ok;                     // the device can co-schedule the two-CTA clusters of the fused kernel
    unsigned long long* stack_dbg;    // [sm_count][DBG_SLOTS] cycle counters of the last fused launch (debug hook)
    bool stack_dbg_on;
    // nat_rvq_encode_host_f32 (one stack, no caller context): a private host context, created on first use and
    // serialised by a mutex -- concurrent host-buffer calls should bring their own nat_host_ctx
    struct nat_host_ctx* host_ctx;
    std::mutex* host_mutex;
    // internal streams of the two-lane encode (created on first use)
    cudaStream_t side[2];
    cudaEvent_t side_ev[3];
};

extern "C" {

const char* nat_last_error(void) { return g_last_error.c_str(); }
int nat_abi_version(void) { return NAT_B200_ABI_VERSION; }

int nat_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, size_t name_len) {
    int dev = 0;
    NAT_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    NAT_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len) {
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return NAT_OK;
}

// ------------------------------------------------------------------------------------------------- codebooks
static int upload_codebooks(nat_rvq_codebooks* cb, const float* const* codebooks_dev, cudaStream_t st) {
    using namespace nat;
    NAT_CUDA(cudaMemsetAsync(cb->scratch, 0, sizeof(int) * cb->L * prepare::kScratchPerLayer, st));
    for (int l = 0; l < cb->L; ++l) {
        if (codebooks_dev[l] == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "codebook %d is null", l);
        float* dst = cb->cbf + static_cast<long long>(l) * cb->K * cb->dp;
        int* scr = cb->scratch + l * prepare::kScratchPerLayer;
        const long long total = static_cast<long long>(cb->K) * cb->dp;
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
        NAT_LAUNCH(6, st, prepare::pack_absmax_kernel<<<grid, 256, 0, st>>>(codebooks_dev[l], cb->K, cb->D, cb->dp, dst, scr));
        const int grid2 = std::min((cb->kp + 7) / 8, 148 * 8);
        NAT_LAUNCH(6, st, prepare::convert_norms_kernel<<<grid2, 256, 0, st>>>(
            dst, cb->K, cb->kp, cb->dp, cb->cbh + static_cast<long long>(l) * cb->kp * cb->dp,
            cb->cn32 + static_cast<long long>(l) * cb->kp, cb->cn64 + static_cast<long long>(l) * cb->K, scr));
    }
    for (int l = 0; l < cb->L; ++l) {
        const float* src = cb->cbf + static_cast<long long>(l) * cb->K * cb->dp;
        const int grid = std::min((cb->K + 7) / 8, 148 * 8);
        NAT_LAUNCH(6, st, prepare::row_hash_kernel<<<grid, 256, 0, st>>>(src, cb->K, cb->dp, cb->row_hash));
        NAT_LAUNCH(6, st, prepare::mask_duplicates_kernel<<<std::min((cb->kp + 7) / 8, 148 * 8), 256, 0, st>>>(
            src, cb->row_hash, cb->cn32 + static_cast<long long>(l) * cb->kp, cb->cn32m + static_cast<long long>(l) * cb->kp,
            cb->K, cb->kp, cb->dp));
    }
    NAT_LAUNCH(6, st, prepare::finish_consts_kernel<<<1, 32, 0, st>>>(cb->scratch, cb->L, cb->lc));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

int nat_rvq_codebooks_create(const float* const* codebooks_dev, int L, int K, int D, void* stream,
                             nat_rvq_codebooks** out) {
    if (out == nullptr || codebooks_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (L < 1 || L > 16) return fail(NAT_ERR_UNSUPPORTED, "num_quantizers per stack must be in [1, 16], got %d", L);
    if (K < 1 || K > 65536) return fail(NAT_ERR_UNSUPPORTED, "codebook_size must be in [1, 65536], got %d", K);
    if (D < 1 || D > 2048) return fail(NAT_ERR_UNSUPPORTED, "input_dim must be in [1, 2048], got %d", D);
    nat_rvq_codebooks* cb = new nat_rvq_codebooks();
    memset(cb, 0, sizeof *cb);
    cb->L = L; cb->K = K; cb->D = D;
    cb->dp = static_cast<int>(round_up(D, 64));
    cb->kp = static_cast<int>(round_up(K, 256));
    cb->host_mutex = new std::mutex();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = NAT_OK;
    auto guard = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == NAT_OK) rc = fail(NAT_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    guard(cudaGetDevice(&cb->device), "cudaGetDevice");
    if (rc == NAT_OK) guard(cudaDeviceGetAttribute(&cb->sm_count, cudaDevAttrMultiProcessorCount, cb->device), "attr");
    const size_t n_f = static_cast<size_t>(L) * K * cb->dp, n_h = static_cast<size_t>(L) * cb->kp * cb->dp;
    if (rc == NAT_OK) guard(cudaMalloc(&cb->cbf, n_f * 4), "cudaMalloc(cbf)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->cbh, n_h * 2), "cudaMalloc(cbh)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->cn32, static_cast<size_t>(L) * cb->kp * 4), "cudaMalloc(cn32)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->cn32m, static_cast<size_t>(L) * cb->kp * 4), "cudaMalloc(cn32m)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->row_hash, static_cast<size_t>(K) * 8), "cudaMalloc(row_hash)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->cn64, static_cast<size_t>(L) * K * 8), "cudaMalloc(cn64)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->lc, sizeof(nat::rows::LayerConst) * L), "cudaMalloc(lc)");
    if (rc == NAT_OK) guard(cudaMalloc(&cb->scratch, sizeof(int) * L * nat::prepare::kScratchPerLayer), "cudaMalloc");
    if (rc == NAT_OK) guard(cudaMemsetAsync(cb->cbh, 0, n_h * 2, st), "cudaMemsetAsync(cbh)");
    if (rc == NAT_OK) rc = upload_codebooks(cb, codebooks_dev, st);
    if (rc == NAT_OK) rc = make_map_f16(&cb->map_b, cb->cbh, static_cast<long long>(L) * cb->kp, cb->dp, 256);
    if (rc == NAT_OK) rc = make_map_f16(&cb->map_b_half, cb->cbh, static_cast<long long>(L) * cb->kp, cb->dp, 128);
    if (rc == NAT_OK) {
        guard(cudaFuncSetAttribute(nat::gemm::rvq_gemm_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   nat::gemm::SMEM_BYTES), "cudaFuncSetAttribute");
        guard(cudaFuncSetAttribute(nat::gemm::rvq_gemm_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   nat::gemm::SMEM_BYTES), "cudaFuncSetAttribute");
#ifdef NAT_ONLY_NV
        guard(stack_kernel_attrs<NAT_ONLY_NV>(), "cudaFuncSetAttribute");
#else
        guard(stack_kernel_attrs<2>(), "cudaFuncSetAttribute");
        guard(stack_kernel_attrs<4>(), "cudaFuncSetAttribute");
        guard(stack_kernel_attrs<6>(), "cudaFuncSetAttribute");
        guard(stack_kernel_attrs<8>(), "cudaFuncSetAttribute");
#endif
        {   // CTA pairs need two co-resident CTAs of 768 threads / 224 KB in one cluster (a TPC); ask the runtime
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(2, 1, 1);
            cfg.blockDim = dim3(nat::stack::NUM_THREADS, 1, 1);
            cfg.dynamicSmemBytes = nat::stack::SMEM_BYTES;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int clusters = 0;
            const cudaError_t qe = cudaOccupancyMaxActiveClusters(&clusters, nat::stack::rvq_stack_kernel<6, 2, false>, &cfg);
            (void)cudaGetLastError();
            cb->pair_ok = !(qe == cudaSuccess && clusters == 0);      // only a definite "no cluster fits" turns pairs off
            if (getenv("NAT_B200_VERBOSE"))
                fprintf(stderr, "nat_b200: cluster occupancy query: %s, %d active clusters -> pairs %s\n",
                        cudaGetErrorString(qe), clusters, cb->pair_ok ? "on" : "off");
        }
        guard(cudaFuncSetAttribute(nat::rows::prep_bct_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024), "cudaFuncSetAttribute");
        guard(cudaFuncSetAttribute(nat::rows::reconstruct_bct_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024), "cudaFuncSetAttribute");
        guard(cudaFuncSetAttribute(nat::rows::reconstruct_bct_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024), "cudaFuncSetAttribute");
        guard(cudaFuncSetAttribute(nat::rows::sample_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
              "cudaFuncSetAttribute");
        guard(cudaFuncSetAttribute(nat::rows::decode_bct_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024), "cudaFuncSetAttribute");
        guard(cudaFuncSetAttribute(nat::rows::decode_bct_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024), "cudaFuncSetAttribute");
    }
    if (rc != NAT_OK) {
        nat_rvq_codebooks_destroy(cb);
        return rc;
    }
    *out = cb;
    return NAT_OK;
}

int nat_rvq_codebooks_update(nat_rvq_codebooks* cb, const float* const* codebooks_dev, void* stream) {
    if (cb == nullptr || codebooks_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    return upload_codebooks(cb, codebooks_dev, static_cast<cudaStream_t>(stream));
}

int nat_rvq_codebooks_destroy(nat_rvq_codebooks* cb) {
    if (cb == nullptr) return NAT_OK;
    cudaFree(cb->cbf); cudaFree(cb->cbh); cudaFree(cb->cn32); cudaFree(cb->cn32m); cudaFree(cb->row_hash);
    cudaFree(cb->cn64); cudaFree(cb->lc);
    cudaFree(cb->scratch); cudaFree(cb->stack_dbg);
    if (cb->host_ctx) nat_host_ctx_destroy(cb->host_ctx);
    delete cb->host_mutex;
    for (auto& s2 : cb->side) if (s2) cudaStreamDestroy(s2);
    for (auto& e : cb->side_ev) if (e) cudaEventDestroy(e);
    delete cb;
    return NAT_OK;
}

int nat_rvq_codebooks_dims(const nat_rvq_codebooks* cb, int* L, int* K, int* D) {
    if (cb == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null handle");
    if (L) *L = cb->L;
    if (K) *K = cb->K;
    if (D) *D = cb->D;
    return NAT_OK;
}

// Inputs whose tiles x codebook chunks fit one wave of SMs take the split-GEMM + row-argmin path (low latency); its
// score matrix [rows, kp] fp32 sits at the tail of the caller's workspace.
static size_t small_input_scores_bytes(const nat_rvq_codebooks* cb, long long n_frames) {
    const long long tiles = (n_frames + 127) / 128;
    if (tiles * (cb->kp / nat::gemm::BLOCK_N) > cb->sm_count) return 0;
    return static_cast<size_t>(tiles) * 128 * cb->kp * sizeof(float);
}

size_t nat_rvq_workspace_bytes(const nat_rvq_codebooks* cb, int64_t n_frames) {
    if (cb == nullptr || n_frames <= 0) return kWsFixed + 256 * 10 + 128 * ws_per_row(64, 16);
    // two lanes, each holding half of the frames rounded up to a tile (see nat_rvq_encode_f32)
    const long long per_lane = std::min<long long>(round_up((n_frames + 1) / 2, 128), chunk_cap_rows() / 2);
    const size_t lane_bytes = static_cast<size_t>(round_up(kWsFixed + 256 * 10 + per_lane * ws_per_row(cb->dp, cb->L), 256));
    return 2 * lane_bytes + small_input_scores_bytes(cb, n_frames);
}

// ------------------------------------------------------------------------------------------------- encode
// Layer-0 preparation of frames [n0, n0 + n): fp32 rows, fp16 operand rows, {alpha, bias, window} and max |x| per
// frame. `rowinfo_b` / `lc_b`: a second stack quantising the same frames gets its own window from the same sums.
static int launch_layer0_prep(const nat_rvq_codebooks* cb, const Workspace& ws, const float* x, int layout,
                              long long T, long long n0, int n, cudaStream_t st, float4* rowinfo_b = nullptr,
                              const nat::rows::LayerConst* lc_b = nullptr, long long T_in = 0) {
    using namespace nat;
    if (T_in > 0) {          // time-base alignment fused into the loads: [B, D, T_in] -> T frames
        const size_t smem = static_cast<size_t>(rows::kPrepFrames) * (cb->dp + 1) * sizeof(float);
        if (layout != NAT_LAYOUT_BCT || smem > 200 * 1024)
            return fail(NAT_ERR_UNSUPPORTED, "time-base alignment is fused into the [B, D, T] preparation only");
        const float scale = static_cast<float>(T_in) / static_cast<float>(T);      // area_pixel_compute_scale<float>
        NAT_LAUNCH(0, st, rows::prep_bct_fused_kernel<<<(n + rows::kPrepFrames - 1) / rows::kPrepFrames, rows::kPrepThreads, smem, st>>>(
            x, T, cb->D, n0, n, cb->dp, ws.r, ws.a, ws.rowinfo, ws.rowamax, cb->lc, rowinfo_b, lc_b, T_in, scale));
        NAT_CUDA(cudaGetLastError());
        return NAT_OK;
    }
    const int warps_grid = std::min((n + 7) / 8, cb->sm_count * 16);
    if (layout == NAT_LAYOUT_ROWS) {
        NAT_LAUNCH(0, st, rows::prep_rows_kernel<<<warps_grid, 256, 0, st>>>(x + n0 * cb->D, cb->D, n, cb->D, cb->dp, ws.r, ws.a,
                                                           ws.rowinfo, ws.rowamax, cb->lc, false, rowinfo_b, lc_b));
    } else {
        const size_t smem = static_cast<size_t>(rows::kPrepFrames) * (cb->dp + 1) * sizeof(float);
        if (smem <= 200 * 1024) {
            NAT_LAUNCH(0, st, rows::prep_bct_fused_kernel<<<(n + rows::kPrepFrames - 1) / rows::kPrepFrames, rows::kPrepThreads, smem, st>>>(
                x, T, cb->D, n0, n, cb->dp, ws.r, ws.a, ws.rowinfo, ws.rowamax, cb->lc, rowinfo_b, lc_b));
        } else {
            dim3 grid((n + 31) / 32, cb->dp / 32);
            NAT_LAUNCH(0, st, rows::bct_to_rows_kernel<<<grid, 256, 0, st>>>(x, T, cb->D, n0, n, cb->dp, ws.r));
            NAT_LAUNCH(0, st, rows::prep_rows_kernel<<<warps_grid, 256, 0, st>>>(nullptr, 0, n, cb->D, cb->dp, ws.r, ws.a,
                                                                              ws.rowinfo, ws.rowamax, cb->lc, true, rowinfo_b, lc_b));
        }
    }
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

static bool fused_enabled() {
    const char* e = getenv("NAT_RVQ_FUSED");         // read every call: tests flip it to cross-check both paths
    return e == nullptr || atoi(e) != 0;
}
static bool small_path_enabled() {
    const char* e = getenv("NAT_RVQ_SMALL");           // read every call: tests flip it to cover the fused kernel on small inputs
    return e == nullptr || atoi(e) != 0;
}
static bool fused_pair() {
    const char* e = getenv("NAT_RVQ_PAIR");
    return e == nullptr || atoi(e) != 1;
}
static bool pdl_enabled() {
    const char* e = getenv("NAT_RVQ_PDL");              // read every call: the probes flip it for A/B measurements
    return (e == nullptr || atoi(e) != 0) && g_prof == nullptr;   // timed per launch: every kernel alone
}
static int fused_group(int pair) {
    const char* e = getenv("NAT_RVQ_GROUP");
    const int v = e ? atoi(e) : 0;
    return v >= 1 ? v : (pair == 2 ? 3 : 2);      // measured on B200: 270k x 768, K = 1024 (profiles/)
}

// One launch of the fused stack kernel: one stack over one range of frames.
struct StackLaunch {
    const nat_rvq_codebooks* cb = nullptr;
    const float* r0 = nullptr; const float4* rowinfo0 = nullptr; const float* rowamax0 = nullptr;   // prepared rows (layer 0)
    float* r = nullptr; __half* a = nullptr; float4* rowinfo = nullptr; float* rowamax = nullptr;      // rows the kernel writes
    void* codes = nullptr; long long codes_ld = 0, code_off = 0;
    double* row_loss = nullptr; long long loss_ld = 0;
    unsigned long long* stats = nullptr;
    CUtensorMap map_a0, map_a;   // operand rows: prepared / written by the kernel (may be the same buffer)
    int n_rows = 0, code_dtype = NAT_CODES_I16;
    bool programmatic = false;   // may overlap the tail of the launch before it (shares nothing that launch writes)
};

static int launch_fused_stack(const StackLaunch& sl, cudaStream_t st) {
    using namespace nat;
    const nat_rvq_codebooks* cb = sl.cb;
    const int n_tiles = (sl.n_rows + stack::BLOCK_M - 1) / stack::BLOCK_M;
    stack::StackArgs sa;
    memset(&sa, 0, sizeof sa);
    stack::StackRef& r = sa.s;
    r.cbf = cb->cbf; r.cn64 = cb->cn64; r.cn32 = cb->cn32m; r.lc = cb->lc;
    r.r0 = sl.r0; r.rowinfo0 = sl.rowinfo0; r.rowamax0 = sl.rowamax0;
    r.r = sl.r; r.a = sl.a; r.rowinfo = sl.rowinfo; r.rowamax = sl.rowamax;
    r.codes = sl.codes; r.codes_ld = sl.codes_ld; r.code_off = sl.code_off;
    r.row_loss = sl.row_loss; r.loss_ld = sl.loss_ld;
    r.stats = sl.stats;
    r.L = cb->L;
    // Residual write-back policy of the hot update form: the residual entering the last layer is never needed in
    // memory (only its fp16 operand is), and writing back after every other layer only (later layers replay the
    // skipped update from the emitted code, one extra L2 gather) measured faster than every layer or none.
    r.store_mask = 0x55555555 & ((1 << std::max(0, cb->L - 2)) - 1);
    { const char* e = getenv("NAT_RVQ_STORE_MASK"); if (e) r.store_mask &= atoi(e); }
    { const char* e = getenv("NAT_RVQ_STORE_MASK_SET"); if (e) r.store_mask = atoi(e); }      // A/B measurements
    sa.n_rows = sl.n_rows; sa.n_tiles = n_tiles; sa.K = cb->K; sa.kp = cb->kp; sa.dp = cb->dp;
    sa.code_dtype = sl.code_dtype;
    sa.dbg = cb->stack_dbg_on ? cb->stack_dbg : nullptr;
    { const char* e = getenv("NAT_RVQ_DBG_MODE"); sa.dbg_mode = e ? atoi(e) : 0; }
    // CTA pairs (one tcgen05.mma.cta_group::2 per two SMs) once there is more than one tile; NAT_RVQ_PAIR=1 keeps
    // the single-CTA form for A/B measurements.
    const int pair = (n_tiles >= 2 && cb->sm_count >= 2 && cb->pair_ok && fused_pair()) ? 2 : 1;
    const int grid = pair == 2 ? std::min((n_tiles + 1) & ~1, cb->sm_count & ~1) : std::min(n_tiles, cb->sm_count);
    StackMaps maps;
    maps.a0 = sl.map_a0; maps.aw = sl.map_a;
    maps.b = pair == 2 ? cb->map_b_half : cb->map_b;
    sa.group = fused_group(pair);
    const int nv = (cb->dp / 4 + 31) / 32;
    const bool pdl = sl.programmatic && pdl_enabled();
    cudaError_t le = cudaSuccess;
    NAT_LAUNCH(1, st, {
#ifdef NAT_ONLY_NV                                   /* A/B builds: one instantiation, a quarter of the compile time */
        le = launch_stack<NAT_ONLY_NV>(pair, grid, st, maps, sa, pdl);
#else
        if (nv <= 2) le = launch_stack<2>(pair, grid, st, maps, sa, pdl);
        else if (nv <= 4) le = launch_stack<4>(pair, grid, st, maps, sa, pdl);
        else if (nv <= 6) le = launch_stack<6>(pair, grid, st, maps, sa, pdl);
        else le = launch_stack<8>(pair, grid, st, maps, sa, pdl);
#endif
    });
    NAT_CUDA(le);
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

// Everything one chunk of frames [n0, n0 + n) needs, in order, on one stream.
struct EncodeCall {
    const nat_rvq_codebooks* cb;
    const float* x; int layout; long long T, N;
    void* codes; int code_dtype; float* quantized; bool want_loss; unsigned long long* stats; bool exact;
    // sampling mode (nat_rvq_sample_f32): per-layer temperatures (<= 0: that layer takes the exact argmin), optional
    // host-drawn noise [L, N, K], Philox key / draw base otherwise
    const float* temperatures = nullptr; const float* noise = nullptr;
    unsigned long long seed = 0, draw_base = 0;
    float* scores = nullptr;       // [chunk rows, kp] accumulator dump of the Philox fast path (stream-ordered allocation)
};

static int encode_chunk(const EncodeCall& c, const Workspace& ws, const CUtensorMap& map_a, long long n0, int n,
                        cudaStream_t st) {
    using namespace nat;
    const nat_rvq_codebooks* cb = c.cb;
    const int cbytes = code_bytes(c.code_dtype);
    const long long cb_layer_ld = static_cast<long long>(cb->K) * cb->dp;
    if (int rc = launch_layer0_prep(cb, ws, c.x, c.layout, c.T, n0, n, st)) return rc;
    const int n_tiles = (n + gemm::BLOCK_M - 1) / gemm::BLOCK_M;
    const bool fused = !c.exact && c.temperatures == nullptr && fused_enabled() && cb->dp <= 1024 && c.scores == nullptr;
    if (!fused) NAT_CUDA(cudaMemsetAsync(ws.scan_count, 0, sizeof(int) * cb->L, st));   // only the per-layer kernels list scans
    if (fused) {
        // one persistent launch for all L layers (rvq_stack_sm100.cuh); the update warps write in place
        StackLaunch sl;
        sl.cb = cb;
        // Codes only: the update warps work in place. With the quantised sum or the losses asked for (the form
        // ResidualVectorQuantizer.forward returns, nat.py:1410-1415) the kernel still runs its codes-only form, but
        // writes its residuals to the second set of rows: the prepared x rows stay intact and ONE replay of the chain
        // from the emitted codes (bit-identical op order) yields the quantised sum and every layer's loss sums.
        const bool extras = c.want_loss || c.quantized != nullptr;
        sl.r0 = ws.r; sl.rowinfo0 = ws.rowinfo; sl.rowamax0 = ws.rowamax;
        sl.r = extras ? ws.r_work : ws.r; sl.a = ws.a; sl.rowinfo = ws.rowinfo; sl.rowamax = ws.rowamax;
        sl.codes = c.codes; sl.codes_ld = c.N; sl.code_off = n0;
        sl.stats = c.stats;
        sl.map_a0 = map_a; sl.map_a = map_a;
        sl.n_rows = n; sl.code_dtype = c.code_dtype;
        if (int rc = launch_fused_stack(sl, st)) return rc;
        if (extras) {
            double* row_loss = c.want_loss ? ws.row_loss : nullptr;
            const size_t tile_smem = static_cast<size_t>(rows::kReplayFrames) * cb->dp * sizeof(float);
            if (c.quantized != nullptr && c.layout == NAT_LAYOUT_BCT && tile_smem <= 200 * 1024) {
                // replay + transposed write in one pass: the [frames, Dp] intermediate is never written
                const int n_tiles = (n + rows::kReplayFrames - 1) / rows::kReplayFrames;
                const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(2048 / rows::kReplayThreads, (220 * 1024) / (tile_smem + 1024))));
                auto kern = cb->L == 4 ? rows::reconstruct_bct_kernel<4> : rows::reconstruct_bct_kernel<0>;
                NAT_LAUNCH(5, st, kern<<<std::min(n_tiles, cb->sm_count * per_sm), rows::kReplayThreads, tile_smem, st>>>(
                    ws.r, n, cb->dp, cb->D, cb->cbf, cb_layer_ld, cb->L, c.codes, c.code_dtype, c.N, n0, row_loss, ws.rows,
                    c.T, n0, c.quantized));
            } else {
                const int warps_grid = std::min((n + 7) / 8, cb->sm_count * 16);
                // [frames, D] output without padding: the replay writes the caller's rows directly
                const bool direct = c.quantized != nullptr && c.layout == NAT_LAYOUT_ROWS && cb->dp == cb->D;
                float* dst = c.quantized == nullptr ? nullptr : direct ? c.quantized + n0 * cb->D : ws.r_work;
                NAT_LAUNCH(5, st, rows::reconstruct_rows_kernel<<<warps_grid, 256, 0, st>>>(
                    ws.r, dst, n, cb->dp, cb->cbf, cb_layer_ld, cb->L, c.codes, c.code_dtype, c.N, n0, row_loss, ws.rows));
                if (c.quantized != nullptr && !direct) {
                    if (c.layout == NAT_LAYOUT_ROWS) {
                        NAT_LAUNCH(5, st, rows::copy_rows_out_kernel<<<cb->sm_count * 8, 256, 0, st>>>(ws.r_work, n, cb->dp, cb->D,
                                                                                                    c.quantized + n0 * cb->D));
                    } else {
                        dim3 grid((n + 31) / 32, (cb->D + 31) / 32);
                        NAT_LAUNCH(5, st, rows::rows_to_bct_kernel<<<grid, 256, 0, st>>>(ws.r_work, cb->dp, c.T, cb->D, n0, n, c.quantized));
                    }
                }
            }
            if (c.want_loss)
                NAT_LAUNCH(4, st, rows::reduce_loss_kernel<<<cb->L, 1024, 0, st>>>(ws.row_loss, n, ws.loss_acc, ws.rows));
        }
        NAT_CUDA(cudaGetLastError());
        return NAT_OK;
    } else
    for (int l = 0; l < cb->L; ++l) {
        rows::UpdateArgs ua;
        ua.r = ws.r; ua.a = ws.a; ua.rowinfo = ws.rowinfo; ua.rowamax = ws.rowamax;
        ua.cb = cb->cbf + l * cb_layer_ld;
        ua.cn64 = cb->cn64 + static_cast<long long>(l) * cb->K;
        ua.lc_next = (l + 1 < cb->L) ? cb->lc + l + 1 : nullptr;
        ua.codes = static_cast<char*>(c.codes) + (static_cast<long long>(l) * c.N + n0) * cbytes;
        ua.row_loss = c.want_loss ? ws.row_loss : nullptr;
        ua.stats = c.stats ? c.stats + l * NAT_RVQ_STAT_FIELDS : nullptr;
        ua.n = n; ua.K = cb->K; ua.dp = cb->dp; ua.code_dtype = c.code_dtype;
        const int scan_grid = cb->sm_count * 4;
        const size_t scan_smem = static_cast<size_t>(cb->dp) * sizeof(float);
        if (c.temperatures != nullptr && c.temperatures[l] > 0.f) {
            rows::SampleArgs sargs;
            sargs.noise = c.noise != nullptr ? c.noise + (static_cast<long long>(l) * c.N + n0) * cb->K : nullptr;
            sargs.seed = c.seed;
            sargs.row0 = static_cast<unsigned long long>(n0);
            sargs.draw = static_cast<unsigned>(c.draw_base + l);
            sargs.temperature = c.temperatures[l];
            if (c.noise == nullptr && !c.exact && c.scores != nullptr) {
                // Philox noise: distances from the tensor-core pass (score matrix through HBM), one warp per frame
                NAT_LAUNCH(1, st, gemm::rvq_gemm_topk_kernel<true><<<std::min(n_tiles, cb->sm_count), gemm::NUM_THREADS,
                                                                  gemm::SMEM_BYTES, st>>>(
                    map_a, cb->map_b, n, n_tiles, cb->kp / gemm::BLOCK_N, cb->dp / gemm::BLOCK_K, l * cb->kp, 1, ws.rowinfo,
                    cb->cn32 + static_cast<long long>(l) * cb->kp, ws.cand, c.scores, cb->kp));
                NAT_LAUNCH(2, st, rows::sample_from_acc_kernel<<<std::min((n + 7) / 8, cb->sm_count * 16), 256, 0, st>>>(
                    ua, sargs, c.scores, cb->kp, cb->cn32 + static_cast<long long>(l) * cb->kp));
            } else {
                const size_t smem = scan_smem + static_cast<size_t>(cb->K) * sizeof(float);
                NAT_LAUNCH(3, st, rows::sample_scan_kernel<<<std::min(n, scan_grid), rows::kScanThreads, smem, st>>>(ua, sargs));
            }
        } else if (!c.exact && c.temperatures == nullptr && c.scores != nullptr) {
            // few tiles: the coarse GEMM dealt over tiles x codebook chunks, then one warp per frame (low latency)
            const int n_chunks = cb->kp / gemm::BLOCK_N;
            NAT_LAUNCH(1, st, gemm::rvq_gemm_topk_kernel<true><<<std::min(n_tiles * n_chunks, cb->sm_count), gemm::NUM_THREADS,
                                                              gemm::SMEM_BYTES, st>>>(
                map_a, cb->map_b, n, n_tiles, n_chunks, cb->dp / gemm::BLOCK_K, l * cb->kp, n_chunks, ws.rowinfo,
                cb->cn32m + static_cast<long long>(l) * cb->kp, ws.cand, c.scores, cb->kp));
            NAT_LAUNCH(2, st, rows::argmin_from_acc_kernel<<<std::min((n + 7) / 8, cb->sm_count * 16), 256, 0, st>>>(
                ua, c.scores, cb->kp, cb->cn32m + static_cast<long long>(l) * cb->kp));
        } else if (!c.exact && c.temperatures == nullptr) {
            NAT_LAUNCH(1, st, gemm::rvq_gemm_topk_kernel<false><<<std::min(n_tiles, cb->sm_count), gemm::NUM_THREADS,
                                                                   gemm::SMEM_BYTES, st>>>(
                map_a, cb->map_b, n, n_tiles, cb->kp / gemm::BLOCK_N, cb->dp / gemm::BLOCK_K, l * cb->kp, 1, ws.rowinfo,
                cb->cn32m + static_cast<long long>(l) * cb->kp, ws.cand, nullptr, 0));
            NAT_LAUNCH(2, st, rows::decide_update_kernel<<<std::min((n + 7) / 8, cb->sm_count * 16), 256, 0, st>>>(
                ua, ws.cand, ws.scan_list, ws.scan_count + l));
            NAT_LAUNCH(3, st, rows::full_scan_kernel<<<scan_grid, rows::kScanThreads, scan_smem, st>>>(
                ua, ws.scan_list, ws.scan_count + l, 0, false));
        } else {
            NAT_LAUNCH(3, st, rows::full_scan_kernel<<<std::min(n, scan_grid), rows::kScanThreads, scan_smem, st>>>(
                ua, nullptr, nullptr, n, true));
        }
        if (c.want_loss) NAT_LAUNCH(4, st, rows::reduce_loss_kernel<<<1, 1024, 0, st>>>(ws.row_loss, n, ws.loss_acc + l));
        NAT_CUDA(cudaGetLastError());
    }
    if (c.quantized != nullptr) {
        // replay the chain from the emitted codes on a fresh copy of x (bit-identical op order, nat.py:2167/1405/1408)
        const int warps_grid = std::min((n + 7) / 8, cb->sm_count * 16);
        if (c.layout == NAT_LAYOUT_ROWS) {
            NAT_LAUNCH(0, st, rows::prep_rows_kernel<<<warps_grid, 256, 0, st>>>(c.x + n0 * cb->D, cb->D, n, cb->D, cb->dp,
                                                                              ws.r, ws.a, ws.rowinfo, ws.rowamax, cb->lc, false));
        } else {
            dim3 grid((n + 31) / 32, cb->dp / 32);
            NAT_LAUNCH(0, st, rows::bct_to_rows_kernel<<<grid, 256, 0, st>>>(c.x, c.T, cb->D, n0, n, cb->dp, ws.r));
        }
        NAT_LAUNCH(5, st, rows::reconstruct_rows_kernel<<<warps_grid, 256, 0, st>>>(ws.r, ws.r, n, cb->dp, cb->cbf, cb_layer_ld,
                                                                                 cb->L, c.codes, c.code_dtype, c.N, n0));
        if (c.layout == NAT_LAYOUT_ROWS) {
            NAT_LAUNCH(5, st, rows::copy_rows_out_kernel<<<cb->sm_count * 8, 256, 0, st>>>(ws.r, n, cb->dp, cb->D,
                                                                                        c.quantized + n0 * cb->D));
        } else {
            dim3 grid((n + 31) / 32, (cb->D + 31) / 32);
            NAT_LAUNCH(5, st, rows::rows_to_bct_kernel<<<grid, 256, 0, st>>>(ws.r, cb->dp, c.T, cb->D, n0, n, c.quantized));
        }
        NAT_CUDA(cudaGetLastError());
    }
    return NAT_OK;
}

static bool overlap_enabled() {
    static bool on = [] { const char* e = getenv("NAT_RVQ_STREAMS"); return e && atoi(e) == 2; }();
    return on;
}

static int encode_impl(const nat_rvq_codebooks* cb_const, const float* x_dev, int layout, int64_t B, int64_t T,
                       void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                       float commitment_weight, unsigned long long* stats_dev, void* workspace_dev,
                       size_t workspace_bytes, int flags, void* stream, const float* temperatures,
                       const float* noise_dev, unsigned long long seed, unsigned long long draw_base) {
    using namespace nat;
    nat_rvq_codebooks* cb = const_cast<nat_rvq_codebooks*>(cb_const);
    if (cb == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null codebook handle");
    if (B < 0 || T < 0) return fail(NAT_ERR_INVALID_ARGUMENT, "negative batch or time extent");
    if (layout != NAT_LAYOUT_BCT && layout != NAT_LAYOUT_ROWS) return fail(NAT_ERR_INVALID_ARGUMENT, "bad layout %d", layout);
    if (code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16) return fail(NAT_ERR_INVALID_ARGUMENT, "bad code dtype %d", code_dtype);
    if (code_dtype == NAT_CODES_I16 && cb->K > 32768) return fail(NAT_ERR_INVALID_ARGUMENT, "int16 codes need codebook_size <= 32768");
    const long long N = B * T;
    if (N == 0) return NAT_OK;
    if (N > (1LL << 40)) return fail(NAT_ERR_UNSUPPORTED, "too many frames");
    if (x_dev == nullptr || codes_out_dev == nullptr || workspace_dev == nullptr)
        return fail(NAT_ERR_INVALID_ARGUMENT, "null device pointer");
    int dev = -1;
    NAT_CUDA(cudaGetDevice(&dev));
    if (dev != cb->device) return fail(NAT_ERR_INVALID_ARGUMENT, "codebooks live on device %d, current device is %d", cb->device, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // Two halves of the workspace on two internal streams: the HBM-bound row kernels of one half run under the
    // tensor-bound GEMM of the other. Small inputs (or NAT_RVQ_SINGLE_STREAM / NAT_RVQ_STREAMS=1) stay on `st`.
    const bool two = overlap_enabled() && !(flags & NAT_RVQ_SINGLE_STREAM) && N >= 4LL * 128 * cb->sm_count;
    // small-input path: its score matrix is carved off the tail of the workspace (when the caller sized it with
    // nat_rvq_workspace_bytes for this many frames; a smaller workspace simply keeps the fused kernel)
    size_t scores_bytes = (temperatures == nullptr && !(flags & NAT_RVQ_EXACT_SCAN) && small_path_enabled() && !two)
                              ? small_input_scores_bytes(cb, N) : 0;
    if (scores_bytes != 0) {
        const size_t main_bytes = static_cast<size_t>(round_up(kWsFixed + 256 * 10 + round_up(N, 128) * ws_per_row(cb->dp, cb->L), 256));
        if (workspace_bytes < main_bytes + scores_bytes + 256) scores_bytes = 0;
        else workspace_bytes = (workspace_bytes - scores_bytes) & ~static_cast<size_t>(255);
    }
    const int n_lanes = two ? 2 : 1;
    Workspace ws[2];
    CUtensorMap map_a[2];
    long long want = std::min<long long>((N + n_lanes - 1) / n_lanes, chunk_cap_rows());
    // Philox sampling reads its distances from a score matrix [frames, Kp] (the tensor-core pass dumps its
    // accumulators): it lives in the tail of the caller's workspace, and the chunk shrinks until both fit. (A
    // stream-ordered allocation per call cost 5-260 ms of host time: the default pool hands the gigabyte back to the
    // driver at every synchronisation.)
    const bool philox_bulk = temperatures != nullptr && noise_dev == nullptr && !(flags & NAT_RVQ_EXACT_SCAN);
    if (philox_bulk) {
        const size_t per_row = ws_per_row(cb->dp, cb->L) + static_cast<size_t>(cb->kp) * sizeof(float);
        const size_t fixed = kWsFixed + 256 * 10 + 512;
        const long long fit = workspace_bytes > fixed ? static_cast<long long>((workspace_bytes - fixed) / per_row) / 128 * 128 : 0;
        if (fit < 128)
            return fail(NAT_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one 128-frame tile and its score rows (need %zu)",
                        workspace_bytes, fixed + 128 * per_row);
        want = std::min(round_up(want, 128), fit);
        scores_bytes = static_cast<size_t>(want) * cb->kp * sizeof(float);
        workspace_bytes = (workspace_bytes - scores_bytes) & ~static_cast<size_t>(255);
    }
    const size_t half_bytes = (workspace_bytes / n_lanes) & ~static_cast<size_t>(255);
    for (int i = 0; i < n_lanes; ++i) {
        if (!carve(static_cast<char*>(workspace_dev) + i * half_bytes, half_bytes, cb->dp, cb->L, want, &ws[i]))
            return fail(NAT_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one 128-frame tile per lane (need %zu)",
                        workspace_bytes, nat_rvq_workspace_bytes(cb, 128 * n_lanes));
        if (int rc = make_map_f16(&map_a[i], ws[i].a, ws[i].rows, cb->dp, 128)) return rc;
    }
    cudaStream_t lane_st[2] = {st, st};
    if (two) {
        if (cb->side[0] == nullptr) {
            for (int i = 0; i < 2; ++i) NAT_CUDA(cudaStreamCreateWithFlags(&cb->side[i], cudaStreamNonBlocking));
            for (auto& e : cb->side_ev) NAT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        NAT_CUDA(cudaEventRecord(cb->side_ev[0], st));
        for (int i = 0; i < 2; ++i) {
            NAT_CUDA(cudaStreamWaitEvent(cb->side[i], cb->side_ev[0], 0));
            lane_st[i] = cb->side[i];
        }
    }

    EncodeCall call{cb, x_dev, layout, T, N, codes_out_dev, code_dtype, quantized_out_dev, loss_out_dev != nullptr,
                    stats_dev, (flags & NAT_RVQ_EXACT_SCAN) != 0};
    call.temperatures = temperatures; call.noise = noise_dev; call.seed = seed; call.draw_base = draw_base;
    // Score matrix (accumulator dump): the Philox sampling path, and the small-input argmin path -- inputs whose
    // tiles x codebook chunks fit one wave of SMs would otherwise run a whole stack on a few persistent CTAs.
    if (scores_bytes != 0)
        call.scores = reinterpret_cast<float*>(static_cast<char*>(workspace_dev) + workspace_bytes);   // the tail cut off above
    for (int i = 0; i < n_lanes; ++i) {
        if (loss_out_dev != nullptr) NAT_CUDA(cudaMemsetAsync(ws[i].loss_acc, 0, sizeof(double) * cb->L, lane_st[i]));
        if (cb->dp != cb->D) NAT_CUDA(cudaMemsetAsync(ws[i].a, 0, static_cast<size_t>(ws[i].rows) * cb->dp * 2, lane_st[i]));
    }
    int lane = 0;
    for (long long n0 = 0; n0 < N; lane = (lane + 1) % n_lanes) {
        const int n = static_cast<int>(std::min<long long>(ws[lane].rows, N - n0));
        if (int rc = encode_chunk(call, ws[lane], map_a[lane], n0, n, lane_st[lane])) return rc;
        n0 += n;
    }
    if (two) {
        for (int i = 0; i < 2; ++i) {
            NAT_CUDA(cudaEventRecord(cb->side_ev[1 + i], cb->side[i]));
            NAT_CUDA(cudaStreamWaitEvent(st, cb->side_ev[1 + i], 0));
        }
    }
    if (loss_out_dev != nullptr) {
        NAT_LAUNCH(4, st, rows::finish_loss_kernel<<<1, 32, 0, st>>>(ws[0].loss_acc, two ? ws[1].loss_acc : nullptr, cb->L,
                                                                   static_cast<double>(N) * cb->D, commitment_weight,
                                                                   loss_out_dev));
        NAT_CUDA(cudaGetLastError());
    }
    return NAT_OK;
}

int nat_rvq_encode_f32(const nat_rvq_codebooks* cb, const float* x_dev, int layout, int64_t B, int64_t T,
                       void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                       float commitment_weight, unsigned long long* stats_dev, void* workspace_dev,
                       size_t workspace_bytes, int flags, void* stream) {
    return encode_impl(cb, x_dev, layout, B, T, codes_out_dev, code_dtype, quantized_out_dev, loss_out_dev,
                       commitment_weight, stats_dev, workspace_dev, workspace_bytes, flags, stream, nullptr, nullptr, 0, 0);
}

int nat_rvq_sample_f32(const nat_rvq_codebooks* cb, const float* x_dev, int layout, int64_t B, int64_t T,
                       void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                       float commitment_weight, const float* temperatures_host, const float* noise_dev,
                       unsigned long long philox_seed, unsigned long long philox_draw, void* workspace_dev,
                       size_t workspace_bytes, int flags, void* stream) {
    if (cb == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null codebook handle");
    if (temperatures_host == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null temperature array");
    for (int l = 0; l < cb->L; ++l)
        if (!(temperatures_host[l] == temperatures_host[l]))
            return fail(NAT_ERR_INVALID_ARGUMENT, "temperature of layer %d is NaN", l);
    const size_t smem = static_cast<size_t>(cb->dp) * 4 + static_cast<size_t>(cb->K) * 4;
    if (smem > 200 * 1024)
        return fail(NAT_ERR_UNSUPPORTED, "sampling mode keeps K scores in shared memory: codebook_size %d is too large", cb->K);
    // (sample_scan_kernel's shared-memory limit is raised once per device when a codebook handle is created: the
    // attribute call costs milliseconds of host time, 5.5 of the 12.7 ms this call took on 270 000 frames)
    return encode_impl(cb, x_dev, layout, B, T, codes_out_dev, code_dtype, quantized_out_dev, loss_out_dev,
                       commitment_weight, nullptr, workspace_dev, workspace_bytes,
                       NAT_RVQ_SINGLE_STREAM | (flags & NAT_RVQ_EXACT_SCAN), stream,
                       temperatures_host, noise_dev, philox_seed, philox_draw);
}

static int finish_profile(Profiler& prof, int rc, cudaStream_t stream, float* prof_ms_host) {
    cudaError_t e = cudaStreamSynchronize(stream);
    for (int i = 0; i < NAT_PROF_FIELDS; ++i) prof_ms_host[i] = 0.f;
    for (auto& sp : prof.spans) {
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.cls < 6) prof_ms_host[sp.cls] += ms;
        if (sp.cls == NAT_PROF_GEMM) prof_ms_host[NAT_PROF_GEMM_LAUNCHES] += 1.f;
    }
    if (e == cudaSuccess && !prof.spans.empty())
        cudaEventElapsedTime(&prof_ms_host[NAT_PROF_WALL], prof.spans.front().a, prof.spans.back().b);
    for (auto& sp : prof.spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    if (rc != NAT_OK) return rc;
    if (e != cudaSuccess) return fail(NAT_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    return NAT_OK;
}

int nat_rvq_encode_profile_f32(const nat_rvq_codebooks* cb, const float* x_dev, int layout, int64_t B, int64_t T,
                               void* codes_out_dev, int code_dtype, float* quantized_out_dev, float* loss_out_dev,
                               float commitment_weight, unsigned long long* stats_dev, void* workspace_dev,
                               size_t workspace_bytes, int flags, void* stream, float* prof_ms_host) {
    if (prof_ms_host == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null profile buffer");
    Profiler prof;
    g_prof = &prof;
    const int rc = nat_rvq_encode_f32(cb, x_dev, layout, B, T, codes_out_dev, code_dtype, quantized_out_dev,
                                      loss_out_dev, commitment_weight, stats_dev, workspace_dev, workspace_bytes, flags,
                                      stream);
    g_prof = nullptr;
    return finish_profile(prof, rc, static_cast<cudaStream_t>(stream), prof_ms_host);
}

unsigned long long nat_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nat_rvq_decode_f32(const nat_rvq_codebooks* cb, const void* codes_dev, int code_dtype, int n_code_layers,
                       int64_t B, int64_t T, int layout, float* out_dev, void* stream) {
    using namespace nat;
    if (cb == nullptr || out_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    if (code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16) return fail(NAT_ERR_INVALID_ARGUMENT, "bad code dtype");
    const long long N = B * T;
    if (N <= 0) return NAT_OK;
    const int used = std::max(0, std::min(n_code_layers, cb->L));
    if (used > 0 && codes_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null codes");
    const size_t tile_smem = static_cast<size_t>(rows::kReplayFrames) * cb->dp * sizeof(float);
    if (layout == NAT_LAYOUT_BCT && used > 0 && tile_smem <= 200 * 1024) {
        const long long n_tiles = (N + rows::kReplayFrames - 1) / rows::kReplayFrames;
        const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(2048 / rows::kReplayThreads, (220 * 1024) / (tile_smem + 1024))));
        auto kern = used == 4 ? rows::decode_bct_kernel<4> : rows::decode_bct_kernel<0>;
        NAT_LAUNCH(5, static_cast<cudaStream_t>(stream),
                   kern<<<static_cast<int>(std::min<long long>(n_tiles, cb->sm_count * per_sm)), rows::kReplayThreads, tile_smem,
                          static_cast<cudaStream_t>(stream)>>>(cb->cbf, static_cast<long long>(cb->K) * cb->dp, cb->dp, cb->D, used,
                                                               codes_dev, code_dtype, N, T, out_dev));
        NAT_CUDA(cudaGetLastError());
        return NAT_OK;
    }
    const long long total = N * cb->D;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, cb->sm_count * 16));
    NAT_LAUNCH(5, static_cast<cudaStream_t>(stream), rows::decode_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        cb->cbf, static_cast<long long>(cb->K) * cb->dp, cb->dp, cb->D, used, codes_dev, code_dtype, N, T, layout, out_dev));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

// ------------------------------------------------------------------------------------------------- time alignment
int nat_interp_linear_f32(const float* x_dev, int64_t rows, int64_t t_in, int64_t t_out, float* out_dev, void* stream) {
    using namespace nat;
    if (rows < 0 || t_in < 1 || t_out < 0 || t_in > 0x7FFFFFFF || t_out > 0x7FFFFFFF)
        return fail(NAT_ERR_INVALID_ARGUMENT, "bad extents for linear interpolation");
    if (rows == 0 || t_out == 0) return NAT_OK;
    if (x_dev == nullptr || out_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 148;
    NAT_CUDA(cudaGetDevice(&dev));
    NAT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const float scale = static_cast<float>(t_in) / static_cast<float>(t_out);   // area_pixel_compute_scale<float>
    const long long total = rows * t_out;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, sms * 16)));
    NAT_LAUNCH(5, st, interp::interp_linear_kernel<<<grid, 256, 0, st>>>(x_dev, rows, static_cast<int>(t_in),
                                                                        static_cast<int>(t_out), scale, out_dev));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

// ------------------------------------------------------------------------------------------------- token statistics
int nat_token_histogram(const void* codes_dev, int code_dtype, int64_t n_tokens, int vocab,
                        unsigned long long* counts_dev, unsigned long long* outliers_dev, void* stream) {
    using namespace nat;
    if (code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16) return fail(NAT_ERR_INVALID_ARGUMENT, "bad code dtype %d", code_dtype);
    if (vocab < 1 || n_tokens < 0) return fail(NAT_ERR_INVALID_ARGUMENT, "bad vocabulary size or token count");
    if (counts_dev == nullptr || outliers_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null output");
    if (n_tokens == 0) return NAT_OK;
    if (codes_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null codes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 148;
    NAT_CUDA(cudaGetDevice(&dev));
    NAT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = vocab <= stats::kHistSmemBins ? static_cast<size_t>(vocab) * 4 : 0;
    const long long want = (n_tokens + stats::kHistThreads * 8 - 1) / (stats::kHistThreads * 8);
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, sms * 4)));
    NAT_LAUNCH(5, st, stats::token_histogram_kernel<<<grid, stats::kHistThreads, smem, st>>>(
        codes_dev, code_dtype, n_tokens, vocab, counts_dev, outliers_dev));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

int nat_token_joint_histogram(const void* a_dev, const void* b_dev, int code_dtype, int64_t n, const double* edges_a_dev,
                              const double* edges_b_dev, int bins, unsigned long long* hist_dev, void* stream) {
    using namespace nat;
    if (code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16) return fail(NAT_ERR_INVALID_ARGUMENT, "bad code dtype %d", code_dtype);
    if (bins < 1 || bins > stats::kJointMaxBins) return fail(NAT_ERR_UNSUPPORTED, "bins must be in [1, %d], got %d", stats::kJointMaxBins, bins);
    if (n < 0 || hist_dev == nullptr || edges_a_dev == nullptr || edges_b_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "bad argument");
    if (n == 0) return NAT_OK;
    if (a_dev == nullptr || b_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null codes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 148;
    NAT_CUDA(cudaGetDevice(&dev));
    NAT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (n + stats::kJointThreads * 8 - 1) / (stats::kJointThreads * 8);
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, sms * 4)));
    NAT_LAUNCH(5, st, stats::joint_histogram_kernel<<<grid, stats::kJointThreads, 0, st>>>(
        a_dev, b_dev, code_dtype, n, edges_a_dev, edges_b_dev, bins, hist_dev));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

int nat_debug_stack_counters(nat_rvq_codebooks* cb, int enable, unsigned long long* out_host, int max_ctas,
                              int* n_ctas, int* n_slots) {
    if (cb == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null handle");
    const size_t bytes = sizeof(unsigned long long) * cb->sm_count * nat::stack::DBG_SLOTS;
    if (cb->stack_dbg == nullptr) {
        NAT_CUDA(cudaMalloc(&cb->stack_dbg, bytes));
        NAT_CUDA(cudaMemset(cb->stack_dbg, 0, bytes));
    }
    if (out_host != nullptr) {
        NAT_CUDA(cudaDeviceSynchronize());                     // debug hook, not a hot path
        const int n = std::min(max_ctas, cb->sm_count);
        NAT_CUDA(cudaMemcpy(out_host, cb->stack_dbg, sizeof(unsigned long long) * n * nat::stack::DBG_SLOTS,
                            cudaMemcpyDeviceToHost));
        NAT_CUDA(cudaMemset(cb->stack_dbg, 0, bytes));
    }
    cb->stack_dbg_on = enable != 0;
    if (n_ctas) *n_ctas = cb->sm_count;
    if (n_slots) *n_slots = nat::stack::DBG_SLOTS;
    return NAT_OK;
}

int nat_debug_rvq_scores(const nat_rvq_codebooks* cb, int layer, const float* rows_dev, int64_t N,
                         float* scores_out_dev, float* row_scale_out_dev, float* cb_scale_out_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace nat;
    if (cb == nullptr || rows_dev == nullptr || scores_out_dev == nullptr || workspace_dev == nullptr)
        return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    if (layer < 0 || layer >= cb->L) return fail(NAT_ERR_INVALID_ARGUMENT, "layer out of range");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace ws;
    if (!carve(workspace_dev, workspace_bytes, cb->dp, cb->L, N, &ws) || ws.rows < N)
        return fail(NAT_ERR_WORKSPACE, "debug call needs a workspace for all %lld rows", (long long)N);
    CUtensorMap map_a;
    if (int rc = make_map_f16(&map_a, ws.a, ws.rows, cb->dp, 128)) return rc;
    const int n = static_cast<int>(N);
    if (cb->dp != cb->D) NAT_CUDA(cudaMemsetAsync(ws.a, 0, static_cast<size_t>(ws.rows) * cb->dp * 2, st));
    NAT_LAUNCH(0, st, rows::prep_rows_kernel<<<std::min((n + 7) / 8, cb->sm_count * 16), 256, 0, st>>>(
        rows_dev, cb->D, n, cb->D, cb->dp, ws.r, ws.a, ws.rowinfo, ws.rowamax, cb->lc + layer, false));
    const int n_tiles = (n + gemm::BLOCK_M - 1) / gemm::BLOCK_M;
    NAT_LAUNCH(1, st, gemm::rvq_gemm_topk_kernel<true><<<std::min(n_tiles, cb->sm_count), gemm::NUM_THREADS, gemm::SMEM_BYTES, st>>>(
        map_a, cb->map_b, n, n_tiles, cb->kp / gemm::BLOCK_N, cb->dp / gemm::BLOCK_K, layer * cb->kp, 1, ws.rowinfo,
        cb->cn32 + static_cast<long long>(layer) * cb->kp, ws.cand, scores_out_dev, cb->kp));
    NAT_CUDA(cudaGetLastError());
    if (row_scale_out_dev)
        NAT_CUDA(cudaMemcpy2DAsync(row_scale_out_dev, 4, reinterpret_cast<const char*>(ws.rowinfo) + 12, 16, 4, n,
                                   cudaMemcpyDeviceToDevice, st));
    if (cb_scale_out_dev)
        NAT_CUDA(cudaMemcpyAsync(cb_scale_out_dev, &cb->lc[layer].sc, 4, cudaMemcpyDeviceToDevice, st));
    return NAT_OK;
}

// ------------------------------------------------------------------------------------------------- several stacks
// The stacks of one tokenizer (S0-S3, A0-A3) in one call. Stacks that quantise the same frames (equal x pointers)
// share one layer-0 preparation, and the second stack's persistent launch is a programmatic dependent launch that
// fills the SMs the first one has already left. Shapes the fused kernel does not cover (different K or D, D > 1024,
// inputs of a few tiles, NAT_RVQ_FUSED=0) run stack after stack through nat_rvq_encode_f32.
namespace {

bool stacks_fusable(const nat_rvq_codebooks* const* stacks, int n_stacks, long long N) {
    if (n_stacks != 2 || !fused_enabled()) return false;
    const nat_rvq_codebooks *a = stacks[0], *b = stacks[1];
    if (a->D != b->D || a->K != b->K || a->dp > 1024 || a->device != b->device) return false;
    if (small_path_enabled() && small_input_scores_bytes(a, N) != 0) return false;     // latency path wins there
    return true;
}

struct MultiWs {
    float* r_prep[2]; __half* a_prep[2]; float* rowamax[2]; float4* rowinfo[2];
    float* r_work[2]; __half* a_work[2]; float4* rowinfo_work[2]; float* rowamax_work[2];
    long long rows;
};

// per frame: prepared rows (fp32 + fp16 + max) per distinct input, a layer-0 window per stack, and the rows the
// update warps write (fp32 + fp16 + window + max) for each of the two stacks
size_t multi_per_row(int dp, int n_inputs) {
    return static_cast<size_t>(n_inputs) * (static_cast<size_t>(dp) * 6 + 4) + 2 * 16 +
           2 * (static_cast<size_t>(dp) * 6 + 16 + 4);
}

bool carve_multi(void* base, size_t bytes, int dp, int n_inputs, long long want_rows, MultiWs* ws) {
    const size_t slack = 256 * 20;
    if (bytes <= slack) return false;
    long long rows = static_cast<long long>((bytes - slack) / multi_per_row(dp, n_inputs));
    rows = std::min(rows, round_up(want_rows, 128)) / 128 * 128;
    if (rows < 128) return false;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t n) { char* q = p; p += round_up((long long)n, 256); return q; };
    for (int i = 0; i < 2; ++i) {
        const int src = i < n_inputs ? i : 0;
        if (i < n_inputs) {
            ws->r_prep[i] = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * dp * 4));
            ws->a_prep[i] = reinterpret_cast<__half*>(take(static_cast<size_t>(rows) * dp * 2));
            ws->rowamax[i] = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * 4));
        } else {
            ws->r_prep[i] = ws->r_prep[src]; ws->a_prep[i] = ws->a_prep[src]; ws->rowamax[i] = ws->rowamax[src];
        }
        ws->rowinfo[i] = reinterpret_cast<float4*>(take(static_cast<size_t>(rows) * 16));
    }
    for (int i = 0; i < 2; ++i) {
        ws->r_work[i] = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * dp * 4));
        ws->a_work[i] = reinterpret_cast<__half*>(take(static_cast<size_t>(rows) * dp * 2));
        ws->rowinfo_work[i] = reinterpret_cast<float4*>(take(static_cast<size_t>(rows) * 16));
        ws->rowamax_work[i] = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * 4));
    }
    ws->rows = rows;
    return static_cast<size_t>(p - static_cast<char*>(base)) <= bytes;
}

size_t multi_ws_bytes(int dp, int n_inputs, long long n_frames) {
    const long long rows = std::min<long long>(round_up(n_frames, 128), chunk_cap_rows());
    return static_cast<size_t>(rows) * multi_per_row(dp, n_inputs) + 256 * 20;
}

}  // namespace

size_t nat_rvq_stacks_workspace_bytes(const nat_rvq_codebooks* const* stacks, int n_stacks, int64_t n_frames) {
    size_t need = 0;
    if (stacks == nullptr) return nat_rvq_workspace_bytes(nullptr, n_frames);
    for (int i = 0; i < n_stacks; ++i) need = std::max(need, nat_rvq_workspace_bytes(stacks[i], n_frames));
    if (n_stacks == 2 && stacks[0] != nullptr && stacks[1] != nullptr && n_frames > 0 &&
        stacks[0]->D == stacks[1]->D && stacks[0]->K == stacks[1]->K && stacks[0]->dp <= 1024)
        need = std::max(need, multi_ws_bytes(stacks[0]->dp, 2, n_frames));
    return need;
}

int nat_rvq_stacks_fused(const nat_rvq_codebooks* const* stacks, int n_stacks, int64_t n_frames) {
    if (stacks == nullptr || n_stacks < 1) return 0;
    for (int i = 0; i < n_stacks; ++i) if (stacks[i] == nullptr) return 0;
    return stacks_fusable(stacks, n_stacks, n_frames) ? 1 : 0;
}

static int encode_stacks_impl(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                              const int64_t* t_in, int layout, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                              void* workspace_dev, size_t workspace_bytes, int flags, void* stream);

int nat_rvq_encode_stacks_f32(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                              int layout, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                              void* workspace_dev, size_t workspace_bytes, int flags, void* stream) {
    return encode_stacks_impl(stacks, n_stacks, x_dev, nullptr, layout, B, T, codes_out_dev, code_dtype, workspace_dev,
                              workspace_bytes, flags, stream);
}

int nat_rvq_encode_stacks_aligned_f32(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                                      const int64_t* t_in, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                                      void* workspace_dev, size_t workspace_bytes, int flags, void* stream) {
    if (t_in == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null time extents");
    for (int i = 0; i < n_stacks; ++i)
        if (t_in[i] < 1 || t_in[i] > 0x7FFFFFFF || T > 0x7FFFFFFF) return fail(NAT_ERR_INVALID_ARGUMENT, "bad time extent");
    return encode_stacks_impl(stacks, n_stacks, x_dev, t_in, NAT_LAYOUT_BCT, B, T, codes_out_dev, code_dtype, workspace_dev,
                              workspace_bytes, flags, stream);
}

static int encode_stacks_impl(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                              const int64_t* t_in, int layout, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                              void* workspace_dev, size_t workspace_bytes, int flags, void* stream) {
    using namespace nat;
    if (stacks == nullptr || x_dev == nullptr || n_stacks < 1) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    for (int i = 0; i < n_stacks; ++i)
        if (stacks[i] == nullptr || x_dev[i] == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "stack %d: null handle or input", i);
    if (B < 0 || T < 0) return fail(NAT_ERR_INVALID_ARGUMENT, "negative batch or time extent");
    if (layout != NAT_LAYOUT_BCT && layout != NAT_LAYOUT_ROWS) return fail(NAT_ERR_INVALID_ARGUMENT, "bad layout %d", layout);
    if (code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16) return fail(NAT_ERR_INVALID_ARGUMENT, "bad code dtype %d", code_dtype);
    const long long N = B * T;
    if (N == 0) return NAT_OK;
    if (codes_out_dev == nullptr || workspace_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null device pointer");
    const int cbytes = code_bytes(code_dtype);
    if (!stacks_fusable(stacks, n_stacks, N) || (flags & NAT_RVQ_EXACT_SCAN)) {
        if (t_in != nullptr)
            for (int i = 0; i < n_stacks; ++i)
                if (t_in[i] != T)
                    return fail(NAT_ERR_UNSUPPORTED, "time-base alignment is fused into the two-stack preparation only "
                                "(nat_rvq_stacks_fused() says when); align with nat_interp_linear_f32 first");
        long long layer0 = 0;
        for (int i = 0; i < n_stacks; ++i) {
            if (int rc = nat_rvq_encode_f32(stacks[i], x_dev[i], layout, B, T, static_cast<char*>(codes_out_dev) + layer0 * N * cbytes,
                                            code_dtype, nullptr, nullptr, 0.25f, nullptr, workspace_dev, workspace_bytes,
                                            flags | NAT_RVQ_SINGLE_STREAM, stream)) return rc;
            layer0 += stacks[i]->L;
        }
        return NAT_OK;
    }
    const nat_rvq_codebooks *s0 = stacks[0], *s1 = stacks[1];
    if (code_dtype == NAT_CODES_I16 && s0->K > 32768) return fail(NAT_ERR_INVALID_ARGUMENT, "int16 codes need codebook_size <= 32768");
    int dev = -1;
    NAT_CUDA(cudaGetDevice(&dev));
    if (dev != s0->device) return fail(NAT_ERR_INVALID_ARGUMENT, "codebooks live on device %d, current device is %d", s0->device, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long tin0 = (t_in != nullptr && t_in[0] != T) ? t_in[0] : 0, tin1 = (t_in != nullptr && t_in[1] != T) ? t_in[1] : 0;
    const int n_inputs = (x_dev[0] == x_dev[1] && tin0 == tin1) ? 1 : 2;
    MultiWs ws;
    if (!carve_multi(workspace_dev, workspace_bytes, s0->dp, n_inputs, std::min<long long>(N, chunk_cap_rows()), &ws))
        return fail(NAT_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one 128-frame tile (need %zu)", workspace_bytes,
                    multi_ws_bytes(s0->dp, n_inputs, 128));
    CUtensorMap map_prep[2], map_work[2];
    if (int rc = make_map_f16(&map_prep[0], ws.a_prep[0], ws.rows, s0->dp, 128)) return rc;
    if (n_inputs == 2) { if (int rc = make_map_f16(&map_prep[1], ws.a_prep[1], ws.rows, s0->dp, 128)) return rc; }
    else map_prep[1] = map_prep[0];
    for (int i = 0; i < 2; ++i)
        if (int rc = make_map_f16(&map_work[i], ws.a_work[i], ws.rows, s0->dp, 128)) return rc;
    if (s0->dp != s0->D) {          // padded columns of the operand rows are never written by the row kernels
        for (int i = 0; i < n_inputs; ++i) NAT_CUDA(cudaMemsetAsync(ws.a_prep[i], 0, static_cast<size_t>(ws.rows) * s0->dp * 2, st));
        for (int i = 0; i < 2; ++i) NAT_CUDA(cudaMemsetAsync(ws.a_work[i], 0, static_cast<size_t>(ws.rows) * s0->dp * 2, st));
    }
    for (long long n0 = 0; n0 < N; n0 += ws.rows) {
        const int n = static_cast<int>(std::min<long long>(ws.rows, N - n0));
        Workspace w0;
        memset(&w0, 0, sizeof w0);
        w0.r = ws.r_prep[0]; w0.a = ws.a_prep[0]; w0.rowinfo = ws.rowinfo[0]; w0.rowamax = ws.rowamax[0]; w0.rows = ws.rows;
        if (n_inputs == 1) {
            if (int rc = launch_layer0_prep(s0, w0, x_dev[0], layout, T, n0, n, st, ws.rowinfo[1], s1->lc, tin0)) return rc;
        } else {
            if (int rc = launch_layer0_prep(s0, w0, x_dev[0], layout, T, n0, n, st, nullptr, nullptr, tin0)) return rc;
            Workspace w1 = w0;
            w1.r = ws.r_prep[1]; w1.a = ws.a_prep[1]; w1.rowinfo = ws.rowinfo[1]; w1.rowamax = ws.rowamax[1];
            if (int rc = launch_layer0_prep(s1, w1, x_dev[1], layout, T, n0, n, st, nullptr, nullptr, tin1)) return rc;
        }
        for (int i = 0; i < 2; ++i) {
            StackLaunch sl;
            sl.cb = stacks[i];
            sl.r0 = ws.r_prep[i]; sl.rowinfo0 = ws.rowinfo[i]; sl.rowamax0 = ws.rowamax[i];
            sl.r = ws.r_work[i]; sl.a = ws.a_work[i]; sl.rowinfo = ws.rowinfo_work[i]; sl.rowamax = ws.rowamax_work[i];
            sl.codes = static_cast<char*>(codes_out_dev) + (i == 0 ? 0 : static_cast<long long>(s0->L) * N * cbytes);
            sl.codes_ld = N; sl.code_off = n0;
            sl.map_a0 = map_prep[i]; sl.map_a = map_work[i];
            sl.n_rows = n; sl.code_dtype = code_dtype;
            sl.programmatic = i == 1;       // shares only the prepared rows with stack 0's launch, which never writes them
            if (int rc = launch_fused_stack(sl, st)) return rc;
        }
    }
    return NAT_OK;
}

int nat_rvq_encode_stacks_profile_f32(const nat_rvq_codebooks* const* stacks, int n_stacks, const float* const* x_dev,
                                      int layout, int64_t B, int64_t T, void* codes_out_dev, int code_dtype,
                                      void* workspace_dev, size_t workspace_bytes, int flags, void* stream,
                                      float* prof_ms_host) {
    if (prof_ms_host == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null profile buffer");
    Profiler prof;
    g_prof = &prof;
    const int rc = nat_rvq_encode_stacks_f32(stacks, n_stacks, x_dev, layout, B, T, codes_out_dev, code_dtype,
                                             workspace_dev, workspace_bytes, flags, stream);
    g_prof = nullptr;
    return finish_profile(prof, rc, static_cast<cudaStream_t>(stream), prof_ms_host);
}

// ------------------------------------------------------------------------------------------------- host buffers
int nat_host_ctx_create(nat_host_ctx** out) {
    if (out == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    nat_host_ctx* ctx = new nat_host_ctx();
    cudaError_t e = cudaGetDevice(&ctx->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (auto& ev : ctx->ev)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        nat_host_ctx_destroy(ctx);
        return fail(NAT_ERR_CUDA, "host context: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return NAT_OK;
}

int nat_host_ctx_destroy(nat_host_ctx* ctx) {
    if (ctx == nullptr) return NAT_OK;
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    delete ctx;
    return NAT_OK;
}

// Host-buffer entry point: frames stream through the context's two-slot device arena, so the H2D copy of chunk i+1
// overlaps the kernels of chunk i; every chunk is uploaded ONCE and all stacks run on it; the index streams of all
// stacks come back as one [sum L, N] host array.
int nat_tokenize_host_f32(nat_host_ctx* ctx, const nat_rvq_codebooks* const* stacks, int n_stacks, const float* x_host,
                          int layout, int64_t B, int64_t T, void* codes_out_host, int code_dtype, void* stream) {
    if (ctx == nullptr || stacks == nullptr || n_stacks < 1 || n_stacks > kMaxStacks || x_host == nullptr ||
        codes_out_host == nullptr)
        return fail(NAT_ERR_INVALID_ARGUMENT, "null argument or unsupported stack count");
    if (code_dtype < NAT_CODES_I64 || code_dtype > NAT_CODES_I16) return fail(NAT_ERR_INVALID_ARGUMENT, "bad code dtype");
    if (layout != NAT_LAYOUT_BCT && layout != NAT_LAYOUT_ROWS) return fail(NAT_ERR_INVALID_ARGUMENT, "bad layout %d", layout);
    int L_total = 0;
    for (int i = 0; i < n_stacks; ++i) {
        if (stacks[i] == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "stack %d: null handle", i);
        if (stacks[i]->D != stacks[0]->D) return fail(NAT_ERR_INVALID_ARGUMENT, "stacks fed from one host buffer must share input_dim");
        L_total += stacks[i]->L;
    }
    const long long N = B * T;
    if (N <= 0) return NAT_OK;
    int dev = -1;
    NAT_CUDA(cudaGetDevice(&dev));
    if (dev != ctx->device) return fail(NAT_ERR_INVALID_ARGUMENT, "host context belongs to device %d, current device is %d", ctx->device, dev);
    const int D = stacks[0]->D;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cbytes = code_bytes(code_dtype);
    // slot = [x chunk | codes chunk]; plus one shared workspace
    long long rows = 1 << 16;
    { const char* e = getenv("NAT_HOST_CHUNK_ROWS"); if (e && atoll(e) >= 128) rows = round_up(atoll(e), 128); }
    rows = std::min<long long>(round_up(N, 128), rows);
    const size_t x_slot = static_cast<size_t>(rows) * D * 4, c_slot = static_cast<size_t>(rows) * L_total * cbytes;
    const size_t ws_bytes = nat_rvq_stacks_workspace_bytes(stacks, n_stacks, rows);
    const size_t need = 2 * (round_up(x_slot, 256) + round_up(c_slot, 256)) + ws_bytes;
    if (ctx->arena_bytes < need) {
        NAT_CUDA(cudaStreamSynchronize(st));
        NAT_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->arena) NAT_CUDA(cudaFree(ctx->arena));
        ctx->arena = nullptr; ctx->arena_bytes = 0;
        NAT_CUDA(cudaMalloc(&ctx->arena, need));
        ctx->arena_bytes = need;
    }
    char* base = static_cast<char*>(ctx->arena);
    float* xs[2]; char* cs[2];
    for (int s = 0; s < 2; ++s) { xs[s] = reinterpret_cast<float*>(base); base += round_up(x_slot, 256);
                                  cs[s] = base; base += round_up(c_slot, 256); }
    void* wsp = base;
    const float* xin[kMaxStacks];
    // copy stream: H2D(i) ; compute stream waits ev[s], encodes, D2H codes; copy stream waits ev[2+s] before reuse.
    NAT_CUDA(cudaEventRecord(ctx->ev[2], st)); NAT_CUDA(cudaEventRecord(ctx->ev[3], st));
    int slot = 0;
    // The BCT layout is strided per frame range, so it is staged with a 2-D copy per batch item; rows are contiguous.
    const long long outer = layout == NAT_LAYOUT_ROWS ? 1 : B, inner = layout == NAT_LAYOUT_ROWS ? N : T;
    for (long long b = 0; b < outer; ++b) {
        for (long long t0 = 0; t0 < inner; t0 += rows, slot ^= 1) {
            const long long n = std::min(rows, inner - t0);
            NAT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[2 + slot], 0));
            if (layout == NAT_LAYOUT_ROWS)
                NAT_CUDA(cudaMemcpyAsync(xs[slot], x_host + t0 * D, static_cast<size_t>(n) * D * 4, cudaMemcpyHostToDevice,
                                         ctx->copy_stream));
            else       // x[b, :, t0:t0+n] is D rows of n floats with pitch T
                NAT_CUDA(cudaMemcpy2DAsync(xs[slot], static_cast<size_t>(n) * 4, x_host + (b * D) * T + t0,
                                           static_cast<size_t>(T) * 4, static_cast<size_t>(n) * 4, D,
                                           cudaMemcpyHostToDevice, ctx->copy_stream));
            NAT_CUDA(cudaEventRecord(ctx->ev[slot], ctx->copy_stream));
            NAT_CUDA(cudaStreamWaitEvent(st, ctx->ev[slot], 0));
            for (int i = 0; i < n_stacks; ++i) xin[i] = xs[slot];
            if (int rc = nat_rvq_encode_stacks_f32(stacks, n_stacks, xin, layout, 1, n, cs[slot], code_dtype, wsp, ws_bytes,
                                                   NAT_RVQ_SINGLE_STREAM, st)) return rc;
            NAT_CUDA(cudaMemcpy2DAsync(static_cast<char*>(codes_out_host) + (b * inner + t0) * cbytes,
                                       static_cast<size_t>(N) * cbytes, cs[slot], static_cast<size_t>(n) * cbytes,
                                       static_cast<size_t>(n) * cbytes, L_total, cudaMemcpyDeviceToHost, st));
            NAT_CUDA(cudaEventRecord(ctx->ev[2 + slot], st));
        }
    }
    NAT_CUDA(cudaStreamSynchronize(st));     // host buffers are the caller's: results must have landed on return
    return NAT_OK;
}

// One stack, no caller context (kept from ABI 1): the handle's private context, one call at a time.
int nat_rvq_encode_host_f32(const nat_rvq_codebooks* cb_const, const float* x_host, int layout, int64_t B, int64_t T,
                            void* codes_out_host, int code_dtype, void* stream) {
    nat_rvq_codebooks* cb = const_cast<nat_rvq_codebooks*>(cb_const);
    if (cb == nullptr || x_host == nullptr || codes_out_host == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(*cb->host_mutex);
    if (cb->host_ctx == nullptr)
        if (int rc = nat_host_ctx_create(&cb->host_ctx)) return rc;
    const nat_rvq_codebooks* one[1] = {cb};
    return nat_tokenize_host_f32(cb->host_ctx, one, 1, x_host, layout, B, T, codes_out_host, code_dtype, stream);
}

// ------------------------------------------------------------------------------------------------- front-end
namespace {

// A filterbank in the form the mel kernel reads, one device allocation of nat_mel_filterbank_bytes(n_mels):
// band-major weights [n_mels, NBINS] and {first, one past last} non-zero bin per band (fb_to_banded_kernel), then the
// projection layout and its tap-major weights (fb_layout_kernel).
struct FbBlob {
    size_t off_band, off_layout, off_wpack, bytes;
    explicit FbBlob(int n_mels) {
        off_band = static_cast<size_t>(round_up(static_cast<long long>(n_mels) * nat::fe::NBINS * 4, 256));
        off_layout = off_band + static_cast<size_t>(round_up(static_cast<long long>(sizeof(int2)) * n_mels, 256));
        off_wpack = off_layout + static_cast<size_t>(round_up(nat::fe::LAYOUT_INTS * 4LL, 256));
        bytes = off_wpack + static_cast<size_t>(round_up(nat::fe::fb_wpack_capacity(n_mels) * 4, 256));
    }
    float* fbT(void* base) const { return reinterpret_cast<float*>(base); }
    int2* band(void* base) const { return reinterpret_cast<int2*>(static_cast<char*>(base) + off_band); }
    int* layout(void* base) const { return reinterpret_cast<int*>(static_cast<char*>(base) + off_layout); }
    float* wpack(void* base) const { return reinterpret_cast<float*>(static_cast<char*>(base) + off_wpack); }
};

struct FePlan {
    float2* tw = nullptr;        // [NFFT/2]
    char* fb = nullptr;          // built-in HTK filterbank in the prepared form (see FbBlob)
    int sm_count = 148;
    // cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remembered per plan (plans are keyed by device)
    std::atomic<size_t> mel_smem_set{0};
    std::atomic<bool> spectral_smem_set{false};
    FePlan() = default;
    FePlan(const FePlan& o) : tw(o.tw), fb(o.fb), sm_count(o.sm_count),
                              mel_smem_set(o.mel_smem_set.load()), spectral_smem_set(o.spectral_smem_set.load()) {}
};

std::mutex g_plan_mutex;
std::map<std::tuple<int, int, int>, FePlan> g_plans;      // (device, sample_rate, n_mels); n_mels 0 = twiddles only

double hz_to_mel_htk(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }
double mel_to_hz_htk(double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); }

int get_plan(int sample_rate, int n_mels, FePlan** out) {
    int dev = 0;
    NAT_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto key = std::make_tuple(dev, sample_rate, n_mels);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = &it->second; return NAT_OK; }
    FePlan plan;
    NAT_CUDA(cudaDeviceGetAttribute(&plan.sm_count, cudaDevAttrMultiProcessorCount, dev));
    const int N = nat::fe::NFFT, NB = nat::fe::NBINS;
    // [N/2] e^{-2 pi i k / N} (the window), then the per-pass twiddle tables in the order the threads read them
    std::vector<float2> tw(N / 2 + nat::fe::TW1_ELEMS + nat::fe::TW2_ELEMS);
    for (int k = 0; k < N / 2; ++k) tw[k] = nat::fe::fft_twiddle_value(k, N);
    for (int k = 1; k < 16; ++k) {
        for (int t = 0; t < nat::fe::TEAM; ++t)
            tw[N / 2 + (k - 1) * nat::fe::TEAM + t] = nat::fe::fft_twiddle_value(static_cast<long long>(t) * k, N);
        for (int n2 = 0; n2 < 8; ++n2)
            tw[N / 2 + nat::fe::TW1_ELEMS + (k - 1) * 8 + n2] = nat::fe::fft_twiddle_value(static_cast<long long>(n2) * k, 128);
    }
    NAT_CUDA(cudaMalloc(&plan.tw, sizeof(float2) * tw.size()));
    NAT_CUDA(cudaMemcpy(plan.tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    if (n_mels > 0) {
        // HTK triangles exactly as torchaudio.functional.melscale_fbanks(norm=None) defines them (f_min 0, f_max sr//2)
        std::vector<float> fbT(static_cast<size_t>(n_mels) * NB, 0.f);
        std::vector<int2> band(n_mels);
        const double f_max = static_cast<double>(sample_rate / 2);
        const double m_min = hz_to_mel_htk(0.0), m_max = hz_to_mel_htk(f_max);
        std::vector<double> f_pts(n_mels + 2);
        for (int i = 0; i < n_mels + 2; ++i) f_pts[i] = mel_to_hz_htk(m_min + (m_max - m_min) * i / (n_mels + 1));
        for (int m = 0; m < n_mels; ++m) {
            int lo = NB, hi = 0;
            for (int k = 0; k < NB; ++k) {
                const double f = f_max * k / (NB - 1);
                const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
                const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
                const double v = std::max(0.0, std::min(down, up));
                fbT[static_cast<size_t>(m) * NB + k] = static_cast<float>(v);
                if (v > 0.0) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
            }
            band[m] = make_int2(std::min(lo, hi), hi);
        }
        const FbBlob blob(n_mels);
        NAT_CUDA(cudaMalloc(&plan.fb, blob.bytes));
        NAT_CUDA(cudaMemcpy(blob.fbT(plan.fb), fbT.data(), fbT.size() * 4, cudaMemcpyHostToDevice));
        NAT_CUDA(cudaMemcpy(blob.band(plan.fb), band.data(), sizeof(int2) * n_mels, cudaMemcpyHostToDevice));
        nat::fe::fb_layout_kernel<<<1, nat::fe::MAX_MELS>>>(blob.fbT(plan.fb), blob.band(plan.fb), n_mels,
                                                           blob.layout(plan.fb), blob.wpack(plan.fb));
        NAT_CUDA(cudaGetLastError());
        NAT_CUDA(cudaDeviceSynchronize());
    }
    auto ins = g_plans.emplace(key, plan);
    *out = &ins.first->second;
    return NAT_OK;
}

}  // namespace

int64_t nat_mel_num_frames(int64_t S, int hop) { return hop > 0 ? 1 + S / hop : 0; }
int64_t nat_spectral_num_frames(int64_t S, int n_fft, int hop) {
    if (hop <= 0) return 0;
    return S >= n_fft ? 1 + (S - n_fft) / hop : 1;
}

// Prepared form of a dense [n_fft/2+1, n_mels] filterbank (FbBlob): derived once per transform object and handed to
// nat_mel_power_banded_f32.
size_t nat_mel_filterbank_bytes(int n_mels) { return n_mels > 0 ? FbBlob(n_mels).bytes : 0; }

static int prepare_filterbank(const float* fb_dev, int n_mels, void* out_dev, cudaStream_t st) {
    using namespace nat;
    const FbBlob blob(n_mels);
    NAT_LAUNCH(7, st, fe::fb_to_banded_kernel<<<n_mels, 128, 0, st>>>(fb_dev, n_mels, blob.fbT(out_dev), blob.band(out_dev)));
    NAT_CUDA(cudaGetLastError());
    NAT_LAUNCH(7, st, fe::fb_layout_kernel<<<1, fe::MAX_MELS, 0, st>>>(blob.fbT(out_dev), blob.band(out_dev), n_mels,
                                                                     blob.layout(out_dev), blob.wpack(out_dev)));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

int nat_mel_filterbank_prepare(const float* fb_dev, int n_mels, void* banded_out_dev, void* stream) {
    using namespace nat;
    if (fb_dev == nullptr || banded_out_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null device pointer");
    if (n_mels < 1 || n_mels > fe::MAX_MELS) return fail(NAT_ERR_UNSUPPORTED, "unsupported n_mels %d", n_mels);
    return prepare_filterbank(fb_dev, n_mels, banded_out_dev, static_cast<cudaStream_t>(stream));
}

static int mel_power_impl(const float* wave_dev, int64_t B, int64_t S, int sample_rate, int n_fft, int hop, int n_mels,
                          const float* fb_dev, const void* fb_banded_dev, float* mel_out_dev, float* logmel_out_dev,
                          void* stream);

int nat_mel_power_f32(const float* wave_dev, int64_t B, int64_t S, int sample_rate, int n_fft, int hop, int n_mels,
                      const float* fb_dev, float* mel_out_dev, float* logmel_out_dev, void* stream) {
    return mel_power_impl(wave_dev, B, S, sample_rate, n_fft, hop, n_mels, fb_dev, nullptr, mel_out_dev, logmel_out_dev, stream);
}

int nat_mel_power_banded_f32(const float* wave_dev, int64_t B, int64_t S, int sample_rate, int n_fft, int hop, int n_mels,
                             const void* fb_banded_dev, float* mel_out_dev, float* logmel_out_dev, void* stream) {
    if (fb_banded_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null banded filterbank");
    return mel_power_impl(wave_dev, B, S, sample_rate, n_fft, hop, n_mels, nullptr, fb_banded_dev, mel_out_dev, logmel_out_dev, stream);
}

static int mel_power_impl(const float* wave_dev, int64_t B, int64_t S, int sample_rate, int n_fft, int hop, int n_mels,
                          const float* fb_dev, const void* fb_banded_dev, float* mel_out_dev, float* logmel_out_dev,
                          void* stream) {
    using namespace nat;
    if (wave_dev == nullptr || mel_out_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null device pointer");
    if (n_fft != fe::NFFT) return fail(NAT_ERR_UNSUPPORTED, "n_fft must be 2048 (the reference hard-codes it), got %d", n_fft);
    if (hop < 1 || n_mels < 1 || n_mels > fe::MAX_MELS || sample_rate < 2) return fail(NAT_ERR_UNSUPPORTED, "unsupported hop/n_mels/sample_rate");
    if (B < 0) return fail(NAT_ERR_INVALID_ARGUMENT, "negative batch");
    if (B == 0) return NAT_OK;
    if (S <= n_fft / 2) return fail(NAT_ERR_INVALID_ARGUMENT, "reflect padding needs more than n_fft/2 = %d samples, got %lld", n_fft / 2, (long long)S);
    FePlan* plan = nullptr;
    if (int rc = get_plan(sample_rate, n_mels, &plan)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fe::MelArgs p;
    p.wave = wave_dev; p.S = S; p.T = nat_mel_num_frames(S, hop); p.hop = hop; p.n_mels = n_mels; p.tw = plan->tw;
    p.mel = mel_out_dev; p.logmel = logmel_out_dev;
    p.inv_wsum = 1.0f / (3.0f * fe::NFFT / 8.0f);                 // sum of hann^2 over a period = 3N/8
    // A caller-supplied dense filterbank is prepared into scratch that belongs to THIS call (stream-ordered
    // allocation): concurrent calls on different streams never share it.
    const FbBlob blob(n_mels);
    char* fb_scratch = nullptr;
    const void* fb_use = plan->fb;
    if (fb_banded_dev != nullptr) {
        fb_use = fb_banded_dev;
    } else if (fb_dev != nullptr) {
        NAT_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&fb_scratch), blob.bytes, st));
        if (int rc = prepare_filterbank(fb_dev, n_mels, fb_scratch, st)) { cudaFreeAsync(fb_scratch, st); return rc; }
        fb_use = fb_scratch;
    }
    p.layout = blob.layout(const_cast<void*>(fb_use));
    p.wpack = blob.wpack(const_cast<void*>(fb_use));
    const long long groups_per_clip = (p.T + fe::FRAMES_PER_CTA - 1) / fe::FRAMES_PER_CTA;
    const long long total = groups_per_clip * B;
    const size_t smem = fe::mel_smem_bytes(n_mels);
    // The attribute belongs to the function (per device), not to a plan: every plan raises it to the same value, the
    // largest any n_mels needs, so that no plan can lower it under another one's launch.
    const size_t smem_cap = fe::mel_smem_bytes(fe::MAX_MELS);
    if (plan->mel_smem_set.load() < smem_cap) {
        NAT_CUDA(cudaFuncSetAttribute(fe::mel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_cap)));
        plan->mel_smem_set.store(smem_cap);
    }
    const int grid = static_cast<int>(std::min<long long>(total, plan->sm_count * 4LL * 4));
    NAT_LAUNCH(7, st, fe::mel_power_kernel<<<grid, fe::THREADS, smem, st>>>(p, groups_per_clip, total));
    const cudaError_t launch_err = cudaGetLastError();
    if (fb_scratch != nullptr) cudaFreeAsync(fb_scratch, st);
    NAT_CUDA(launch_err);
    return NAT_OK;
}

int nat_spectral_stats_f32(const float* wave_dev, int64_t S, int sample_rate, int n_fft, int hop, float* out_dev,
                           void* stream) {
    using namespace nat;
    if (wave_dev == nullptr || out_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null device pointer");
    if (n_fft != fe::NFFT) return fail(NAT_ERR_UNSUPPORTED, "n_fft must be 2048, got %d", n_fft);
    if (hop < 1 || S < 1 || sample_rate < 1) return fail(NAT_ERR_INVALID_ARGUMENT, "bad hop/length/sample_rate");
    FePlan* plan = nullptr;
    if (int rc = get_plan(sample_rate, 0, &plan)) return rc;
    fe::SpectralArgs p;
    p.wave = wave_dev; p.S = S; p.T = nat_spectral_num_frames(S, n_fft, hop); p.hop = hop;
    p.bin_hz = static_cast<float>(sample_rate) / fe::NFFT; p.tw = plan->tw; p.out = out_dev;
    const long long pairs = (p.T + 1) / 2;
    const size_t smem = fe::spectral_smem_bytes();
    if (!plan->spectral_smem_set.load()) {
        NAT_CUDA(cudaFuncSetAttribute(fe::spectral_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        plan->spectral_smem_set.store(true);
    }
    const int grid = static_cast<int>(std::min<long long>((pairs + 1) / 2, plan->sm_count * 4LL * 4));
    NAT_LAUNCH(7, static_cast<cudaStream_t>(stream), fe::spectral_stats_kernel<<<grid, fe::THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p));
    NAT_CUDA(cudaGetLastError());
    return NAT_OK;
}

}  // extern "C"
