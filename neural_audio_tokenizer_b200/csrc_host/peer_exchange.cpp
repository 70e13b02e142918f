// All-gather of the index streams over NVLink peer memory by the copy engines. Host code only, its own translation
// unit (csrc/ holds the kernel sources and their launcher: the sources bench.build_id ties the ncu captures to).
//
// One process per GPU. Every rank owns three output buffers [rows, world * cols] (step s uses buffer s % 3) and one step
// counter per sender, all plain device memory exported through CUDA IPC. A step, on the caller's stream:
//   1. one 2-D peer copy per rank writes this rank's [rows, cols] block straight into its column range of every
//      rank's output buffer (the final layout: no staging on the receiver, no un-pad or permute kernel);
//   2. the step number is copied into this rank's counter on every rank (same stream: after the data);
//   3. the stream waits (cuStreamWaitValue32, flushing remote writes) until every sender's counter has reached the
//      step.
// No kernel runs: NCCL's all-gather kernel has to squeeze in beside a persistent kernel that owns every SM, these
// copies do not touch an SM. A sender can start step s + 2 as soon as every rank has SENT step s + 1, which a receiver
// that consumes its results one step late (the overlapped use) does before it has read step s: with two buffers the
// remote writes of s + 2 would land on the buffer it is still reading. With three, buffer s % 3 is written again by
// step s + 3, which needs every rank's step s + 2 call; a result is therefore valid until the call after next.
// Teardown is two-phase: every rank unmaps its peers' buffers (nat_peer_disconnect), the ranks meet at a barrier of
// the caller's, and only then does anyone free the memory the others had mapped (nat_peer_destroy): freeing an
// exported allocation that an importer still maps is undefined.
#include "../../include/nat_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace {

thread_local std::string g_peer_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_peer_error = buf;
    return code;
}

#define NAT_CUDA(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(NAT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

}  // namespace

const char* nat_peer_last_error(void) { return g_peer_error.c_str(); }

constexpr unsigned kPeerBuffers = 3;

struct nat_peer_ctx {
    int world = 0, rank = 0, device = 0;
    size_t rows = 0, col_bytes = 0, pitch = 0;      // block: rows x col_bytes; output row pitch = world * col_bytes
    char* out = nullptr;                            // [kPeerBuffers][rows][pitch]
    uint32_t* flags = nullptr;                      // [world] step counters written by the senders
    uint32_t* step_word = nullptr;                  // the value this rank sends
    std::vector<char*> peer_out;
    std::vector<uint32_t*> peer_flags;
    bool connected = false;
    bool can_flush = false;
    unsigned step = 0;
};

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");

namespace {
// The library does not link libcuda (it must load on a box without a driver): driver entry points come from the runtime.
template <class Fn>
Fn peer_driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<Fn>(p);
}
using DeviceGetAttributeFn = CUresult (*)(int*, CUdevice_attribute, CUdevice);
using MemsetD32AsyncFn = CUresult (*)(CUdeviceptr, unsigned int, size_t, CUstream);
using StreamWaitValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
}  // namespace

int nat_peer_create(int world, int rank, size_t rows, size_t col_bytes, nat_peer_ctx** out) {
    if (out == nullptr || world < 1 || world > 64 || rank < 0 || rank >= world || rows == 0 || col_bytes == 0)
        return fail(NAT_ERR_INVALID_ARGUMENT, "bad peer exchange geometry");
    auto* ctx = new nat_peer_ctx;
    ctx->world = world; ctx->rank = rank; ctx->rows = rows; ctx->col_bytes = col_bytes; ctx->pitch = col_bytes * world;
    ctx->peer_out.assign(world, nullptr); ctx->peer_flags.assign(world, nullptr);
    auto cleanup = [&](int rc) { cudaFree(ctx->out); cudaFree(ctx->flags); cudaFree(ctx->step_word); delete ctx; return rc; };
    if (cudaGetDevice(&ctx->device) != cudaSuccess) return cleanup(fail(NAT_ERR_CUDA, "cudaGetDevice failed"));
    const size_t bytes = kPeerBuffers * rows * ctx->pitch;
    // An IPC handle exports the whole allocation block a pointer lives in: both exported buffers get blocks of their
    // own (sizes rounded up to 2 MiB), so no unrelated small allocation of this process shares a block with them.
    constexpr size_t kIpcGranule = size_t(2) << 20;
    const size_t out_alloc = (bytes + kIpcGranule - 1) / kIpcGranule * kIpcGranule;
    if (cudaMalloc(&ctx->out, out_alloc) != cudaSuccess || cudaMalloc(&ctx->flags, kIpcGranule) != cudaSuccess ||
        cudaMalloc(&ctx->step_word, sizeof(uint32_t)) != cudaSuccess)
        return cleanup(fail(NAT_ERR_CUDA, "cudaMalloc of the exchange buffers failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaMemset(ctx->out, 0, bytes) != cudaSuccess || cudaMemset(ctx->flags, 0, sizeof(uint32_t) * world) != cudaSuccess)
        return cleanup(fail(NAT_ERR_CUDA, "cudaMemset failed"));
    int flush = 0;
    static const auto get_attr = peer_driver_fn<DeviceGetAttributeFn>("cuDeviceGetAttribute");
    if (get_attr != nullptr) get_attr(&flush, CU_DEVICE_ATTRIBUTE_CAN_FLUSH_REMOTE_WRITES, static_cast<CUdevice>(ctx->device));
    ctx->can_flush = flush != 0;
    ctx->peer_out[rank] = ctx->out; ctx->peer_flags[rank] = ctx->flags;
    ctx->connected = world == 1;
    *out = ctx;
    return NAT_OK;
}

int nat_peer_export(const nat_peer_ctx* ctx, void* handles_out) {
    if (ctx == nullptr || handles_out == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    auto* h = static_cast<cudaIpcMemHandle_t*>(handles_out);
    NAT_CUDA(cudaIpcGetMemHandle(&h[0], ctx->out));
    NAT_CUDA(cudaIpcGetMemHandle(&h[1], ctx->flags));
    return NAT_OK;
}

int nat_peer_connect(nat_peer_ctx* ctx, const void* handles_all) {
    if (ctx == nullptr || handles_all == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    const auto* h = static_cast<const cudaIpcMemHandle_t*>(handles_all);
    for (int p = 0; p < ctx->world; ++p) {
        if (p == ctx->rank || ctx->peer_out[p] != nullptr) continue;
        void* a = nullptr; void* b = nullptr;
        NAT_CUDA(cudaIpcOpenMemHandle(&a, h[2 * p], cudaIpcMemLazyEnablePeerAccess));
        NAT_CUDA(cudaIpcOpenMemHandle(&b, h[2 * p + 1], cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_out[p] = static_cast<char*>(a);
        ctx->peer_flags[p] = static_cast<uint32_t*>(b);
    }
    ctx->connected = true;
    return NAT_OK;
}

int nat_peer_all_gather(nat_peer_ctx* ctx, const void* block_dev, size_t block_pitch, void* stream, void** gathered_out) {
    if (ctx == nullptr || block_dev == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    if (!ctx->connected) return fail(NAT_ERR_INVALID_ARGUMENT, "nat_peer_connect has not been called");
    if (block_pitch < ctx->col_bytes) return fail(NAT_ERR_INVALID_ARGUMENT, "block pitch smaller than a row of the block");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned s = ++ctx->step;
    const size_t buf = static_cast<size_t>(s % kPeerBuffers) * ctx->rows * ctx->pitch;
    for (int i = 0; i < ctx->world; ++i) {
        const int p = (ctx->rank + i) % ctx->world;                 // every rank starts with a different receiver
        NAT_CUDA(cudaMemcpy2DAsync(ctx->peer_out[p] + buf + ctx->rank * ctx->col_bytes, ctx->pitch, block_dev, block_pitch,
                                   ctx->col_bytes, ctx->rows, cudaMemcpyDeviceToDevice, st));
    }
    static const auto memset_d32 = peer_driver_fn<MemsetD32AsyncFn>("cuMemsetD32Async");
    static const auto wait_value = peer_driver_fn<StreamWaitValue32Fn>("cuStreamWaitValue32");
    if (memset_d32 == nullptr || wait_value == nullptr)
        return fail(NAT_ERR_CUDA, "the driver does not export cuMemsetD32Async / cuStreamWaitValue32");
    if (memset_d32(reinterpret_cast<CUdeviceptr>(ctx->step_word), s, 1, st) != CUDA_SUCCESS)
        return fail(NAT_ERR_CUDA, "cuMemsetD32Async failed");
    for (int i = 0; i < ctx->world; ++i) {
        const int p = (ctx->rank + i) % ctx->world;
        NAT_CUDA(cudaMemcpyAsync(ctx->peer_flags[p] + ctx->rank, ctx->step_word, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    const unsigned flags = CU_STREAM_WAIT_VALUE_GEQ | (ctx->can_flush ? CU_STREAM_WAIT_VALUE_FLUSH : 0u);
    for (int q = 0; q < ctx->world; ++q)
        if (wait_value(st, reinterpret_cast<CUdeviceptr>(ctx->flags + q), s, flags) != CUDA_SUCCESS)
            return fail(NAT_ERR_CUDA, "cuStreamWaitValue32 failed (stream memory operations unavailable?)");
    if (gathered_out != nullptr) *gathered_out = ctx->out + buf;
    return NAT_OK;
}

void* nat_peer_buffer(const nat_peer_ctx* ctx, int which) {
    return ctx == nullptr ? nullptr : ctx->out + static_cast<size_t>(static_cast<unsigned>(which) % kPeerBuffers) * ctx->rows * ctx->pitch;
}

int nat_peer_disconnect(nat_peer_ctx* ctx) {
    if (ctx == nullptr) return fail(NAT_ERR_INVALID_ARGUMENT, "null argument");
    cudaError_t first = cudaSuccess;
    for (int p = 0; p < ctx->world; ++p) {
        if (p == ctx->rank) continue;
        if (ctx->peer_out[p] != nullptr) { const cudaError_t e = cudaIpcCloseMemHandle(ctx->peer_out[p]); if (first == cudaSuccess) first = e; }
        if (ctx->peer_flags[p] != nullptr) { const cudaError_t e = cudaIpcCloseMemHandle(ctx->peer_flags[p]); if (first == cudaSuccess) first = e; }
        ctx->peer_out[p] = nullptr;
        ctx->peer_flags[p] = nullptr;
    }
    ctx->connected = ctx->world == 1;
    if (first != cudaSuccess) return fail(NAT_ERR_CUDA, "cudaIpcCloseMemHandle failed: %s", cudaGetErrorString(first));
    return NAT_OK;
}

void nat_peer_destroy(nat_peer_ctx* ctx) {
    if (ctx == nullptr) return;
    nat_peer_disconnect(ctx);            // a no-op after an explicit disconnect
    cudaFree(ctx->out); cudaFree(ctx->flags); cudaFree(ctx->step_word);
    delete ctx;
}
