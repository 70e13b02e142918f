// Ad-hoc microbenchmark (not a test, not part of the library): L2 -> SM read bandwidth of the whole chip on a buffer
// that fits L2, the denominator behind DESIGN.md's statement that the stack kernel runs at the fabric's limit.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_bandwidth tests/host/l2_bandwidth.cu && ./l2_bandwidth
#include <cstdio>
#include <cuda_runtime.h>

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldcg256(const float* p) {
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p) : "memory");
    return r;
}

template <int UNROLL>
__global__ void __launch_bounds__(512) read_kernel(const float* __restrict__ buf, size_t n_f8, int passes, float* out) {
    float acc = 0.f;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t first = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t full = n_f8 / (UNROLL * stride) * (UNROLL * stride);      // every thread reads the same count
    for (int p = 0; p < passes; ++p) {
        for (size_t i = first; i < full; i += UNROLL * stride) {
            F8 r[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) r[u] = ldcg256(buf + (i + u * stride) * 8);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc += r[u].v[0] + r[u].v[7];
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (size_t mb : {16, 32, 48, 64, 96}) {
        const size_t bytes = mb << 20, n_f8 = bytes / 32;
        float *buf, *out;
        cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
        cudaMemset(buf, 0, bytes);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int ctas : {1, 2, 4}) {
            const int passes = 200;
            read_kernel<4><<<sms * ctas, 512>>>(buf, n_f8, 5, out);          // warm: the buffer is in L2
            cudaEventRecord(e0);
            read_kernel<4><<<sms * ctas, 512>>>(buf, n_f8, passes, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const size_t stride = (size_t)sms * ctas * 512, full = n_f8 / (4 * stride) * (4 * stride);
            printf("buffer %3zu MB (%.1f MB read per pass), %d CTAs x 512 threads per SM: %.2f TB/s\n", mb, full * 32 / 1048576.0, ctas,
                   full * 32.0 * passes / (ms * 1e-3) / 1e12);
        }
        cudaFree(buf); cudaFree(out);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
