// Time-base alignment before quantisation: F.interpolate(x, size=T_out, mode='linear',
// align_corners=False) on [B, C, T] (nat.py:3225-3236), with the floating-point steps of ATen's CPU kernel in this
// image (UpSampleKernel.cpp, compute_source_index_and_lambda + the two-tap interpolation, built with FMA contraction):
//   scale = float(T_in) / float(T_out)
//   real  = max(fma(scale, i + 0.5f, -0.5f), 0);  i0 = int(real);  i1 = i0 + (i0 < T_in - 1)
//   l1    = min(max(real - i0, 0), 1);  l0 = 1 - l1
//   out   = fma(l0, x[i0], l1 * x[i1])
// Bit-identical to torch 2.11 CPU on every geometry tried (tests/golden/interp_cases.npz, oracle/interp_oracle.py).
#pragma once

#include "nat_common.cuh"

namespace nat {
namespace interp {

__global__ void __launch_bounds__(256)
interp_linear_kernel(const float* __restrict__ x, long long rows, int t_in, int t_out, float scale,
                     float* __restrict__ out) {
    const long long total = rows * t_out;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long row = e / t_out;
        const int i = static_cast<int>(e - row * t_out);
        const float real = fmaxf(__fmaf_rn(scale, static_cast<float>(i) + 0.5f, -0.5f), 0.f);
        const int i0 = min(static_cast<int>(real), t_in - 1);
        const int i1 = i0 + (i0 < t_in - 1 ? 1 : 0);
        const float l1 = fminf(fmaxf(__fsub_rn(real, static_cast<float>(i0)), 0.f), 1.f);
        const float l0 = __fsub_rn(1.f, l1);
        const float* src = x + row * t_in;
        out[e] = __fmaf_rn(l0, __ldg(src + i0), __fmul_rn(l1, __ldg(src + i1)));
    }
}

}  // namespace interp
}  // namespace nat
