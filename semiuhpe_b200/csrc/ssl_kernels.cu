// ssl_kernels.cu -- the small kernels around K2 that let a whole semi-supervised training step's loss
// head run as one fixed launch sequence (suhpe_ssl_step_f32, capi.cu):
//   scale_rows_kernel     per-sample gradient x incoming d L / d loss_i (the backward of the loss mirrors)
//   ssl_finalize_kernel   losses.mean() (src/agent.py:83), masked unsupervised mean x mask ratio (:163-166),
//                         loss_all = loss + SSL_lambda * unsuper_loss (:203) and the gradients of loss_all
// Both are launch-latency sized (a training batch is 32 + 128 rotations); nothing here is a hot loop.
#include "kernels.cuh"

namespace suhpe {

namespace {

__global__ void __launch_bounds__(256)
scale_rows_kernel(const float* __restrict__ in, long long n, int width, const float* __restrict__ row_weight,
                  const float* __restrict__ scalar_weight, const uint8_t* __restrict__ keep, float* __restrict__ out) {
    const float sw = scalar_weight ? *scalar_weight : 1.0f;
    const long long total = n * width;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long row = i / width;
        float v = 0.0f;                                   // a filtered row is written as zeros: 0 * NaN never forms
        if (keep == nullptr || keep[row] != 0) {
            v = in[i];
            if (row_weight) v *= row_weight[row];
            if (scalar_weight) v *= sw;
        }
        out[i] = v;
    }
}

constexpr int kFinThreads = 256;

// deterministic block sum (fixed tree): every thread returns the total
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kFinThreads / 32; ++w) t += red[w];
    return t;
}

// one CTA: the batch of a training step is a few hundred rows
__global__ void __launch_bounds__(kFinThreads)
ssl_finalize_kernel(SslFinalizeArgs a) {
    __shared__ float red[kFinThreads / 32];
    const int t = threadIdx.x;
    float s_l = 0.0f, s_u = 0.0f;
    for (long long i = t; i < a.b_l; i += kFinThreads) s_l += a.nll_l[i];
    for (long long i = t; i < a.b_u; i += kFinThreads) {
        const bool m = a.mask[i] != 0;
        const float l = m ? a.loss_u[i] : 0.0f;
        s_u += l;
        if (a.losses_u_out) a.losses_u_out[i] = l;
    }
    s_l = block_sum(s_l, red);
    s_u = block_sum(s_u, red);
    const float inv_l = a.b_l > 0 ? 1.0f / (float)a.b_l : 0.0f;
    const float inv_u = a.b_u > 0 ? 1.0f / (float)a.b_u : 0.0f;
    if (t == 0) {
        const float loss_sup = s_l * inv_l;
        const float unsup = s_u * inv_u;                  // == mean(l[mask]) * mask_ratio (src/agent.py:163,166)
        a.losses[0] = loss_sup;
        a.losses[1] = unsup;
        a.losses[2] = (a.b_u > 0 && a.kept) ? (float)(*a.kept) * inv_u : 0.0f;
        a.losses[3] = fmaf(a.ssl_lambda, unsup, loss_sup);
    }
    if (a.grad_l)
        for (long long i = t; i < a.b_l * 9; i += kFinThreads) a.grad_l[i] *= inv_l;
    if (a.grad_u) {
        const float w = a.ssl_lambda * inv_u;
        for (long long i = t; i < a.b_u * 9; i += kFinThreads)
            a.grad_u[i] = a.mask[i / 9] ? a.grad_u[i] * w : 0.0f;
    }
}

}  // namespace

cudaError_t launch_scale_rows(const float* in, long long n, int width, const float* row_weight, const float* scalar_weight,
                              const uint8_t* keep, float* out, cudaStream_t stream) {
    if (n <= 0 || width <= 0) return cudaSuccess;
    const long long total = n * width;
    long long blocks = (total + 256 * 4 - 1) / (256 * 4);
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    scale_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, n, width, row_weight, scalar_weight, keep, out);
    return cudaGetLastError();
}

cudaError_t launch_ssl_finalize(SslFinalizeArgs a, cudaStream_t stream) {
    ssl_finalize_kernel<<<1, kFinThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace suhpe
