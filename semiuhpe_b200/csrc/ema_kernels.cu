// ema_kernels.cu -- EMA / EMAN teacher update: one launch per group of up to 48
// parameter tensors instead of two torch ops per tensor.
//
// Replaces SSLAgent.update_ema_variables (src/agent.py:277-299), both branches:
//   mode 0 (config.eman, state_dict form, :293)   ema = fl(ema * alpha) + fl((1 - alpha) * src)
//   mode 1 (parameters form, :298)                ema.mul_(alpha).add_(src, alpha = 1 - alpha)
//                                                 = fma(1 - alpha, src, fl(ema * alpha))   (ATen's add is a*b+c fused)
// alpha and 1 - alpha arrive already rounded to fp32 (what ATen does with a Python scalar against an
// fp32 tensor).  HBM-bound: 12 B per element (two reads, one write).
//
// Multi-tensor layout: the tensor table (pointers, sizes) and the CTA -> (tensor, chunk) map travel
// in the kernel's parameter block (2.6 KB), so nothing is staged through device memory and the
// call needs no host synchronisation; the host splits long lists over several launches.
#include "kernels.cuh"

namespace suhpe {

namespace {
constexpr int kEmaThreads = 256;

__global__ void __launch_bounds__(kEmaThreads)
ema_update_kernel(EmaLaunch t, float alpha, float one_minus_alpha, int mode) {
    const int ti = t.block_tensor[blockIdx.x];
    float* __restrict__ e = t.ema[ti];
    const float* __restrict__ s = t.src[ti];
    const long long begin = (long long)t.block_chunk[blockIdx.x] * kEmaChunk;
    const long long n = t.numel[ti];
    const long long end = begin + kEmaChunk < n ? begin + kEmaChunk : n;
    auto blend = [&](float ev, float sv) {
        const float a = __fmul_rn(ev, alpha);
        return mode == 0 ? __fadd_rn(a, __fmul_rn(one_minus_alpha, sv)) : __fmaf_rn(one_minus_alpha, sv, a);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(s)) & 15u) == 0;   // begin is a multiple of 4
    if (vec) {
        const long long v0 = begin >> 2, v1 = end >> 2;
        float4* e4 = reinterpret_cast<float4*>(e);
        const float4* s4 = reinterpret_cast<const float4*>(s);
        for (long long i = v0 + threadIdx.x; i < v1; i += kEmaThreads) {
            float4 a = e4[i];
            const float4 b = __ldg(s4 + i);
            a.x = blend(a.x, b.x); a.y = blend(a.y, b.y); a.z = blend(a.z, b.z); a.w = blend(a.w, b.w);
            e4[i] = a;
        }
        for (long long i = (v1 << 2) + threadIdx.x; i < end; i += kEmaThreads) e[i] = blend(e[i], __ldg(s + i));
    } else {
        for (long long i = begin + threadIdx.x; i < end; i += kEmaThreads) e[i] = blend(e[i], __ldg(s + i));
    }
}
}  // namespace

cudaError_t launch_ema_update(float* const* ema, const float* const* src, const long long* numel, int count,
                              float alpha, float one_minus_alpha, int mode, cudaStream_t stream) {
    EmaLaunch t;
    int nt = 0, nb = 0;
    auto flush = [&]() -> cudaError_t {
        if (nb == 0) { nt = 0; return cudaSuccess; }
        ema_update_kernel<<<nb, kEmaThreads, 0, stream>>>(t, alpha, one_minus_alpha, mode);
        nt = 0; nb = 0;
        return cudaGetLastError();
    };
    for (int i = 0; i < count; ++i) {
        if (numel[i] <= 0) continue;
        const long long chunks = (numel[i] + kEmaChunk - 1) / kEmaChunk;
        long long c = 0;
        while (c < chunks) {
            if (nt == kEmaMaxTensors || nb == kEmaMaxBlocks) {
                const cudaError_t err = flush();
                if (err != cudaSuccess) return err;
            }
            // (re-)enter tensor i in the current launch's table
            t.ema[nt] = ema[i]; t.src[nt] = src[i]; t.numel[nt] = numel[i];
            while (c < chunks && nb < kEmaMaxBlocks) {
                t.block_tensor[nb] = (unsigned char)nt;
                t.block_chunk[nb] = (int)c;
                ++nb; ++c;
            }
            ++nt;
        }
    }
    return flush();
}

}  // namespace suhpe
