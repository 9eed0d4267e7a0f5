// metrics_kernels.cu -- K4: fused error metrics over (prediction, ground-truth)
// rotation pairs: geodesic angle in degrees, ||I - Rp Rg^T||_F, Euler angles
// (pitch,yaw,roll), per-angle absolute error and 3-angle MAE, plus fp64 running
// sums for the eval means.  HBM-bound: 104 B of traffic per pair when every
// output is requested (36+36+12 in, 4+12+4 out).
//
// Replaces  src/agent.py:447-455          compute_err_deg_from_matrices
//           src/utils.py:232-260          compute_euler_angles_from_rotation_matrices
//                                         (a per-sample Python loop in the reference)
//           eval.py:76-98,125-133         per-angle errors, geodesic, Frobenius, means
//           pytorch3d so3_relative_angle  (restated; PARITY UNPINNED, see oracle/)
#include "kernels.cuh"
#include "so3_math.cuh"

namespace suhpe {

namespace {

constexpr int kMetThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void stage_in(float* dst, const float* __restrict__ src, int floats, bool vec) {
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < floats / 4; i += kMetThreads) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = threadIdx.x; i < floats; i += kMetThreads) dst[i] = __ldg(src + i);
    }
}
__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* src, int floats, bool vec) {
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < floats / 4; i += kMetThreads) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < floats; i += kMetThreads) dst[i] = src[i];
    }
}

__global__ void __launch_bounds__(kMetThreads)
metrics_kernel(MetricsArgs p, bool vec_ok) {
    __shared__ __align__(16) float sp[kMetThreads * 9];
    __shared__ __align__(16) float sg[kMetThreads * 9];
    __shared__ __align__(16) float se[kMetThreads * 3];   // gt euler in -> abs err out
    __shared__ __align__(16) float so[kMetThreads * 3];   // euler out
    __shared__ double red[6][kMetThreads / 32];

    const int t = threadIdx.x;
    const long long tiles = (p.n + kMetThreads - 1) / kMetThreads;
    const bool want_euler = p.euler || p.abs_err || p.mae || (p.sums && p.gt_euler);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    bool bad = false;

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long base = tile * kMetThreads;
        const int count = (int)min((long long)kMetThreads, p.n - base);
        const bool full = vec_ok && count == kMetThreads;
        stage_in(sp, p.Rp + base * 9, count * 9, full);
        if (p.Rg) stage_in(sg, p.Rg + base * 9, count * 9, full);
        if (p.gt_euler) stage_in(se, p.gt_euler + base * 3, count * 3, full);
        __syncthreads();
        if (t < count) {
            float Rp[9], Rg[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) Rp[k] = sp[t * 9 + k];
            const long long i = base + t;
            if (p.Rg) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Rg[k] = sg[t * 9 + k];
                if (p.geo_deg || p.sums) {
                    bool ok;
                    const float g = geodesic_degrees(relative_trace(Rp, Rg), &ok);
                    if (!ok) bad = true;
                    if (p.geo_deg) p.geo_deg[i] = g;
                    acc[0] += (double)g;
                }
                if (p.frob || p.sums) {
                    const float f = frobenius_to_identity(Rp, Rg);
                    if (p.frob) p.frob[i] = f;
                    acc[1] += (double)f;
                }
            }
            if (want_euler) {
                float e[3];
                euler_from_rotation(Rp, p.full_range != 0, e);
                so[t * 3] = e[0]; so[t * 3 + 1] = e[1]; so[t * 3 + 2] = e[2];
                if (p.gt_euler) {
                    float d[3], gt[3] = {se[t * 3], se[t * 3 + 1], se[t * 3 + 2]};
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float deg = div_rn(mul_rn(e[k], 180.0f), 3.14159265358979323846f);
                        d[k] = fabsf(deg - gt[k]);
                        acc[2 + k] += (double)d[k];
                    }
                    const float m = euler_mae_degrees(e, gt);
                    if (p.mae) p.mae[i] = m;
                    acc[5] += (double)m;
                    se[t * 3] = d[0]; se[t * 3 + 1] = d[1]; se[t * 3 + 2] = d[2];
                }
            }
        }
        __syncthreads();
        if (p.euler) stage_out(p.euler + base * 3, so, count * 3, full);
        if (p.abs_err && p.gt_euler) stage_out(p.abs_err + base * 3, se, count * 3, full);
        __syncthreads();
    }
    if (bad && p.status) atomicOr(p.status, kStatusTraceRange);
    if (p.sums) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = acc[k];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
            if ((t & 31) == 0) red[k][t >> 5] = v;
        }
        __syncthreads();
        if (t < 6) {
            double v = 0;
            for (int w = 0; w < kMetThreads / 32; ++w) v += red[t][w];
            atomicAdd(p.sums + t, v);
        }
    }
}

}  // namespace

cudaError_t launch_metrics(MetricsArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (p.n + kMetThreads - 1) / kMetThreads;
    long long blocks = tiles;
    const long long cap = (long long)sms * 6;
    if (blocks > cap) blocks = cap;
    auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec_ok = al(p.Rp) && al(p.Rg) && al(p.gt_euler) && al(p.euler) && al(p.abs_err);
    metrics_kernel<<<(unsigned)blocks, kMetThreads, 0, stream>>>(p, vec_ok);
    return cudaGetLastError();
}

}  // namespace suhpe
