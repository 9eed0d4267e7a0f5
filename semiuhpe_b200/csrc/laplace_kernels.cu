// laplace_kernels.cu -- K2L: rotation-Laplace NLL forward + backward in one pass.
//
// Replaces src/laplace/rotation_laplace.py:24-34,58-72,76-115,140-173
//   p(A,X)  = -sqrt(max(T - <A,X>, 1e-8)),  T = s1+s2+s3*sign(det A)
//   logF    = c + log( sum_k exp(p_k - c)/(-p_k) / N ),  c = max_k p_k
//   nll     = logF - p(A,R) + log(-p(A,R))
// and its autograd backward (closed form):
//   d nll/dA = -sum_k chat_k (R* - R_k) + c_x (R* - R_gt),   R* = U diag(1,1,sgn) V^T
// The reference materialises a (b,N,3,3) tensor (166 KB per sample for the
// 4608-point grid) plus autograd copies; here the grid is transposed once into
// shared memory ([9][N] floats, 166 KB for N=4608) and every sample streams it
// with an online max-shifted sum: running min of q, Z = sum w, C = sum chat,
// M = sum chat R_k (9) -- 12 accumulators, nothing written per (sample, grid) pair.
// FP32 FMA only: the reference forbids TF32 here (rotation_laplace.py:13), and a
// K=9 contraction has no tensor-core shape anyway.
//
// Two decompositions of the same loop (template L = lanes per sample):
//   L = 1   thread per sample, grid values are smem broadcasts   (large batches)
//   L = 32  warp per sample, lanes stride the grid, shuffle merge (small batches)
#include "kernels.cuh"
#include "so3_math.cuh"

namespace suhpe {

namespace {

constexpr int kLapThreads = 256;
constexpr int kGridChunk = 4608;           // grid points resident in shared memory at once
constexpr unsigned kFull = 0xffffffffu;
constexpr float kLapEps = 1e-8f;           // rotation_laplace.py:11

// sqrt with one Newton step on top of MUFU.RSQ: returns q ~ sqrt(d) and rs ~ 1/sqrt(d)
__device__ __forceinline__ void sqrt_pair(float d, float& q, float& rs) {
    rs = mufu_rsqrt(d);
    q = d * rs;
    const float err = fmaf(-q, q, d);
    q = fmaf(0.5f * rs, err, q);
}

struct Accum {
    float qmin, Z, C, M[9];
};

__device__ __forceinline__ void accum_point(Accum& a, const float* A, float T, const float* r) {
    float t = A[0] * r[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) t = fmaf(A[i], r[i], t);
    const float d = T - t;
    const bool live = d >= kLapEps;          // clamp_min passes the gradient where input >= min
    float q, rs;
    sqrt_pair(fmaxf(d, kLapEps), q, rs);
    if (q < a.qmin) {                         // new running maximum of p = -q: rescale (rare)
        const float sc = mufu_ex2((q - a.qmin) * kLog2e);
        a.Z *= sc; a.C *= sc;
#pragma unroll
        for (int i = 0; i < 9; ++i) a.M[i] *= sc;
        a.qmin = q;
    }
    const float w = mufu_ex2((a.qmin - q) * kLog2e) * rs;    // exp(p - c) / (-p)
    a.Z += w;
    const float cw = live ? w * fmaf(rs, rs, rs) : 0.0f;      // w (1/q + 1/q^2); the 1/2 is applied once at the end
    a.C += cw;
#pragma unroll
    for (int i = 0; i < 9; ++i) a.M[i] = fmaf(cw, r[i], a.M[i]);
}

template <int L>
__global__ void __launch_bounds__(kLapThreads)
laplace_kernel(LaplaceArgs p, int chunk, int stride) {
    extern __shared__ __align__(16) float gs[];      // [9][stride]
    constexpr int kSamplesPerBlock = kLapThreads / L;
    const int sub = threadIdx.x % L;
    const int slot = threadIdx.x / L;
    const long long tiles = (p.n + kSamplesPerBlock - 1) / kSamplesPerBlock;
    const bool single_chunk = p.N <= chunk;
    bool bad = false;

    auto load_chunk = [&](int c0, int cn) {
        // global (cn,9) row-major -> shared [9][stride]
        const float* src = p.grid + (size_t)c0 * 9;
        for (int i = threadIdx.x; i < cn * 9; i += kLapThreads) {
            const int k = i / 9, ij = i - 9 * k;
            gs[ij * stride + k] = __ldg(src + i);
        }
    };
    if (single_chunk) { load_chunk(0, p.N); __syncthreads(); }

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long sample = tile * kSamplesPerBlock + slot;
        const bool valid = sample < p.n;
        float A[9], U[9], V[9], s[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = valid ? __ldg(p.A + sample * 9 + i) : ((i % 4 == 0) ? 1.f : 0.f);
        if (!proper_svd3(A, U, V, s) && valid) bad = true;
        const float T = s[0] + s[1] + s[2];

        Accum a;
        a.qmin = INFINITY; a.Z = 0.f; a.C = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) a.M[i] = 0.f;

        for (int c0 = 0; c0 < p.N; c0 += chunk) {
            const int cn = min(chunk, p.N - c0);
            if (!single_chunk) { __syncthreads(); load_chunk(c0, cn); __syncthreads(); }
            for (int k = sub; k < cn; k += L) {
                float r[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) r[i] = gs[i * stride + k];
                accum_point(a, A, T, r);
            }
        }

        if (L > 1) {
            // merge the lanes' partial sums under the global running minimum
            float qg = a.qmin;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) qg = fminf(qg, __shfl_xor_sync(kFull, qg, off));
            const float sc = (a.qmin == INFINITY) ? 0.f : mufu_ex2((qg - a.qmin) * kLog2e);
            a.Z *= sc; a.C *= sc;
#pragma unroll
            for (int i = 0; i < 9; ++i) a.M[i] *= sc;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                a.Z += __shfl_xor_sync(kFull, a.Z, off);
                a.C += __shfl_xor_sync(kFull, a.C, off);
#pragma unroll
                for (int i = 0; i < 9; ++i) a.M[i] += __shfl_xor_sync(kFull, a.M[i], off);
            }
            a.qmin = qg;
        }

        if (valid && sub == 0) {
            float Rs[9];
            u_diag_vt(U, V, 1.f, 1.f, 1.f, Rs);
            // logF = c + log(sum * (1/N)), c = -qmin   (rotation_laplace.py:69-71)
            const float logF = -a.qmin + logf(a.Z * (1.0f / (float)p.N));
            float tx = 0.f;
            float Rg[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) { Rg[i] = __ldg(p.Rgt + sample * 9 + i); tx = fmaf(A[i], Rg[i], tx); }
            const float dx = T - tx;
            const float qx = sqrt_rn(fmaxf(dx, kLapEps));
            p.nll[sample] = logF + qx + logf(qx);
            if (p.logF) p.logF[sample] = logF;
            if (p.mode) {
#pragma unroll
                for (int i = 0; i < 9; ++i) p.mode[sample * 9 + i] = Rs[i];
            }
            if (p.grad) {
                const float invZ = 0.5f / a.Z;
                const float cs = a.C * invZ;                                   // sum_k chat_k
                const float cx = (dx >= kLapEps) ? 0.5f * (1.0f + 1.0f / qx) / qx : 0.0f;
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    p.grad[sample * 9 + i] = fmaf(a.M[i], invZ, fmaf(cx - cs, Rs[i], -cx * Rg[i]));
            }
        }
    }
    if (bad && p.status) atomicOr(p.status, kStatusNonFinite);
}

}  // namespace

cudaError_t launch_laplace(LaplaceArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    if (p.N <= 0) return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int chunk = p.N < kGridChunk ? p.N : kGridChunk;
    const int stride = (chunk + 3) & ~3;
    const size_t smem = (size_t)9 * stride * sizeof(float);
    const bool per_thread = p.n >= (long long)sms * kLapThreads;
    cudaError_t err;
    if (per_thread) {
        err = cudaFuncSetAttribute(laplace_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        const long long tiles = (p.n + kLapThreads - 1) / kLapThreads;
        const unsigned blocks = (unsigned)(tiles < sms ? tiles : sms);
        laplace_kernel<1><<<blocks, kLapThreads, smem, stream>>>(p, chunk, stride);
    } else {
        err = cudaFuncSetAttribute(laplace_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        // spread the samples over all SMs: up to 8 per block pass, as few as 1 when the batch is tiny
        const long long tiles = (p.n + (kLapThreads / 32) - 1) / (kLapThreads / 32);
        const unsigned blocks = (unsigned)(tiles < sms ? tiles : sms);
        laplace_kernel<32><<<blocks, kLapThreads, smem, stream>>>(p, chunk, stride);
    }
    return cudaGetLastError();
}

}  // namespace suhpe
