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
// Two decompositions of the same sum:
//   stream  thread per sample (large batches): 512-thread persistent CTAs; the grid sits in shared
//           memory as interleaved point PAIRS so one LDS.128 broadcast feeds two f32x2 operands and
//           every FMA of the loop is a packed FFMA2 over two grid points (29 per pair); block sums
//           fold into the totals every 128 points so the fp32 summation error does not grow with N
//   warp    warp per sample (small batches): lanes stride the grid, shuffle merge
// Per-sample set-up (proper SVD, T, the ground-truth term) runs in fp64: see laplace_setup.
#include "kernels.cuh"
#include "so3_math.cuh"

namespace suhpe {

namespace {

constexpr int kLapThreads = 256;
constexpr int kGridChunk = 4608;           // grid points resident in shared memory at once
constexpr unsigned kFull = 0xffffffffu;

// ---- packed fp32x2 helpers (Blackwell FFMA2) ------------------------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 dup(float x) { return pk(x, x); }

constexpr int kStreamThreads = 512;
constexpr int kStreamChunk = 6144;         // grid points per shared-memory chunk (216 KB), multiple of 4

// block sums of the packed loop: lo half = even grid points, hi half = odd
struct PackedSums { f2 z, c, m[9]; };

__device__ __forceinline__ void packed_scale(PackedSums& s, float sc) {
    const f2 k = dup(sc);
    s.z = mul2(s.z, k); s.c = mul2(s.c, k);
#pragma unroll
    for (int i = 0; i < 9; ++i) s.m[i] = mul2(s.m[i], k);
}

// two grid points (r[i] = (R_k[i], R_k+1[i])) in one straight-line FFMA2 body; same arithmetic per
// point as laplace_accum_point
__device__ __forceinline__ void packed_pair(LaplaceAccum& a, PackedSums& s, const float* A, float T, const f2* r) {
    f2 t = mul2(dup(A[0]), r[0]);
#pragma unroll
    for (int i = 1; i < 9; ++i) t = fma2(dup(A[i]), r[i], t);
    float d0, d1;
    upk(fma2(t, dup(-1.0f), dup(T)), d0, d1);
    const bool live0 = d0 >= kLapEps, live1 = d1 >= kLapEps;
    d0 = fmaxf(d0, kLapEps); d1 = fmaxf(d1, kLapEps);
    const f2 d = pk(d0, d1), rs = pk(mufu_rsqrt(d0), mufu_rsqrt(d1));
    f2 q = mul2(d, rs);
    q = fma2(mul2(rs, dup(0.5f)), fma2(mul2(q, dup(-1.0f)), q, d), q);
    float q0, q1;
    upk(q, q0, q1);
    const float qm = fminf(q0, q1);
    if (qm < a.qmin) {                        // new running maximum of p = -q: rescale (rare)
        const float sc = mufu_ex2((qm - a.qmin) * kLog2e);
        laplace_accum_scale(a, sc);
        packed_scale(s, sc);
        a.qmin = qm;
    }
    float e0, e1;
    upk(fma2(q, dup(-kLog2e), dup(a.qmin * kLog2e)), e0, e1);
    const f2 w = mul2(pk(mufu_ex2(e0), mufu_ex2(e1)), rs);    // exp(p - c) / (-p)
    s.z = add2(s.z, w);
    float c0, c1;
    upk(mul2(w, fma2(rs, rs, rs)), c0, c1);                    // w (1/q + 1/q^2)
    const f2 cw = pk(live0 ? c0 : 0.0f, live1 ? c1 : 0.0f);
    s.c = add2(s.c, cw);
#pragma unroll
    for (int i = 0; i < 9; ++i) s.m[i] = fma2(cw, r[i], s.m[i]);
}

__device__ __forceinline__ void packed_flush(LaplaceAccum& a, PackedSums& s) {
    float lo, hi;
    upk(s.z, lo, hi); a.z += lo + hi; s.z = pk(0.f, 0.f);
    upk(s.c, lo, hi); a.c += lo + hi; s.c = pk(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 9; ++i) { upk(s.m[i], lo, hi); a.m[i] += lo + hi; s.m[i] = pk(0.f, 0.f); }
    laplace_accum_flush(a);
}

__global__ void __launch_bounds__(kStreamThreads, 1)
laplace_stream_kernel(LaplaceArgs p, int chunk) {
    extern __shared__ __align__(16) float gp[];      // [chunk/2][9][2]: point pairs interleaved per component
    const long long tiles = (p.n + kStreamThreads - 1) / kStreamThreads;
    const bool single_chunk = p.N <= chunk;
    bool bad = false;

    auto load_chunk = [&](int c0, int cn) {
        const float* src = p.grid + (size_t)c0 * 9;
        for (int i = threadIdx.x; i < cn * 9; i += kStreamThreads) {
            const int k = i / 9, ij = i - 9 * k;
            gp[(k >> 1) * 18 + 2 * ij + (k & 1)] = __ldg(src + i);
        }
    };
    if (single_chunk) { load_chunk(0, p.N); __syncthreads(); }

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long sample = tile * kStreamThreads + threadIdx.x;
        const bool valid = sample < p.n;
        float A[9], Rs[9];
        double Td;
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = valid ? __ldg(p.A + sample * 9 + i) : ((i % 4 == 0) ? 1.f : 0.f);
        if (!laplace_setup(A, Rs, &Td) && valid) bad = true;
        const float T = (float)Td;

        LaplaceAccum a;
        laplace_accum_init(a);
        PackedSums s;
        s.z = s.c = pk(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 9; ++i) s.m[i] = pk(0.f, 0.f);

        for (int c0 = 0; c0 < p.N; c0 += chunk) {
            const int cn = min(chunk, p.N - c0);
            if (!single_chunk) { __syncthreads(); load_chunk(c0, cn); __syncthreads(); }
            const int groups = cn >> 2;                       // 4 points = 2 pairs = 9 float4
            const float4* g4 = reinterpret_cast<const float4*>(gp);
            for (int g0 = 0; g0 < groups; g0 += 32) {         // fold into the totals every 128 points
                const int g1 = min(g0 + 32, groups);
#pragma unroll 1
                for (int g = g0; g < g1; ++g) {
                    f2 r[18];
#pragma unroll
                    for (int v = 0; v < 9; ++v) {
                        const float4 x = g4[g * 9 + v];
                        r[2 * v] = pk(x.x, x.y); r[2 * v + 1] = pk(x.z, x.w);
                    }
                    packed_pair(a, s, A, T, r);
                    packed_pair(a, s, A, T, r + 9);
                }
                packed_flush(a, s);
            }
            for (int k = groups * 4; k < cn; ++k) {           // up to 3 trailing points
                float r[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) r[i] = gp[(k >> 1) * 18 + 2 * i + (k & 1)];
                laplace_accum_point(a, A, T, r);
            }
            laplace_accum_flush(a);
        }

        if (valid) {
            float Rg[9], grad[9], nll, logF;
#pragma unroll
            for (int i = 0; i < 9; ++i) Rg[i] = __ldg(p.Rgt + sample * 9 + i);
            laplace_finish(a, laplace_gt_gap(A, Rg, Td), Rs, Rg, p.N, &nll, &logF, grad);
            p.nll[sample] = nll;
            if (p.logF) p.logF[sample] = logF;
            if (p.mode) {
#pragma unroll
                for (int i = 0; i < 9; ++i) p.mode[sample * 9 + i] = Rs[i];
            }
            if (p.grad) {
#pragma unroll
                for (int i = 0; i < 9; ++i) p.grad[sample * 9 + i] = grad[i];
            }
        }
    }
    if (bad && p.status) atomicOr(p.status, kStatusNonFinite);
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
        float A[9], Rs[9];
        double Td;
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = valid ? __ldg(p.A + sample * 9 + i) : ((i % 4 == 0) ? 1.f : 0.f);
        if (!laplace_setup(A, Rs, &Td) && valid) bad = true;
        const float T = (float)Td;

        LaplaceAccum a;
        laplace_accum_init(a);

        for (int c0 = 0; c0 < p.N; c0 += chunk) {
            const int cn = min(chunk, p.N - c0);
            if (!single_chunk) { __syncthreads(); load_chunk(c0, cn); __syncthreads(); }
            if (L == 1) {
                // thread per sample: fold a block of 128 points into the totals at a time
                for (int k0 = 0; k0 < cn; k0 += 128) {
                    const int k1 = min(k0 + 128, cn);
                    for (int k = k0; k < k1; ++k) {
                        float r[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) r[i] = gs[i * stride + k];
                        laplace_accum_point(a, A, T, r);
                    }
                    laplace_accum_flush(a);
                }
            } else {
                for (int k = sub; k < cn; k += L) {
                    float r[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) r[i] = gs[i * stride + k];
                    laplace_accum_point(a, A, T, r);
                }
                laplace_accum_flush(a);
            }
        }

        if (L > 1) {
            // merge the lanes' partial sums under the global running minimum
            float qg = a.qmin;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) qg = fminf(qg, __shfl_xor_sync(kFull, qg, off));
            laplace_accum_rebase(a, qg);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                a.Z += __shfl_xor_sync(kFull, a.Z, off);
                a.C += __shfl_xor_sync(kFull, a.C, off);
#pragma unroll
                for (int i = 0; i < 9; ++i) a.M[i] += __shfl_xor_sync(kFull, a.M[i], off);
            }
        }

        if (valid && sub == 0) {
            float Rg[9], grad[9], nll, logF;
#pragma unroll
            for (int i = 0; i < 9; ++i) Rg[i] = __ldg(p.Rgt + sample * 9 + i);
            laplace_finish(a, laplace_gt_gap(A, Rg, Td), Rs, Rg, p.N, &nll, &logF, grad);
            p.nll[sample] = nll;
            if (p.logF) p.logF[sample] = logF;
            if (p.mode) {
#pragma unroll
                for (int i = 0; i < 9; ++i) p.mode[sample * 9 + i] = Rs[i];
            }
            if (p.grad) {
#pragma unroll
                for (int i = 0; i < 9; ++i) p.grad[sample * 9 + i] = grad[i];
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
    const bool per_thread = p.n >= (long long)sms * kLapThreads;
    cudaError_t err;
    if (per_thread) {
        const int chunk = p.N < kStreamChunk ? ((p.N + 3) & ~3) : kStreamChunk;
        const size_t smem = (size_t)chunk * 9 * sizeof(float);
        err = cudaFuncSetAttribute(laplace_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        const long long tiles = (p.n + kStreamThreads - 1) / kStreamThreads;
        const unsigned blocks = (unsigned)(tiles < sms ? tiles : sms);
        laplace_stream_kernel<<<blocks, kStreamThreads, smem, stream>>>(p, chunk);
    } else {
        const int chunk = p.N < kGridChunk ? p.N : kGridChunk;
        const int stride = (chunk + 3) & ~3;
        const size_t smem = (size_t)9 * stride * sizeof(float);
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
