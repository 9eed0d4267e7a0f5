// fisher_kernels.cu -- K1 (batched proper 3x3 SVD -> rotation) and K2 (fused
// matrix-Fisher NLL + gradient + entropy [+ first radix histogram of the
// entropy keys]) for sm_100a.
//
// Replaces, for the whole batch in ONE launch and without leaving the GPU:
//   src/fisher/fisher_utils.py:14-48   vmf_loss / KL_Fisher / batch_torch_A_to_R
//   src/fisher/fisher_utils.py:51-81   fisher_log_pdf / fisher_entropy
//   src/fisher/torch_norm_factor.py:66-92  logC_F forward AND backward
// (the reference runs ~1,980 + ~4,600 ATen ops and CPU LAPACK SVDs for these).
//
// Work decomposition (no tensor cores: nothing here is a dense contraction)
//   phase 1  thread-per-sample : coalesced float4 tile load -> smem -> 9 regs,
//            Hestenes SVD in registers, U/V parked in smem
//   phase 2  warp-per-sample   : the 3 x 512 quadrature nodes of one sample are
//            spread over the 32 lanes (16 iterations x 3 integrand families);
//            consecutive nodes sit in consecutive lanes so the |a| <= 3.75
//            polynomial switch is warp-uniform except in the <= 2 iterations
//            that straddle a crossover; one MUFU.EX2 per node (the small-branch
//            exp(-a) factors are merged into the tail exponential)
//   phase 3  thread-per-sample : closing arithmetic, gradient
//            -R_gt + overreg * U diag(g) V^T, entropy, staged float4 stores
#include "kernels.cuh"
#include "so3_math.cuh"

namespace suhpe {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;

// per-warp shared scratch (floats)
constexpr int kTileFloats = 32 * 9;              // one 32-sample tile of 3x3 records
struct __align__(16) WarpScratch {
    float a[kTileFloats];      // A in  -> gradient out
    float r[kTileFloats];      // R_gt in -> projected rotation out
    float uv[18 * 32];         // U,V parked during the quadrature, [k][lane]
};

// coalesced load of `count` 3x3 records starting at `src` into smem `dst`
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int count, bool vec, int lane) {
    if (vec && count == 32) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = lane + 32 * k;
            if (i < kTileFloats / 4) d4[i] = __ldg(s4 + i);
        }
    } else {
        for (int i = lane; i < count * 9; i += 32) dst[i] = __ldg(src + i);
    }
}
__device__ __forceinline__ void store_tile(float* __restrict__ dst, const float* src, int count, bool vec, int lane) {
    if (vec && count == 32) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = lane + 32 * k;
            if (i < kTileFloats / 4) d4[i] = s4[i];
        }
    } else {
        for (int i = lane; i < count * 9; i += 32) dst[i] = src[i];
    }
}

// One Bessel factor with a warp-uniform branch whenever the whole warp sits on
// one side of the 3.75 switch.  `e` accumulates the merged exponent.
__device__ __forceinline__ float bessel_factor(float a, float& e) {
    const bool small = (a <= kBesselSwitch);
    if (__all_sync(kFull, small)) { e -= a; return i0_small_poly(a); }
    if (!__any_sync(kFull, small)) { return i0e_large(a); }
    if (small) { e -= a; return i0_small_poly(a); }
    return i0e_large(a);
}

__device__ __forceinline__ float family_node(const Family& f, float u, float v) {
    float e = -f.c * u;
    const float pd = bessel_factor(fabsf(f.fd * u), e);
    const float ps = bessel_factor(fabsf(f.fs * v), e);
    return pd * ps * mufu_ex2(e * kLog2e);
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
    return x;
}

}  // namespace

// ---------------------------------------------------------------------------
// K2: fused Fisher kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fisher_fused_kernel(FisherArgs p) {
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    __shared__ unsigned int hist_s[kHistBins1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const bool want_hist = (p.hist != nullptr);
    if (want_hist) {
        for (int i = threadIdx.x; i < kHistBins1; i += kThreads) hist_s[i] = 0u;
        __syncthreads();
    }

    const long long warps_total = (long long)gridDim.x * kWarpsPerBlock;
    const long long gwarp = (long long)blockIdx.x * kWarpsPerBlock + warp;
    const int spw = p.samples_per_warp;
    const long long tiles = (p.n + spw - 1) / spw;
    bool bad = false;

    for (long long tile = gwarp; tile < tiles; tile += warps_total) {
        const long long base = tile * spw;
        const int count = (int)min((long long)spw, p.n - base);
        const bool mine = lane < count;

        // ---- phase 1: load + SVD (thread per sample) -------------------------
        float s[3] = {0.f, 0.f, 0.f};
        float dot = 0.f;
        if (p.Sin) {
            // logC_F entry: singular values given directly (torch_norm_factor.py:92), U = V = I
            if (mine) {
                const long long i = base + lane;
                s[0] = __ldg(p.Sin + 3 * i); s[1] = __ldg(p.Sin + 3 * i + 1); s[2] = __ldg(p.Sin + 3 * i + 2);
            }
        } else {
            load_tile(ws.a, p.A + base * 9, count, p.vec_ok, lane);
            if (p.Rgt) load_tile(ws.r, p.Rgt + base * 9, count, p.vec_ok, lane);
            __syncwarp();
            if (mine) {
                float A[9], U[9], V[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) A[k] = ws.a[lane * 9 + k];
                if (p.Rgt) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) dot = fmaf(A[k], ws.r[lane * 9 + k], dot);
                }
                if (!proper_svd3(A, U, V, s)) bad = true;
#pragma unroll
                for (int k = 0; k < 9; ++k) { ws.uv[k * 32 + lane] = U[k]; ws.uv[(9 + k) * 32 + lane] = V[k]; }
            }
            __syncwarp();
        }

        // ---- phase 2: quadrature (warp per sample) ---------------------------
        float F = 1.f, N0 = 0.f, N1 = 0.f, N2 = 0.f;
        for (int j = 0; j < count; ++j) {
            float sj[3];
            sj[0] = __shfl_sync(kFull, s[0], j);
            sj[1] = __shfl_sync(kFull, s[1], j);
            sj[2] = __shfl_sync(kFull, s[2], j);
            Family fam[3];
            fisher_families(sj, fam);
            float aF = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const float x = quad_node((float)(32 * it + lane));
                const float u = 1.0f - x, v = 1.0f + x;
                float y0 = family_node(fam[0], u, v);
                float y1 = family_node(fam[1], u, v);
                float y2 = family_node(fam[2], u, v);
                if (it == 0 || it == 15) {   // trapezoid end weights 1/2
                    const float w = ((it == 0 && lane == 0) || (it == 15 && lane == 31)) ? 0.5f : 1.0f;
                    y0 *= w; y1 *= w; y2 *= w;
                }
                aF += y0;
                a0 = fmaf(x, y0, a0);
                a1 = fmaf(x, y1, a1);
                a2 = fmaf(x, y2, a2);
            }
            aF = warp_sum(aF); a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane == j) { F = aF; N0 = a0; N1 = a1; N2 = a2; }
        }

        // ---- phase 3: closing arithmetic + stores ----------------------------
        __syncwarp();
        if (mine) {
            FisherStats st = fisher_finish(s, F, N0, N1, N2);
            float U[9], V[9], M[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { U[k] = ws.uv[k * 32 + lane]; V[k] = ws.uv[(9 + k) * 32 + lane]; }
            const long long i = base + lane;
            if (p.nll) p.nll[i] = fmaf(p.overreg, st.logC, -dot);
            if (p.entropy) p.entropy[i] = st.entropy;
            if (p.logC) p.logC[i] = st.logC;
            if (p.S) { p.S[3 * i] = s[0]; p.S[3 * i + 1] = s[1]; p.S[3 * i + 2] = s[2]; }
            if (p.G) { p.G[3 * i] = st.g[0]; p.G[3 * i + 1] = st.g[1]; p.G[3 * i + 2] = st.g[2]; }
            if (want_hist) atomicAdd(&hist_s[entropy_key(st.entropy) >> kHistShift1], 1u);
            if (p.grad) {
                u_diag_vt(U, V, st.g[0], st.g[1], st.g[2], M);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float rg = p.Rgt ? ws.r[lane * 9 + k] : 0.f;
                    ws.a[lane * 9 + k] = fmaf(p.overreg, M[k], -rg);
                }
            }
            if (p.Rout) {
                u_diag_vt(U, V, 1.f, 1.f, 1.f, M);
#pragma unroll
                for (int k = 0; k < 9; ++k) ws.r[lane * 9 + k] = M[k];
            }
        }
        __syncwarp();
        if (p.grad) store_tile(p.grad + base * 9, ws.a, count, p.vec_ok, lane);
        if (p.Rout) store_tile(p.Rout + base * 9, ws.r, count, p.vec_ok, lane);
        __syncwarp();
    }

    if (bad && p.status) atomicOr(p.status, kStatusNonFinite);
    if (want_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < kHistBins1; i += kThreads) {
            const unsigned int c = hist_s[i];
            if (c) atomicAdd(p.hist + i, (unsigned long long)c);
        }
    }
}

// ---------------------------------------------------------------------------
// K1: proper SVD only (batch_torch_A_to_R, analytical_mode, proper_svd)
// thread per sample, tile of 128 records staged through shared memory
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
proper_svd_kernel(SvdArgs p) {
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const long long warps_total = (long long)gridDim.x * kWarpsPerBlock;
    const long long tiles = (p.n + 31) / 32;
    bool bad = false;
    for (long long tile = (long long)blockIdx.x * kWarpsPerBlock + warp; tile < tiles; tile += warps_total) {
        const long long base = tile * 32;
        const int count = (int)min(32LL, p.n - base);
        load_tile(ws.a, p.A + base * 9, count, p.vec_ok, lane);
        __syncwarp();
        if (lane < count) {
            float A[9], U[9], V[9], s[3], M[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) A[k] = ws.a[lane * 9 + k];
            if (!proper_svd3(A, U, V, s)) bad = true;
            const long long i = base + lane;
            if (p.S) { p.S[3 * i] = s[0]; p.S[3 * i + 1] = s[1]; p.S[3 * i + 2] = s[2]; }
            u_diag_vt(U, V, 1.f, 1.f, 1.f, M);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                ws.r[lane * 9 + k] = M[k];
                ws.uv[k * 32 + lane] = U[k];
                ws.uv[(9 + k) * 32 + lane] = V[k];
            }
        }
        __syncwarp();
        if (p.R) store_tile(p.R + base * 9, ws.r, count, p.vec_ok, lane);
        if (p.U || p.V) {
            // transpose the parked [k][lane] layout back to records
            for (int i = lane; i < count * 9; i += 32) {
                const int smp = i / 9, k = i - 9 * smp;
                if (p.U) p.U[base * 9 + i] = ws.uv[k * 32 + smp];
                if (p.V) p.V[base * 9 + i] = ws.uv[(9 + k) * 32 + smp];
            }
        }
        __syncwarp();
    }
    if (bad && p.status) atomicOr(p.status, kStatusNonFinite);
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

cudaError_t launch_fisher_fused(FisherArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int sms = sm_count();
    // resident warps the chip can hold for this kernel (occupancy-limited)
    int blocks_per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fisher_fused_kernel, kThreads, 0);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const long long resident_warps = (long long)sms * blocks_per_sm * kWarpsPerBlock;
    // small batches: spread samples over as many warps as possible (latency);
    // large batches: 32 samples per warp so the tile I/O is float4-coalesced
    long long spw = (p.n + resident_warps - 1) / resident_warps;
    if (spw < 1) spw = 1;
    if (spw > 32 || p.n >= resident_warps * 8) spw = 32;
    p.samples_per_warp = (int)spw;
    const long long tiles = (p.n + spw - 1) / spw;
    long long blocks = (tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long max_blocks = (long long)sms * blocks_per_sm;
    if (blocks > max_blocks) blocks = max_blocks;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.vec_ok = (spw == 32) && aligned(p.A) && aligned(p.Rgt) && aligned(p.grad) && aligned(p.Rout);
    fisher_fused_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_proper_svd(SvdArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int sms = sm_count();
    const long long tiles = (p.n + 31) / 32;
    long long blocks = (tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long max_blocks = (long long)sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.vec_ok = aligned(p.A) && aligned(p.R);
    proper_svd_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace suhpe
