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

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2) ------------------
// One issue slot does two FMAs: the quadrature loop packs the two nodes a lane owns
// in a pair of iterations (it, it+1) into one 64-bit register pair.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 dup(float x) { return pk(x, x); }
__device__ __forceinline__ f2 rsq2(f2 a) { float lo, hi; upk(a, lo, hi); return pk(mufu_rsqrt(lo), mufu_rsqrt(hi)); }
__device__ __forceinline__ f2 ex22(f2 a) { float lo, hi; upk(a, lo, hi); return pk(mufu_ex2(lo), mufu_ex2(hi)); }

// same operation order as i0e_large / i0_small_poly in so3_math.cuh, two nodes at a time
__device__ __forceinline__ f2 large2(f2 a) {
    const f2 rs = rsq2(a);
    const f2 r = mul2(rs, rs);
    f2 p = dup(kLg8);
    p = fma2(p, r, dup(kLg7)); p = fma2(p, r, dup(kLg6)); p = fma2(p, r, dup(kLg5)); p = fma2(p, r, dup(kLg4));
    p = fma2(p, r, dup(kLg3)); p = fma2(p, r, dup(kLg2)); p = fma2(p, r, dup(kLg1)); p = fma2(p, r, dup(kLg0));
    return mul2(p, rs);
}
__device__ __forceinline__ f2 small2(f2 a) {
    const f2 q = mul2(a, a);
    f2 p = dup(kSm6);
    p = fma2(p, q, dup(kSm5)); p = fma2(p, q, dup(kSm4)); p = fma2(p, q, dup(kSm3));
    p = fma2(p, q, dup(kSm2)); p = fma2(p, q, dup(kSm1)); p = fma2(p, q, dup(1.0f));
    return p;
}

// node table in shared memory: tab[pair][lane] = (u_lo, u_hi, v_lo, v_hi) for the nodes
// 64*pair + lane (lo) and 64*pair + 32 + lane (hi);  u = 1-x, v = 1+x
struct __align__(16) NodePair { f2 u, v; };

enum RunType { kLS = 0, kLL = 1, kSS = 2, kSL = 3, kMixed = 4 };

struct FamilyPacked { f2 fd, fs, ncl, ncdl, nfsl; float fdf, fsf, nclf, ncdlf, nfslf; };

// one run of consecutive node pairs of uniform type T
template <int T>
__device__ __forceinline__ void quad_run(const NodePair* __restrict__ tab, int lo, int hi, const FamilyPacked& k,
                                         f2& accY, f2& accUY) {
#pragma unroll 1
    for (int p = lo; p < hi; ++p) {
        const NodePair n = tab[p * 32];
        const f2 ad = mul2(k.fd, n.u), as = mul2(k.fs, n.v);
        f2 y;
        if (T == kMixed) {
            // per-lane, per-half switch: both polynomials, selected
            const f2 Ld = large2(ad), Ls = large2(as), Sd = small2(ad), Ss = small2(as);
            float adl, adh, asl, ash, a0, a1, b0, b1;
            upk(ad, adl, adh); upk(as, asl, ash);
            const bool dl = adl <= kBesselSwitch, dh = adh <= kBesselSwitch;
            const bool sl = asl <= kBesselSwitch, sh = ash <= kBesselSwitch;
            upk(Ld, a0, a1); upk(Sd, b0, b1);
            const f2 Vd = pk(dl ? b0 : a0, dh ? b1 : a1);
            upk(Ls, a0, a1); upk(Ss, b0, b1);
            const f2 Vs = pk(sl ? b0 : a0, sh ? b1 : a1);
            const f2 kd = pk(dl ? k.ncdlf : k.nclf, dh ? k.ncdlf : k.nclf);
            const f2 ks = pk(sl ? k.nfslf : 0.0f, sh ? k.nfslf : 0.0f);
            const f2 e = fma2(ks, n.v, mul2(kd, n.u));
            y = mul2(mul2(Vd, Vs), ex22(e));
        } else {
            const bool d_small = (T == kSS || T == kSL), s_small = (T == kSS || T == kLS);
            const f2 Vd = d_small ? small2(ad) : large2(ad);
            const f2 Vs = s_small ? small2(as) : large2(as);
            f2 e = mul2(d_small ? k.ncdl : k.ncl, n.u);
            if (s_small) e = fma2(k.nfsl, n.v, e);
            y = mul2(mul2(Vd, Vs), ex22(e));
        }
        accY = add2(accY, y);
        accUY = fma2(n.u, y, accUY);
    }
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
    __shared__ NodePair node_tab[8 * 32];
    __shared__ unsigned int hist_s[kHistBins1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const bool want_hist = (p.hist != nullptr);
    if (want_hist) {
        for (int i = threadIdx.x; i < kHistBins1; i += kThreads) hist_s[i] = 0u;
    }
    // quadrature nodes of every (pair, lane): x rounded exactly like the reference
    for (int i = threadIdx.x; i < 8 * 32; i += kThreads) {
        const int pr = i >> 5, ln = i & 31;
        const float x0 = quad_node((float)(64 * pr + ln)), x1 = quad_node((float)(64 * pr + 32 + ln));
        node_tab[i].u = pk(add_rn(1.0f, -x0), add_rn(1.0f, -x1));
        node_tab[i].v = pk(add_rn(1.0f, x0), add_rn(1.0f, x1));
    }
    __syncthreads();
    const NodePair* tab = node_tab + lane;
    // trapezoid end points (weight 1/2): node 0 and node 511
    float u_first, u_last, v_first, v_last, dummy;
    upk(node_tab[0].u, u_first, dummy);       upk(node_tab[0].v, v_first, dummy);
    upk(node_tab[7 * 32 + 31].u, dummy, u_last); upk(node_tab[7 * 32 + 31].v, dummy, v_last);

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
        // per-sample run descriptors and trapezoid end-point corrections, still thread per sample
        unsigned runs0, runs1, runs2;
        float cY0, cUY0, cN1, cN2;
        {
            Family fam[3];
            fisher_families(s, fam);
            runs0 = family_runs(fam[0]); runs1 = family_runs(fam[1]); runs2 = family_runs(fam[2]);
            const float f0 = fisher_node(fam[0], u_first, v_first), l0 = fisher_node(fam[0], u_last, v_last);
            const float f1 = fisher_node(fam[1], u_first, v_first), l1 = fisher_node(fam[1], u_last, v_last);
            const float f2v = fisher_node(fam[2], u_first, v_first), l2 = fisher_node(fam[2], u_last, v_last);
            cY0 = 0.5f * (f0 + l0);
            cUY0 = 0.5f * fmaf(u_first, f0, u_last * l0);
            cN1 = 0.5f * ((f1 + l1) - fmaf(u_first, f1, u_last * l1));
            cN2 = 0.5f * ((f2v + l2) - fmaf(u_first, f2v, u_last * l2));
        }

        // ---- phase 2: quadrature (warp per sample) ---------------------------
        float Y0 = 1.f, UY0 = 0.f, N1 = 0.f, N2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < count; ++j) {
            const float s0 = __shfl_sync(kFull, s[0], j);
            const float s1 = __shfl_sync(kFull, s[1], j);
            const float s2 = __shfl_sync(kFull, s[2], j);
            const unsigned r0 = __shfl_sync(kFull, runs0, j);
            const unsigned r1 = __shfl_sync(kFull, runs1, j);
            const unsigned r2 = __shfl_sync(kFull, runs2, j);
            float pY0 = 0.f, pUY0 = 0.f, pN1 = 0.f, pN2 = 0.f;
#pragma unroll 1
            for (int f = 0; f < 3; ++f) {
                // family f: (lo, hi, c) as in fisher_families()
                const float lo = (f == 2) ? s1 : s2;
                const float hi = (f == 0) ? s1 : s0;
                const float c = (f == 0) ? s0 + s2 : s1 + s2;
                const unsigned runs = (f == 0) ? r0 : ((f == 1) ? r1 : r2);
                const Family fm = make_family(lo, hi, c);
                FamilyPacked k;
                k.fd = dup(fm.fd); k.fs = dup(fm.fs); k.ncl = dup(fm.ncl); k.ncdl = dup(fm.ncdl); k.nfsl = dup(fm.nfsl);
                k.fdf = fm.fd; k.fsf = fm.fs; k.nclf = fm.ncl; k.ncdlf = fm.ncdl; k.nfslf = fm.nfsl;
                const int b1 = runs & 15, m0 = (runs >> 4) & 15, m1 = (runs >> 8) & 15, b4 = (runs >> 12) & 15;
                f2 accY = pk(0.f, 0.f), accUY = pk(0.f, 0.f);
                quad_run<kLS>(tab, 0, b1, k, accY, accUY);
                quad_run<kMixed>(tab, b1, m0, k, accY, accUY);
                if (runs & (1u << 16)) quad_run<kSS>(tab, m0, m1, k, accY, accUY);
                else                   quad_run<kLL>(tab, m0, m1, k, accY, accUY);
                quad_run<kMixed>(tab, m1, b4, k, accY, accUY);
                quad_run<kSL>(tab, b4, 8, k, accY, accUY);
                float ylo, yhi, ulo, uhi;
                upk(accY, ylo, yhi); upk(accUY, ulo, uhi);
                const float Y = ylo + yhi, UY = ulo + uhi;
                if (f == 0) { pY0 = Y; pUY0 = UY; }
                else if (f == 1) pN1 = Y - UY;
                else pN2 = Y - UY;
            }
            pY0 = warp_sum(pY0); pUY0 = warp_sum(pUY0); pN1 = warp_sum(pN1); pN2 = warp_sum(pN2);
            if (lane == j) { Y0 = pY0; UY0 = pUY0; N1 = pN1; N2 = pN2; }
        }

        // ---- phase 3: closing arithmetic + stores ----------------------------
        __syncwarp();
        if (mine) {
            const float F = Y0 - cY0;
            const float N0 = F - (UY0 - cUY0);
            FisherStats st = fisher_finish(s, F, N0, N1 - cN1, N2 - cN2);
            float U[9], V[9], M[9];
            const long long i = base + lane;
            if (p.nll) p.nll[i] = fmaf(p.overreg, st.logC, -dot);
            if (p.entropy) p.entropy[i] = st.entropy;
            if (p.logC) p.logC[i] = st.logC;
            if (p.S) { p.S[3 * i] = s[0]; p.S[3 * i + 1] = s[1]; p.S[3 * i + 2] = s[2]; }
            if (p.G) { p.G[3 * i] = st.g[0]; p.G[3 * i + 1] = st.g[1]; p.G[3 * i + 2] = st.g[2]; }
            if (want_hist) atomicAdd(&hist_s[entropy_key(st.entropy) >> kHistShift1], 1u);
            if (p.grad || p.Rout) {
#pragma unroll
                for (int k = 0; k < 9; ++k) { U[k] = ws.uv[k * 32 + lane]; V[k] = ws.uv[(9 + k) * 32 + lane]; }
            }
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
