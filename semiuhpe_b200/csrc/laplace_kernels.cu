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
// Three decompositions of the same sum, picked by launch_laplace:
//   stream  thread per TWO samples (batches from ~4k rotations): 512-thread persistent CTAs; a thread keeps its two
//           samples in the halves of packed registers and the grid point is the broadcast scalar operand, so every
//           FMA of the loop is a packed FFMA2 in the two-register-pair form (25 packed ops and 2.25 LDS.128 per
//           sample and 2 grid points); one comparison per sample and trip covers the clamp and the exponent offset;
//           block sums fold into the totals every 128 points so the fp32 summation error does not grow with N.
//           Launched as thread-block clusters of 2-8 CTAs that slice the grid and merge through distributed shared
//           memory when the batch has fewer 1024-sample tiles than the device has SMs
//   warp    warp per sample (up to a few thousand samples): grid in shared memory, lanes stride it, shuffle merge
//   block   CTA per sample (training-sized batches): lanes stride the grid in L2, shuffle + shared-memory merge
// Per-sample set-up (proper SVD, T, the ground-truth term) runs in fp64: see laplace_setup.
#include "kernels.cuh"
#include "so3_math.cuh"
#include <cooperative_groups.h>

namespace suhpe {

namespace {

#ifndef SUHPE_K2L_FORCE_STREAM
#define SUHPE_K2L_FORCE_STREAM 0     // 1: always one of the stream kernel's forms, whatever the batch size (A/B builds)
#endif
#ifndef SUHPE_K2L_DIAG_NOMUFU
#define SUHPE_K2L_DIAG_NOMUFU 0
#endif
#ifndef SUHPE_K2L_BLOCK_KERNEL
#define SUHPE_K2L_BLOCK_KERNEL 512    // batches up to this many samples take the CTA-per-sample kernel (0: never)
#endif
constexpr int kLapThreads = 256;
constexpr int kGridChunk = 4608;           // grid points resident in shared memory at once
constexpr unsigned kFull = 0xffffffffu;

// ---- packed fp32x2 helpers (Blackwell FFMA2) ------------------------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 dup(float x) { return pk(x, x); }


// block sums of the packed loop: lo / hi half = the thread's first / second sample
struct PackedSums { f2 z, c, m[9]; };

// accumulate in place: pins every accumulator to one register pair for the whole loop (without
// this the compiler renames them across the unrolled pairs and copies 18 registers back per trip)
__device__ __forceinline__ void acc_add2(f2& acc, f2 y) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(y)); }
__device__ __forceinline__ void acc_fma2(f2& acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

constexpr int kParkSlots = 23;       // per-sample state parked in shared memory: Z, C, M[9], offset, R*[9], T (fp64)

// ---------------------------------------------------------------------------------------------------
// Stream kernel: a thread owns TWO samples in the halves of its packed registers and walks the grid one point at a
// time per half.  Against the round-1 form (two grid POINTS in the halves, one sample per thread) this makes the grid
// value the broadcast scalar operand of every packed FMA (one register read less in the dot product AND in the nine
// M sums -- the three-register-pair form runs at 2/3 rate, probe variant 8), halves the LDS per (sample, point) pair
// and needs no lo/hi merge at the folds.  The parked per-sample state doubles, so the grid stays in shared memory in
// chunks (two for the 4608-point grid).
// ---------------------------------------------------------------------------------------------------
constexpr int kS2Threads = 512;      // 640 threads leave 96 registers: 410 bytes of spills in the loop
constexpr int kS2Slots = kParkSlots + 1;          // + the exponent offset in force
constexpr int kS2Chunk = ((227 * 1024 - kS2Slots * kS2Threads * 2 * 4) / 36) & ~3;   // grid points per smem chunk
__host__ __device__ constexpr int s2_grid_floats(int chunk) { return chunk * 9; }

// q of the running maximum n = max(-q^2) a sample has seen, exactly as a grid point with that n would compute it
__device__ __forceinline__ float q_of(float n) { return -(n * mufu_rsqrt(-n)); }

constexpr float kLazyRescale = 32.0f;      // the offset follows the running minimum of q only in steps of at least this much

template <bool GRAD>
__device__ __forceinline__ void park_fold2(float* park, PackedSums& s) {
    constexpr int kStride = 2 * kS2Threads;               // park[slot * kStride + half * kS2Threads]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float off = park[23 * kStride + h * kS2Threads];         // the offset the sums in registers are scaled by
        const float sc = mufu_ex2((off - park[11 * kStride + h * kS2Threads]) * kLog2e);        // first fold: 2^-inf = 0
        float v[2];
        upk(s.z, v[0], v[1]); park[h * kS2Threads] = fmaf(park[h * kS2Threads], sc, v[h]);
        if (GRAD) {
            upk(s.c, v[0], v[1]); park[kStride + h * kS2Threads] = fmaf(park[kStride + h * kS2Threads], sc, v[h]);
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                upk(s.m[i], v[0], v[1]);
                park[(2 + i) * kStride + h * kS2Threads] = fmaf(park[(2 + i) * kStride + h * kS2Threads], sc, v[h]);
            }
        }
        park[11 * kStride + h * kS2Threads] = off;
    }
    s.z = pk(0.f, 0.f);
    if (GRAD) {
        s.c = pk(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 9; ++i) s.m[i] = pk(0.f, 0.f);
    }
}

#if SUHPE_K2L_DIAG_NOMUFU   // timing diagnostic only (wrong results): what do the 16 MUFU of a trip cost?
__device__ __forceinline__ float s2_rsqrt(float x) { return x * -0.37f; }
__device__ __forceinline__ float s2_ex2(float x) { return x * 0.011f; }
#else
__device__ __forceinline__ float s2_rsqrt(float x) { return mufu_rsqrt(x); }
__device__ __forceinline__ float s2_ex2(float x) { return mufu_ex2(x); }
#endif

// the sums of NP grid points for the two samples, written phase by phase over the NP points so that the 2*NP MUFU
// results of a phase are in flight together (four warps per scheduler do not hide a serial RSQ -> EX2 chain).
// EDGE = some n is above -eps (clamp it, mask its gradient: clamp_min passes the gradient where input >= min);
// otherwise every n is below the running maximum, itself <= -eps, and rs holds 1/q computed before the branch.
template <bool GRAD, bool EDGE, int NP>
__device__ __forceinline__ void sample_pair_sums(f2* t, f2* rs, const float* r, f2 qminLpk, PackedSums& s) {
    bool live[NP][2];
    if (EDGE) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float n0, n1;
            upk(t[p], n0, n1);
            live[p][0] = n0 <= -kLapEps; live[p][1] = n1 <= -kLapEps;
            n0 = fminf(n0, -kLapEps); n1 = fminf(n1, -kLapEps);
            t[p] = pk(n0, n1);
            rs[p] = pk(s2_rsqrt(-n0), s2_rsqrt(-n1));
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) t[p] = fma2(mul2(t[p], rs[p]), dup(kLog2e), qminLpk);   // (-q + qmin) log2 e
    f2 w[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float e0, e1;
        upk(t[p], e0, e1);
        w[p] = pk(s2_ex2(e0), s2_ex2(e1));
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        w[p] = mul2(w[p], rs[p]);                                         // exp(p - c) / (-p)
        if (GRAD) rs[p] = fma2(rs[p], rs[p], rs[p]);                      // 1/q + 1/q^2
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) acc_add2(s.z, w[p]);
    if (GRAD) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            w[p] = mul2(w[p], rs[p]);
            if (EDGE) {
                float c0, c1;
                upk(w[p], c0, c1);
                w[p] = pk(live[p][0] ? c0 : 0.0f, live[p][1] ? c1 : 0.0f);
            }
            acc_add2(s.c, w[p]);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int p = 0; p < NP; ++p) acc_fma2(s.m[i], w[p], dup(r[9 * p + i]));
    }
}

// NP grid points (r: NP*9 scalars) against the thread's two samples.  The sums carry exp(-q + off): off is an
// exponent offset, not necessarily the running minimum of q -- any value that keeps the terms inside the fp32 range
// gives the same sums after the final log.  It is moved only when a point undercuts it by kLazyRescale (terms then
// stay below e^32 / q^3 < 2^84), and thr[h] is that condition expressed on n = -q^2, capped at -eps: ONE comparison
// per sample and trip decides "rescale the sums" and "a point needs the clamp", and the common path carries no clamp,
// no mask and no min/max per point.  (A new minimum of q itself turns up in some lane of a warp on almost half of the
// trips; an offset change in about one trip in twenty.)
// dot products of NP points with the two samples, speculative 1/q (right unless a clamp is needed), the trip's max n
template <int NP>
__device__ __forceinline__ void stage_dots(const f2* Apk, f2 nTpk, const float* r, f2* t, f2* rs, float* nmax) {
#pragma unroll
    for (int p = 0; p < NP; ++p) t[p] = fma2(Apk[0], dup(r[9 * p]), nTpk);
#pragma unroll
    for (int i = 1; i < 9; ++i)
#pragma unroll
        for (int p = 0; p < NP; ++p) t[p] = fma2(Apk[i], dup(r[9 * p + i]), t[p]);
    nmax[0] = nmax[1] = -INFINITY;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float n0, n1;
        upk(t[p], n0, n1);
        rs[p] = pk(s2_rsqrt(-n0), s2_rsqrt(-n1));
        nmax[0] = fmaxf(nmax[0], n0); nmax[1] = fmaxf(nmax[1], n1);
    }
}

// the one comparison per sample and trip; returns true when a point of the trip needs the clamp path
template <bool GRAD>
__device__ __forceinline__ bool offset_step(const float* nmax, PackedSums& s, float* thr, f2& offLpk, float* park) {
    bool edge = false;
    // forward-only launches take the branch warp-uniformly: the block is correct for the lanes that did not ask for
    // it (their factor is 2^0) and a uniform branch needs no reconvergence barrier around the trip (-3 %; with the
    // gradient sums in the loop the divergent form measures 1 % faster, profiles/r02w_ab_k2l_uniform.txt)
    bool moved = nmax[0] > thr[0] || nmax[1] > thr[1];
    if (!GRAD) moved = __any_sync(kFull, moved);
    if (moved) {
        float oldL[2], newL[2];
        upk(offLpk, oldL[0], oldL[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            newL[h] = oldL[h];
            if (nmax[h] > thr[h]) {
                const float off = q_of(fminf(nmax[h], -kLapEps));
                const float lim = fmaxf(off - kLazyRescale, 0.0f);
                thr[h] = -fmaxf(lim * lim, kLapEps);                      // any n above -eps trips it
                newL[h] = off * kLog2e;
                park[23 * 2 * kS2Threads + h * kS2Threads] = off;
            }
        }
        const f2 sc = pk(s2_ex2(newL[0] - oldL[0]), s2_ex2(newL[1] - oldL[1]));   // first trip: 2^-inf = 0
        s.z = mul2(s.z, sc);
        if (GRAD) {
            s.c = mul2(s.c, sc);
#pragma unroll
            for (int i = 0; i < 9; ++i) s.m[i] = mul2(s.m[i], sc);
        }
        offLpk = pk(newL[0], newL[1]);
        edge = nmax[0] > -kLapEps || nmax[1] > -kLapEps;
    }
    // a new minimum turns up in some lane of a warp on almost half of the trips, so that block stays small and
    // rejoins; only the clamp (a grid point within 1e-4 rad-ish of the mode: next to never) takes the long way round
    if (!GRAD) edge = __any_sync(kFull, edge);
    return edge;
}

template <bool GRAD, int NP>
__device__ __forceinline__ void sample_pair_points(const f2* Apk, f2 nTpk, const float* r, PackedSums& s, float* thr, f2& offLpk,
                                                   float* park) {
    f2 t[NP], rs[NP];
    float nmax[2];
    stage_dots<NP>(Apk, nTpk, r, t, rs, nmax);
    if (offset_step<GRAD>(nmax, s, thr, offLpk, park)) sample_pair_sums<GRAD, true, NP>(t, rs, r, offLpk, s);      // next to never
    else sample_pair_sums<GRAD, false, NP>(t, rs, r, offLpk, s);
}

template <bool GRAD>
__global__ void __launch_bounds__(kS2Threads, 1)
laplace_stream2_kernel(LaplaceArgs p, int chunk, int slices, int per_slice) {
    extern __shared__ __align__(16) float gp[];      // [chunk][9]: grid points in their natural order
    constexpr int kStride = 2 * kS2Threads;
    float* park = gp + s2_grid_floats(chunk) + threadIdx.x;              // park[slot * kStride + half * kS2Threads]: conflict-free
    const long long tiles = (p.n + 2 * kS2Threads - 1) / (2 * kS2Threads);
    // slices > 1: launched as thread-block clusters of `slices` CTAs.  The CTAs of a cluster take the same tile of
    // samples and one slice of the grid each (per_slice points, a multiple of 4); their totals meet in the cluster's
    // first CTA through distributed shared memory.  This is what fills the SMs when the batch has fewer tiles than that.
    const int rank = slices > 1 ? (int)(blockIdx.x % (unsigned)slices) : 0;
    const int first = rank * per_slice;
    const int last = slices > 1 ? min(p.N, first + per_slice) : p.N;    // this CTA's grid points: [first, last)
    const bool single_chunk = last - first <= chunk;
    bool bad = false;

    auto load_chunk = [&](int c0, int cn) {
        const float* src = p.grid + (size_t)c0 * 9;
        for (int i = threadIdx.x; i < cn * 9; i += kS2Threads) gp[i] = __ldg(src + i);
    };
    if (single_chunk) { load_chunk(first, max(last - first, 0)); __syncthreads(); }

    for (long long tile = blockIdx.x / (unsigned)slices; tile < tiles; tile += gridDim.x / (unsigned)slices) {
        const long long sample0 = tile * (2 * kS2Threads) + 2 * threadIdx.x;
        const bool valid[2] = {sample0 < p.n, sample0 + 1 < p.n};
        f2 Apk[9], nTpk;
        {
            float A[2][9], Tf[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float Rs[9];
                double Td;
#pragma unroll
                for (int i = 0; i < 9; ++i) A[h][i] = valid[h] ? __ldg(p.A + (sample0 + h) * 9 + i) : ((i % 4 == 0) ? 1.f : 0.f);
                if (!laplace_setup(A[h], Rs, &Td) && valid[h]) bad = true;
                Tf[h] = (float)Td;
#pragma unroll
                for (int i = 0; i < 11; ++i) park[i * kStride + h * kS2Threads] = 0.f;
                park[11 * kStride + h * kS2Threads] = INFINITY;
#pragma unroll
                for (int i = 0; i < 9; ++i) park[(12 + i) * kStride + h * kS2Threads] = Rs[i];
                park[21 * kStride + h * kS2Threads] = __int_as_float(__double2loint(Td));
                park[22 * kStride + h * kS2Threads] = __int_as_float(__double2hiint(Td));
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) Apk[i] = pk(A[0][i], A[1][i]);
            nTpk = pk(-Tf[0], -Tf[1]);
        }
        float thr[2] = {-INFINITY, -INFINITY};
        f2 offLpk = pk(INFINITY, INFINITY);
        PackedSums s;
        s.z = s.c = pk(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 9; ++i) s.m[i] = pk(0.f, 0.f);

        for (int c0 = first; c0 < last; c0 += chunk) {
            const int cn = min(chunk, last - c0);
            if (!single_chunk) { __syncthreads(); load_chunk(c0, cn); __syncthreads(); }
            const int groups = cn >> 2;                       // 4 points = 36 floats = 9 float4
            const float4* g4 = reinterpret_cast<const float4*>(gp);
            for (int g0 = 0; g0 < groups; g0 += 32) {         // fold into the totals every 128 points
                const int g1 = min(g0 + 32, groups);
#pragma unroll 1
                for (int g = g0; g < g1; ++g) {
                    float r[36];
#pragma unroll
                    for (int v = 0; v < 9; ++v) {
                        const float4 x = g4[g * 9 + v];
                        r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                    }
                    sample_pair_points<GRAD, 4>(Apk, nTpk, r, s, thr, offLpk, park);
                }
                park_fold2<GRAD>(park, s);
            }
            if (groups * 4 < cn) {                            // up to 3 trailing points
                for (int k = groups * 4; k < cn; ++k) {
                    float r[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) r[i] = gp[k * 9 + i];
                    sample_pair_points<GRAD, 1>(Apk, nTpk, r, s, thr, offLpk, park);
                }
                park_fold2<GRAD>(park, s);
            }
        }

        if (slices > 1) {
            // totals of the other slices: Z, C, M[9] and the offset they are scaled by (slots 0..11) of the same thread
            // in the peer CTA, rebased to the smaller offset.  A slice without points reports offset +inf and zeros.
            namespace cg = cooperative_groups;
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();
            if (rank == 0) {
                for (int peer = 1; peer < slices; ++peer) {
                    const float* theirs = cluster.map_shared_rank(park, peer);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int o = h * kS2Threads;
                        const float mine_off = park[11 * kStride + o], their_off = theirs[11 * kStride + o];
                        const float off = fminf(mine_off, their_off);
                        const float fm = mufu_ex2((off - mine_off) * kLog2e), ft = mufu_ex2((off - their_off) * kLog2e);
#pragma unroll
                        for (int i = 0; i < (GRAD ? 11 : 1); ++i)
                            park[i * kStride + o] = fmaf(park[i * kStride + o], fm, theirs[i * kStride + o] * ft);
                        park[11 * kStride + o] = off;
                    }
                }
            }
            cluster.sync();                                   // the peers' parked state stays in place until it has been read
            if (rank != 0) continue;
        }

#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!valid[h]) continue;
            const long long sample = sample0 + h;
            float A[9], Rg[9], Rs[9], grad[9], nll, logF;
#pragma unroll
            for (int i = 0; i < 9; ++i) { float lo, hi; upk(Apk[i], lo, hi); A[i] = h ? hi : lo; }
            LaplaceAccum a;
            a.qmin = park[11 * kStride + h * kS2Threads]; a.Z = park[h * kS2Threads]; a.C = park[kStride + h * kS2Threads];
#pragma unroll
            for (int i = 0; i < 9; ++i) { a.M[i] = park[(2 + i) * kStride + h * kS2Threads]; Rs[i] = park[(12 + i) * kStride + h * kS2Threads]; }
            const double Td = __hiloint2double(__float_as_int(park[22 * kStride + h * kS2Threads]), __float_as_int(park[21 * kStride + h * kS2Threads]));
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

// Block kernel (training-sized batches, a few hundred samples at most): one 256-thread CTA per sample, the lanes stride
// the grid straight out of global memory (it is L2-resident after the first CTA; staging 166 KB of it in shared memory
// per CTA, as the warp kernel does, costs more than the whole sum at this size), warp partials merge by shuffle and the
// eight warps through 384 bytes of shared memory.
__global__ void __launch_bounds__(kLapThreads)
laplace_block_kernel(LaplaceArgs p) {
    __shared__ float part[kLapThreads / 32][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool bad = false;
    for (long long sample = blockIdx.x; sample < p.n; sample += gridDim.x) {
        float A[9], Rs[9];
        double Td;
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = __ldg(p.A + sample * 9 + i);
        if (!laplace_setup(A, Rs, &Td)) bad = true;
        const float T = (float)Td;
        LaplaceAccum a;
        laplace_accum_init(a);
        for (int k = threadIdx.x; k < p.N; k += kLapThreads) {
            float r[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) r[i] = __ldg(p.grid + (size_t)k * 9 + i);
            laplace_accum_point(a, A, T, r);
        }
        laplace_accum_flush(a);
        auto warp_merge = [&]() {            // the lanes' partial sums under their common minimum
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
        };
        warp_merge();
        if (lane == 0) {
            part[warp][0] = a.qmin; part[warp][1] = a.Z; part[warp][2] = a.C;
#pragma unroll
            for (int i = 0; i < 9; ++i) part[warp][3 + i] = a.M[i];
        }
        __syncthreads();
        if (warp == 0) {
            laplace_accum_init(a);
            if (lane < kLapThreads / 32) {
                a.qmin = part[lane][0]; a.Z = part[lane][1]; a.C = part[lane][2];
#pragma unroll
                for (int i = 0; i < 9; ++i) a.M[i] = part[lane][3 + i];
            }
            warp_merge();
            if (lane == 0) {
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
        __syncthreads();                     // part[] is rewritten by the next sample
    }
    if (bad && p.status && threadIdx.x == 0) atomicOr(p.status, kStatusNonFinite);
}

}  // namespace

cudaError_t launch_laplace(LaplaceArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    if (p.N <= 0) return cudaErrorInvalidValue;
    const int sms = device_sm_count();
    cudaError_t err;
#if SUHPE_K2L_BLOCK_KERNEL
    if (!SUHPE_K2L_FORCE_STREAM && p.n <= (long long)SUHPE_K2L_BLOCK_KERNEL) {
        const unsigned blocks = (unsigned)(p.n < 4ll * sms ? p.n : 4ll * sms);
        laplace_block_kernel<<<blocks, kLapThreads, 0, stream>>>(p);
        return cudaGetLastError();
    }
#endif
    {
        // Which decomposition?  Estimated cost in units of one stream-kernel tile (1024 samples against the whole grid:
        // 0.67 ms for 4608 points), from the timings in profiles/r02af_k2l_batch_sweep.txt:
        //   warp kernel      0.03 + 4.0e-5 per sample (up to 256 samples per SM)
        //   stream kernel    1/C + 0.02 + 0.007 C per round of clusters, C CTAs sharing a tile and slicing the grid
        //                    (every CTA of a cluster repeats the per-sample set-up of the tile); C = 1: plain launch
        // The clusters are what fills the SMs when the batch has fewer than one tile per SM, and what trims the last,
        // mostly empty round of a larger one.
        const long long tiles = (p.n + 2 * kS2Threads - 1) / (2 * kS2Threads);
        auto kernel = p.grad ? laplace_stream2_kernel<true> : laplace_stream2_kernel<false>;
        constexpr size_t kSmemMax = ((size_t)s2_grid_floats(kS2Chunk) + (size_t)kS2Slots * kS2Threads * 2) * sizeof(float);
        static unsigned long long attr_done2[2] = {0ull, 0ull};
        err = allow_dynamic_smem(kernel, kSmemMax, attr_done2[p.grad ? 1 : 0]);
        if (err != cudaSuccess) return err;
        const bool warp_allowed = !SUHPE_K2L_FORCE_STREAM && p.n < (long long)sms * kLapThreads;
        double best = 0.03 + 4.0e-5 * (double)p.n;
        int best_slices = 0, best_per = 0, best_chunk = 0;
        long long best_clusters = 0;
        for (int c = 1; c <= 8; c *= 2) {
            const int per = c == 1 ? p.N : (((p.N + c - 1) / c + 3) & ~3);
            if (c > 1 && (per < 64 || (long long)(c - 1) * per >= p.N)) break;         // slices too thin, or an empty one
            const int chunk = per < kS2Chunk ? ((per + 3) & ~3) : (per <= 2 * kS2Chunk ? ((((per + 1) / 2) + 3) & ~3) : kS2Chunk);
            long long clusters = sms;
            if (c > 1) {
                // how many clusters of c such CTAs fit the device at once (they must share a GPC): asked once per
                // (device, kernel, c) for the largest footprint
                static int fit[64][2][4];
                int dev = 0;
                cudaGetDevice(&dev);
                int& cached = fit[dev & 63][p.grad ? 1 : 0][c == 2 ? 1 : c == 4 ? 2 : 3];
                if (cached == 0) {
                    cudaLaunchConfig_t q = {};
                    q.gridDim = dim3((unsigned)(c * sms), 1, 1);
                    q.blockDim = dim3(kS2Threads, 1, 1);
                    q.dynamicSmemBytes = kSmemMax;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = (unsigned)c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    q.attrs = at; q.numAttrs = 1;
                    int m = 0;
                    if (cudaOccupancyMaxActiveClusters(&m, kernel, &q) != cudaSuccess) { (void)cudaGetLastError(); m = 0; }
                    cached = m > 0 ? m : -1;
                }
                if (cached < 0) continue;
                clusters = cached;
            }
            const long long rounds = (tiles + clusters - 1) / clusters;
            const double cost = (double)rounds * (1.0 / c + (c > 1 ? 0.02 + 0.007 * c : 0.0));
            if (cost < best || (!warp_allowed && best_slices == 0)) {
                best = cost; best_slices = c; best_per = per; best_chunk = chunk;
                best_clusters = tiles < clusters ? tiles : clusters;
            }
        }
        if (best_slices > 0) {
            const size_t smem = ((size_t)s2_grid_floats(best_chunk) + (size_t)kS2Slots * kS2Threads * 2) * sizeof(float);
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3((unsigned)(best_clusters * best_slices), 1, 1);
            q.blockDim = dim3(kS2Threads, 1, 1);
            q.dynamicSmemBytes = smem;
            q.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)best_slices; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.attrs = at; q.numAttrs = best_slices > 1 ? 1 : 0;
            err = cudaLaunchKernelEx(&q, kernel, p, best_chunk, best_slices, best_per);
            if (err != cudaSuccess && best_slices > 1) {
                // the cluster launch was refused (a partitioned or busy device can fit fewer clusters than the occupancy
                // query promised): run the same kernel unclustered, one tile per CTA against the whole grid
                (void)cudaGetLastError();
                const int chunk = p.N < kS2Chunk ? ((p.N + 3) & ~3) : (p.N <= 2 * kS2Chunk ? ((((p.N + 1) / 2) + 3) & ~3) : kS2Chunk);
                q.gridDim = dim3((unsigned)(tiles < sms ? tiles : sms), 1, 1);
                q.dynamicSmemBytes = ((size_t)s2_grid_floats(chunk) + (size_t)kS2Slots * kS2Threads * 2) * sizeof(float);
                q.numAttrs = 0;
                err = cudaLaunchKernelEx(&q, kernel, p, chunk, 1, p.N);
            }
            return err != cudaSuccess ? err : cudaGetLastError();
        }
    }
    {
        const int chunk = p.N < kGridChunk ? p.N : kGridChunk;
        const int stride = (chunk + 3) & ~3;
        const size_t smem = (size_t)9 * stride * sizeof(float);
        static unsigned long long attr_done32 = 0ull;
        err = allow_dynamic_smem(laplace_kernel<32>, (size_t)9 * kGridChunk * sizeof(float), attr_done32);
        if (err != cudaSuccess) return err;
        // spread the samples over the SMs, 8 per block pass
        const long long tiles = (p.n + (kLapThreads / 32) - 1) / (kLapThreads / 32);
        const unsigned blocks = (unsigned)(tiles < sms ? tiles : sms);
        laplace_kernel<32><<<blocks, kLapThreads, smem, stream>>>(p, chunk, stride);
    }
    return cudaGetLastError();
}

}  // namespace suhpe
