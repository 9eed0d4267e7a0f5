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
//   phase 1  thread-per-sample : coalesced float4 tile load -> smem -> 9 regs, Hestenes SVD in
//            registers, U/V parked in smem; per family the (up to) three pair-aligned runs of
//            uniform type, their constants, the negligible-node cut and the edge nodes (pairs that
//            straddle a type boundary) are derived ONCE by the owning thread; constants and run
//            words parked as 3 float4 (family_plan: so3_math.cuh)
//   phase 2  warp-per-sample   : the warp replays the sample's runs; a pass covers 64 (or 32)
//            consecutive node pairs, lane l taking adjacent node pairs in the halves of f32x2
//            registers.  Inside a run every node executes the same straight-line FFMA2 body: the
//            large-argument Bessel branch needs no MUFU (1/u and log2 rsqrt(u) come from a
//            shared-memory node table, 1/f and rsqrt(f) are per-sample constants), so one
//            MUFU.EX2 per node is the only SFU work; pairs past the end of a run (the tail of its
//            last pass) are dropped by one select per pair on the ALU pipe
//   phase 3  thread-per-sample : closing arithmetic, gradient
//            -R_gt + overreg * U diag(g) V^T, entropy, staged float4 stores; with `hist` the first
//            radix-select digit of the entropy key is counted right here (warp-aggregated RED.ADD)
// Geometry: one persistent CTA of 24 warps per SM (78 registers: the per-sample state of phase 3
// rides through phase 2 in shared memory and the 64-bit tile base is re-derived after it), 27 KB of
// node tables shared by the CTA + 8.25 KB of scratch per warp in dynamic shared memory (225 KB).
// What the replay is sensitive to (measured): instruction count and code size -- not
// latency, not the SFU.  Hence 64-pair passes that are not unrolled, Horner chains written round
// robin so that ptxas keeps them interleaved under the register cap, and one butterfly for the
// four warp sums of a rotation.
#include "kernels.cuh"
#include "so3_math.cuh"
#include <cstddef>

namespace suhpe {

namespace {

#ifndef SUHPE_K2_WARPS
#define SUHPE_K2_WARPS 24
#endif
#ifndef SUHPE_K2_RR
#define SUHPE_K2_RR 1
#endif

constexpr int kWarpsPerBlock = SUHPE_K2_WARPS;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;

// per-warp shared scratch
constexpr int kTileFloats = 32 * 9;              // one 32-sample tile of 3x3 records
struct __align__(16) WarpScratch {
    float a[kTileFloats];      // A in -> (phase 2) per-sample state s[3], <A,R_gt>, 4 sums -> gradient out; [sample*9 + k]
    float r[kTileFloats];      // R_gt in -> projected rotation out
    float uv[12 * 32];         // first two columns of U and of V parked during the quadrature, [k][lane]
                               // (u2 = u0 x u1 is how the SVD defines it; v2 = v0 x v1 holds because det V = +1)
    float4 desc[9 * 32];       // 3 families x 3 float4 of run constants, [f*3+q][sample]
};

// Node tables in shared memory, one float4 per PAIR of adjacent nodes (2m, 2m+1) so that a
// lane's LDS.128 lands directly in two f32x2 register pairs.  A run body needs at most two
// such loads per 64 nodes:
//   LL  a = (iu, iv)  b = (u, Lu+Lv)        LS  a = (iu, v)  b = (u, Lu)
//   SL  a = (iv, Lv)  b = LS's b (u, -)      SS  a = (u, v)
// Every column has the same 16-byte stride, so a run walks ONE address register and reaches its
// two columns through immediate offsets.
// 288 entries: a run's last pass may read up to 31 pairs past node 511 (discarded).
constexpr int kTabPairs = 288;
struct __align__(16) QuadTables {
    float4 LLa[kTabPairs], LLb[kTabPairs], LSa[kTabPairs], LSb[kTabPairs];
    float4 SLa[kTabPairs], SSa[kTabPairs];
    NodeVals first, last;                     // nodes 0 and 511 (trapezoid half weights)
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
// One issue slot carries two FMAs: a lane evaluates the two adjacent nodes (2m, 2m+1)
// in the halves of one 64-bit register pair.  Scalars broadcast for free (R.F32 operand).
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 dup(float x) { return pk(x, x); }
__device__ __forceinline__ f2 ex22(f2 a) { float lo, hi; upk(a, lo, hi); return pk(mufu_ex2(lo), mufu_ex2(hi)); }

// same operation order as large_poly / i0_small_poly in so3_math.cuh, two nodes at a time
__device__ __forceinline__ f2 large2(f2 r) {
    f2 p = dup(kLg8);
    p = fma2(p, r, dup(kLg7)); p = fma2(p, r, dup(kLg6)); p = fma2(p, r, dup(kLg5)); p = fma2(p, r, dup(kLg4));
    p = fma2(p, r, dup(kLg3)); p = fma2(p, r, dup(kLg2)); p = fma2(p, r, dup(kLg1)); p = fma2(p, r, dup(kLg0));
    return p;
}
__device__ __forceinline__ f2 small2(f2 a) {
    const f2 q = mul2(a, a);
    f2 p = dup(kSm6);
    p = fma2(p, q, dup(kSm5)); p = fma2(p, q, dup(kSm4)); p = fma2(p, q, dup(kSm3));
    p = fma2(p, q, dup(kSm2)); p = fma2(p, q, dup(kSm1)); p = fma2(p, q, dup(1.0f));
    return p;
}

struct RunConsts { float fd, fs, ifd, ifs, k1L, k1S, k2; };

// unconditional in-place accumulation (body probe)
__device__ __forceinline__ void acc_add2(f2& acc, f2 y) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(y)); }
__device__ __forceinline__ void acc_fma2(f2& acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

// Accumulate one node pair if it lies inside the run: `left` = pairs from this lane's first pair of
// the pass to the run's end, the pair of half w counts iff left > 32*w.  One predicate per pair; the
// select (ALU pipe) also keeps whatever the masked lanes computed from table entries past the run
// (inf / NaN included) away from the accumulators.
template <int LIM>
__device__ __forceinline__ void acc_pair(f2& accY, f2& accUY, f2 y, f2 u, int left) {
    float ylo, yhi;
    upk(y, ylo, yhi);
    const bool in = left > LIM;
    const f2 ym = pk(in ? ylo : 0.0f, in ? yhi : 0.0f);
    acc_add2(accY, ym);
    acc_fma2(accUY, u, ym);
}

// shared-memory loads by 32-bit shared-window address (no generic->shared conversion in the hot loop)
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
template <unsigned OFF>
__device__ __forceinline__ float4 lds128_at(unsigned a) {       // [a + OFF], OFF folded into the instruction
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
    return v;
}
// ordered after the phase-1 stores of the same warp (compiler barrier)
__device__ __forceinline__ float4 lds128_ordered(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

// One pass: W*32 node pairs of a run of type T; lane l takes the adjacent node pairs at
// ta, ta+512 B (nodes 2m, 2m+1).  W = 2 keeps four independent Horner chains in flight per
// lane.  Runs are pair-aligned (family_plan), so inside a run every pair is fully valid; pairs past
// the run's end -- the tail of its last pass -- are dropped by the predicate of acc_pair.  The FMA
// pipe sees one straight-line FFMA2 body per type.
//   ta   : shared address of this lane's first pair, relative to the start of QuadTables
//   left : pairs from this lane's first pair of the pass to the run's end (<= 0: outside)
template <int T, int W>
__device__ __forceinline__ void quad_pass(unsigned ta, int left, const RunConsts& k, f2& accY, f2& accUY) {
    constexpr unsigned offA = (T == kLL) ? offsetof(QuadTables, LLa) : (T == kLS) ? offsetof(QuadTables, LSa)
                            : (T == kSL) ? offsetof(QuadTables, SLa) : offsetof(QuadTables, SSa);
    constexpr unsigned offB = (T == kLL) ? offsetof(QuadTables, LLb) : offsetof(QuadTables, LSb);
    float4 a[W], b[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        a[w] = (w == 0) ? lds128_at<offA>(ta) : lds128_at<offA + 512>(ta);
        if (T != kSS) b[w] = (w == 0) ? lds128_at<offB>(ta) : lds128_at<offB + 512>(ta);
    }
    f2 pd[W], ps[W], e[W], u[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if (T == kLL) {
            u[w] = pk(b[w].x, b[w].y);
            pd[w] = mul2(dup(k.ifd), pk(a[w].x, a[w].y));
            ps[w] = mul2(dup(k.ifs), pk(a[w].z, a[w].w));
            e[w] = fma2(dup(k.k1L), u[w], pk(b[w].z, b[w].w));
        } else if (T == kLS) {
            u[w] = pk(b[w].x, b[w].y);
            const f2 v = pk(a[w].z, a[w].w);
            pd[w] = mul2(dup(k.ifd), pk(a[w].x, a[w].y));
            ps[w] = mul2(dup(k.fs), v);
            e[w] = fma2(dup(k.k2), v, fma2(dup(k.k1L), u[w], pk(b[w].z, b[w].w)));
        } else if (T == kSL) {
            u[w] = pk(b[w].x, b[w].y);
            pd[w] = mul2(dup(k.fd), u[w]);
            ps[w] = mul2(dup(k.ifs), pk(a[w].x, a[w].y));
            e[w] = fma2(dup(k.k1S), u[w], pk(a[w].z, a[w].w));
        } else {
            u[w] = pk(a[w].x, a[w].y);
            const f2 v = pk(a[w].z, a[w].w);
            pd[w] = mul2(dup(k.fd), u[w]);
            ps[w] = mul2(dup(k.fs), v);
            e[w] = fma2(dup(k.k2), v, mul2(dup(k.k1S), u[w]));
        }
    }
#if SUHPE_K2_RR
    // The 2W Horner chains advance one step each in turn (round robin): written this way the chains stay
    // interleaved in the SASS -- evaluating one polynomial after the other left the last chain running
    // alone for 8 dependent FFMA2s (4-5 cycles each against the pipe's 2), which the other warps of the
    // scheduler do not always cover.  Same operations per chain, same order inside a chain: bit-identical.
    {
        constexpr bool dL = (T == kLL || T == kLS), sL = (T == kLL || T == kSL);
        constexpr float cL[9] = {kLg8, kLg7, kLg6, kLg5, kLg4, kLg3, kLg2, kLg1, kLg0};
        constexpr float cS[7] = {kSm6, kSm5, kSm4, kSm3, kSm2, kSm1, 1.0f};
        f2 xd[W], xs[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            xd[w] = dL ? pd[w] : mul2(pd[w], pd[w]);
            xs[w] = sL ? ps[w] : mul2(ps[w], ps[w]);
            pd[w] = dup(dL ? cL[0] : cS[0]);
            ps[w] = dup(sL ? cL[0] : cS[0]);
        }
#pragma unroll
        for (int step = 1; step <= 8; ++step) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (dL) pd[w] = fma2(pd[w], xd[w], dup(cL[step]));
                else if (step <= 6) pd[w] = fma2(pd[w], xd[w], dup(cS[step]));
                if (sL) ps[w] = fma2(ps[w], xs[w], dup(cL[step]));
                else if (step <= 6) ps[w] = fma2(ps[w], xs[w], dup(cS[step]));
            }
        }
    }
#else
#pragma unroll
    for (int w = 0; w < W; ++w) {
        pd[w] = (T == kLL || T == kLS) ? large2(pd[w]) : small2(pd[w]);
        ps[w] = (T == kLL || T == kSL) ? large2(ps[w]) : small2(ps[w]);
    }
#endif
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const f2 y = mul2(mul2(pd[w], ps[w]), ex22(e[w]));
        if (w == 0) {
            // a 64-pair pass only runs while more than 32 pairs remain: its first 32 are all inside
            if (W == 2) { acc_add2(accY, y); acc_fma2(accUY, u[w], y); }
            else acc_pair<0>(accY, accUY, y, u[w], left);
        } else acc_pair<32>(accY, accUY, y, u[w], left);
    }
}

// One run of uniform type T described by a run word (see family_plan in so3_math.cuh); adds the
// run's lane-partial sums of y and u*y, times the run's constant factor, to Y / UY.
// tab: shared address of the QuadTables block + 16*lane; lim32 = 32 - lane.
template <int T>
__device__ __forceinline__ void quad_run(unsigned tab, int lane, int lim32, uint32_t word, float scale,
                                         const RunConsts& k, float& Y, float& UY) {
    unsigned ta = tab + (word & 0x1ff0u);
    int left = (int)(word >> 16) - lane;             // pairs from this lane's pair to the run's end
    f2 accY = pk(0.f, 0.f), accUY = pk(0.f, 0.f);
    // passes of 64 pairs while more than 32 remain: pairs_left = left + lane > 32  <=>  left > 32 - lane
    // (not unrolled: two passes per trip save 4 instructions per pass and double the hot code -- 10.07 -> 11.40 ms;
    //  the instruction cache, not the loop overhead, is what the replay is sensitive to)
#pragma unroll 1
    for (; left > lim32; left -= 64, ta += 1024)
        quad_pass<T, 2>(ta, left, k, accY, accUY);
    if (left > lim32 - 32) quad_pass<T, 1>(ta, left, k, accY, accUY);
    float lo, hi;
    upk(accY, lo, hi); Y = fmaf(scale, lo + hi, Y);
    upk(accUY, lo, hi); UY = fmaf(scale, lo + hi, UY);
}

// node i's constants read back from the pair-interleaved table columns (edge nodes, phase 1)
__device__ __forceinline__ NodeVals table_node(const QuadTables& tb, int i) {
    const int m = i >> 1, h = i & 1;
    NodeVals n;
    const float* q;
    q = reinterpret_cast<const float*>(&tb.SSa[m]); n.u = q[h];  n.v = q[2 + h];
    q = reinterpret_cast<const float*>(&tb.LLa[m]); n.iu = q[h]; n.iv = q[2 + h];
    q = reinterpret_cast<const float*>(&tb.LSb[m]); n.Lu = q[2 + h];
    q = reinterpret_cast<const float*>(&tb.SLa[m]); n.Lv = q[2 + h];
    return n;
}

// One node evaluated by the thread that owns the sample (trapezoid end points, edge nodes): the
// scalar twin of the packed run bodies.  Deliberately NOT inlined -- phase 1 calls it 18 times per
// tile and the kernel's instruction footprint decides whether the 24 warps of a CTA, which sit in
// different phases, keep hitting the instruction cache (arguments and result travel in registers).
__device__ __noinline__ float node_value(int type, float fd, float fs, float ifd, float ifs, float k1L, float k1S, float k2,
                                         float u, float v, float iu, float iv, float Lu, float Lv) {
    FamilyDesc d;
    d.fd = fd; d.fs = fs; d.ifd = ifd; d.ifs = ifs; d.k1L = k1L; d.k1S = k1S; d.k2 = k2;
    NodeVals n;
    n.u = u; n.v = v; n.iu = iu; n.iv = iv; n.Lu = Lu; n.Lv = Lv;
    return node_typed(d, type, n);
}
__device__ __forceinline__ float node_scaled(const FamilyDesc& d, int t, const NodeVals& n) {
    return node_value(t, d.fd, d.fs, d.ifd, d.ifs, d.k1L, d.k1S, d.k2, n.u, n.v, n.iu, n.iv, n.Lu, n.Lv) * type_scale(d, t);
}
// trapezoid half weight of an end node, with the constant factor of the run it sits in
__device__ __forceinline__ float end_node(const FamilyDesc& d, int i, const NodeVals& n) {
    return node_scaled(d, node_type(d, i), n);
}

}  // namespace

// ---------------------------------------------------------------------------
// K2: fused Fisher kernel
// ---------------------------------------------------------------------------
// NFAM = 3: the full head.  NFAM = 1: forward-only calls (NLL / logC without gradient, G or entropy --
// validation under no_grad) need the normaliser family alone, a third of the quadrature.
template <int NFAM>
__global__ void __launch_bounds__(kThreads, 1)
fisher_fused_kernel(FisherArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QuadTables& tb = *reinterpret_cast<QuadTables*>(smem_raw);
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem_raw + sizeof(QuadTables));

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    // node tables (every block builds its own copy: 2 divides + 2 log2 per node, once)
    for (int i = threadIdx.x; i < 2 * kTabPairs; i += kThreads) {
        const NodeVals n = node_vals(i);
        const int m = i >> 1, h = i & 1;
        float* q;
        q = reinterpret_cast<float*>(&tb.LLa[m]); q[h] = n.iu; q[2 + h] = n.iv;
        q = reinterpret_cast<float*>(&tb.LLb[m]); q[h] = n.u;  q[2 + h] = n.Lu + n.Lv;
        q = reinterpret_cast<float*>(&tb.LSa[m]); q[h] = n.iu; q[2 + h] = n.v;
        q = reinterpret_cast<float*>(&tb.LSb[m]); q[h] = n.u;  q[2 + h] = n.Lu;
        q = reinterpret_cast<float*>(&tb.SLa[m]); q[h] = n.iv; q[2 + h] = n.Lv;
        q = reinterpret_cast<float*>(&tb.SSa[m]); q[h] = n.u;  q[2 + h] = n.v;
        if (i == 0) tb.first = n;
        if (i == kQuadNodes - 1) tb.last = n;
    }
    __syncthreads();

    // shared-window addresses, laundered through asm so the hot loop keeps them in registers
    // instead of re-deriving the window base (S2UR SR_CgaCtaId) at every use
    unsigned tab_s = (unsigned)__cvta_generic_to_shared(&tb) + 16u * lane;          // float4 columns, this lane's pair
    unsigned desc_s = (unsigned)__cvta_generic_to_shared(ws.desc);
    asm volatile("" : "+r"(tab_s), "+r"(desc_s));
    // Schedule: `full_rounds` rounds in which every warp of the grid takes one 32-sample tile, then ONE
    // closing round that spreads the remaining samples evenly over all warps (samples_per_warp < 32
    // each), so every warp finishes together whatever n is (a plain round-robin of 32-sample tiles
    // costs a whole extra round for the leftover: 8 % at 2^20 samples).  Small batches are the
    // special case full_rounds = 0.
    // (the tile's first sample index is 64-bit; it is re-derived where it is needed -- phase 1's loads and
    // phase 3's stores -- instead of being carried through the quadrature, whose loop bodies want the registers)
    auto tile_base = [&](int round_) -> long long {
        const long long warps_total = (long long)gridDim.x * kWarpsPerBlock;
        const long long gwarp = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
        return round_ == (int)p.full_rounds ? p.full_rounds * warps_total * 32 + gwarp * p.samples_per_warp
                                            : ((long long)round_ * warps_total + gwarp) * 32;
    };
    bool bad = false;

    for (int round = 0; round <= (int)p.full_rounds; ++round) {
        long long base = tile_base(round);
        const int count = round == (int)p.full_rounds ? (int)max(0LL, min((long long)p.samples_per_warp, p.n - base)) : 32;
        if (count == 0) break;
        const bool mine = lane < count;

        // ---- phase 1: load + SVD + run descriptors (thread per sample) ---------
        // Small tiles (small batches, closing rounds: at most 10 samples in this warp) spread the work of a
        // sample over three lanes, one family each -- phase 1 is a long dependent chain per lane and, with
        // one sample per warp, the whole latency of a small launch (BASELINE configs 1-2).
        const bool spread = NFAM == 3 && count <= 10;
        const int sj = spread ? lane / 3 : lane;                  // the sample this lane sets up
        const int fam0 = spread ? lane - 3 * sj : 0, fam1 = spread ? fam0 + 1 : NFAM;
        const bool work = sj < count;
        float s[3] = {0.f, 0.f, 0.f};
        float dot = 0.f;
        if (p.Sin) {
            // logC_F entry: singular values given directly (torch_norm_factor.py:92), U = V = I
            if (work) {
                const long long i = base + sj;
                s[0] = __ldg(p.Sin + 3 * i); s[1] = __ldg(p.Sin + 3 * i + 1); s[2] = __ldg(p.Sin + 3 * i + 2);
            }
        } else {
            load_tile(ws.a, p.A + base * 9, count, p.vec_ok, lane);
            if (p.Rgt) load_tile(ws.r, p.Rgt + base * 9, count, p.vec_ok, lane);
            __syncwarp();
            if (work) {
                float A[9], U[9], V[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) A[k] = ws.a[sj * 9 + k];
                if (p.Rgt) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) dot = fmaf(A[k], ws.r[sj * 9 + k], dot);
                }
                if (!proper_svd3(A, U, V, s) && (p.keep == nullptr || p.keep[base + sj] != 0)) bad = true;   // (the lanes of a spread sample repeat the SVD: no exchange needed)
                if (fam0 == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        ws.uv[(2 * i) * 32 + sj] = U[3 * i]; ws.uv[(2 * i + 1) * 32 + sj] = U[3 * i + 1];
                        ws.uv[(6 + 2 * i) * 32 + sj] = V[3 * i]; ws.uv[(7 + 2 * i) * 32 + sj] = V[3 * i + 1];
                    }
                }
            }
        }
        // trapezoid end-point corrections (weight 1/2 at nodes 0 and 511) and the parked descriptors
        // One family per trip of a ROLLED loop (the family's inputs are picked by selects): a third of
        // the code of the unrolled form -- see node_value about the instruction footprint.
        float cY0 = 0.f, cUY0 = 0.f, cN1 = 0.f, cN2 = 0.f;
        {
            const float* utab = reinterpret_cast<const float*>(tb.SSa);
            const float cut_thr = cut_threshold(s, p.cut_bits);
            const float uf = tb.first.u, ul = tb.last.u;
#pragma unroll 1
            for (int f = fam0; f < fam1; ++f) {
                const FamilyDesc d = family_of(s, f, utab, utab + 2, cut_thr);
                FamilyPlan pl;
                family_plan(d, pl);
                ws.desc[(f * 3 + 0) * 32 + sj] = make_float4(d.fd, d.fs, d.ifd, d.ifs);
                ws.desc[(f * 3 + 1) * 32 + sj] = make_float4(d.k1L, d.k1S, d.scLS, d.scSL);
                ws.desc[(f * 3 + 2) * 32 + sj] = make_float4(d.scMid, __uint_as_float(pl.word[0]), __uint_as_float(pl.word[1]), __uint_as_float(pl.word[2]));
                // this thread's edge nodes (pairs that straddle a type boundary)
                float eY = 0.f, eUY = 0.f;
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const int node = pl.edge[k];
                    if (node >= 0) {
                        const NodeVals nv = table_node(tb, node);
                        const float y = node_scaled(d, edge_type(d, k), nv);
                        eY += y;
                        eUY = fmaf(nv.u, y, eUY);
                    }
                }
                // trapezoid half weights (node 0 is only corrected where it was evaluated: a family with
                // cut > 0 skipped it), minus the edge nodes: what phase 3 subtracts from the warp's sums
                const float f0 = d.cut ? 0.f : end_node(d, 0, tb.first), l0 = end_node(d, kQuadNodes - 1, tb.last);
                const float cy = 0.5f * (f0 + l0) - eY;
                const float cuy = 0.5f * fmaf(uf, f0, ul * l0) - eUY;
                if (f == 0) { cY0 = cy; cUY0 = cuy; }
                else if (f == 1) cN1 = cy - cuy;
                else cN2 = cy - cuy;
            }
        }
        // Per-sample state that phase 3 needs rides through phase 2 in the sample's nine slots of the
        // (now consumed) A tile instead of in twelve registers: s, <A,R_gt>, and the four corrections,
        // which the sample's lane turns into the corrected sums when its warp-sums arrive.  (Every lane
        // has read its A by now: the barrier orders those reads before the slots are overwritten.)
        __syncwarp();
        {
            float* mine9 = ws.a + sj * 9;
            if (fam0 == 0) { mine9[0] = s[0]; mine9[1] = s[1]; mine9[2] = s[2]; mine9[3] = dot; mine9[4] = cY0; mine9[5] = cUY0; }
            if (fam0 <= 1 && fam1 > 1) mine9[6] = cN1;
            if (fam1 > 2) mine9[7] = cN2;
        }
        __syncwarp();

        // ---- phase 2: quadrature (warp per sample) ---------------------------
        // lane constants of the run walk, laundered so they stay in registers (otherwise every run
        // re-derives them from S2R SR_TID.X, a long-latency read at the head of its dependency chain)
        int lane_r = lane, lim32 = 32 - lane;
        asm volatile("" : "+r"(lane_r), "+r"(lim32));
#pragma unroll 1
        for (int j = 0; j < count; ++j) {
            float pY0 = 0.f, pUY0 = 0.f, pN1 = 0.f, pN2 = 0.f;
#pragma unroll 1
            for (int f = 0; f < NFAM; ++f) {
                const unsigned da = desc_s + 16u * (unsigned)(f * 96 + j);
                const float4 d0 = lds128_ordered(da);
                const float4 d1 = lds128_ordered(da + 512);
                const float4 d2 = lds128_ordered(da + 1024);
                RunConsts k;
                k.fd = d0.x; k.fs = d0.y; k.ifd = d0.z; k.ifs = d0.w; k.k1L = d1.x; k.k1S = d1.y;
                k.k2 = -(d0.y * kLog2e);
                const uint32_t w0 = __float_as_uint(d2.y), w1 = __float_as_uint(d2.z), w2 = __float_as_uint(d2.w);
                float Y = 0.f, UY = 0.f;
                if (w0 >> 16) quad_run<kLS>(tab_s, lane_r, lim32, w0, d1.z, k, Y, UY);
                if (w1 >> 16) {
                    if (w1 & 2u) quad_run<kLL>(tab_s, lane_r, lim32, w1, d2.x, k, Y, UY);
                    else         quad_run<kSS>(tab_s, lane_r, lim32, w1, d2.x, k, Y, UY);
                }
                if (w2 >> 16) quad_run<kSL>(tab_s, lane_r, lim32, w2, d1.w, k, Y, UY);
                if (f == 0) { pY0 = Y; pUY0 = UY; }
                else if (f == 1) pN1 = Y - UY;
                else pN2 = Y - UY;
            }
            {
                // Four warp sums in one butterfly: after the exchanges over lane bits 4 and 3 every lane carries
                // ONE of the four quantities (index 2*bit4 + bit3), so 6 shuffles do the work of 20.  Lanes
                // 0, 8, 16, 24 end up with the totals of F, UY, N1, N2 and subtract the sample's corrections in place.
                // (The forward-only instantiation runs the same sequence with three zeros: its F -- hence its NLL --
                // is then bit-identical to the full launch's, which a plain 5-step shuffle sum was not on the device.)
                const bool b4 = (lane_r & 16) != 0, b3 = (lane_r & 8) != 0;
                float k0 = b4 ? pN1 : pY0, k1 = b4 ? pN2 : pUY0;
                const float s0 = b4 ? pY0 : pN1, s1 = b4 ? pUY0 : pN2;
                k0 += __shfl_xor_sync(kFull, s0, 16);
                k1 += __shfl_xor_sync(kFull, s1, 16);
                float t = b3 ? k1 : k0;
                t += __shfl_xor_sync(kFull, b3 ? k0 : k1, 8);
                t += __shfl_xor_sync(kFull, t, 4);
                t += __shfl_xor_sync(kFull, t, 2);
                t += __shfl_xor_sync(kFull, t, 1);
                // the slot by its shared-window address off desc_s (a generic pointer into ws.a makes the four lanes
                // re-derive the window base -- two S2R and ten more instructions -- once per rotation)
                const unsigned slot_s = desc_s - (unsigned)(offsetof(WarpScratch, desc) - offsetof(WarpScratch, a))
                                      + 4u * (unsigned)(j * 9 + 4 + (lane_r >> 3));
                if ((lane_r & 7) == 0) {
                    float corr;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(corr) : "r"(slot_s) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" : : "r"(slot_s), "f"(t - corr) : "memory");
                }
            }
        }

        // ---- phase 3: closing arithmetic + stores ----------------------------
        __syncwarp();
        { int r_ = round; asm volatile("" : "+r"(r_)); base = tile_base(r_); }
        if (mine) {
            const float* mine9 = ws.a + lane * 9;
            s[0] = mine9[0]; s[1] = mine9[1]; s[2] = mine9[2];
            const float dot = mine9[3];
            const float F = mine9[4];
            const float N0 = F - mine9[5];
            FisherStats st = fisher_finish(s, F, N0, mine9[6], mine9[7]);
            float U[9], V[9], M[9];
            const long long i = base + lane;
            if (p.nll) p.nll[i] = fmaf(p.overreg, st.logC, -dot);
            if (p.entropy) p.entropy[i] = st.entropy;
            if (p.hist) {
                // first radix-select pass, in the kernel: the top 11 key bits of the entropy this lane just
                // produced.  Entropies of a tile fall into a handful of bins, so the lanes that share a bin
                // elect a leader and one RED.ADD per (tile, bin) goes to the 2048 global counters.
                const unsigned bin = entropy_key(st.entropy) >> kHistShift1;
                const unsigned peers = __match_any_sync(__activemask(), bin);
                if ((int)(__ffs(peers) - 1) == lane) atomicAdd(p.hist + bin, (unsigned long long)__popc(peers));
            }
            if (p.logC) p.logC[i] = st.logC;
            if (p.S) { p.S[3 * i] = s[0]; p.S[3 * i + 1] = s[1]; p.S[3 * i + 2] = s[2]; }
            if (p.G) { p.G[3 * i] = st.g[0]; p.G[3 * i + 1] = st.g[1]; p.G[3 * i + 2] = st.g[2]; }
            if (p.grad || p.Rout) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    U[3 * i] = ws.uv[(2 * i) * 32 + lane]; U[3 * i + 1] = ws.uv[(2 * i + 1) * 32 + lane];
                    V[3 * i] = ws.uv[(6 + 2 * i) * 32 + lane]; V[3 * i + 1] = ws.uv[(7 + 2 * i) * 32 + lane];
                }
                // third columns: the cross product of the first two (the same expression proper_svd3 uses for u2)
                U[2] = fmaf(U[3], U[7], -U[6] * U[4]); U[5] = fmaf(U[6], U[1], -U[0] * U[7]); U[8] = fmaf(U[0], U[4], -U[3] * U[1]);
                V[2] = fmaf(V[3], V[7], -V[6] * V[4]); V[5] = fmaf(V[6], V[1], -V[0] * V[7]); V[8] = fmaf(V[0], V[4], -V[3] * V[1]);
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
}

// ---------------------------------------------------------------------------
// K1: proper SVD only (batch_torch_A_to_R, analytical_mode, proper_svd)
// thread per sample, tile of 128 records staged through shared memory
// ---------------------------------------------------------------------------
namespace {
constexpr int kSvdWarps = 4;
constexpr int kSvdThreads = kSvdWarps * 32;
struct __align__(16) SvdScratch {
    float a[kTileFloats];      // A in
    float r[kTileFloats];      // rotation out
    float uv[18 * 32];         // U,V [k][lane]
};
}  // namespace

__global__ void __launch_bounds__(kSvdThreads)
proper_svd_kernel(SvdArgs p) {
    __shared__ SvdScratch scratch[kSvdWarps];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    SvdScratch& ws = scratch[warp];
    const long long warps_total = (long long)gridDim.x * kSvdWarps;
    const long long tiles = (p.n + 31) / 32;
    bool bad = false;
    for (long long tile = (long long)blockIdx.x * kSvdWarps + warp; tile < tiles; tile += warps_total) {
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
// Backward of K1's rotation output (batch_torch_A_to_R is differentiable in the reference through torch.svd:
// src/fisher/fisher_utils.py:39-48).  With A = U diag(s) V^T (proper: det U = det V = +1, s3 signed) the polar
// factor R = U V^T moves by  U^T dR V = X,  X_ij = (M_ij - M_ji) / (s_i + s_j),  M = U^T dA V,  so for an incoming
// G = dL/dR:   dL/dA = U Q V^T,   Q_ij = (P_ij - P_ji) / (s_i + s_j),   P = U^T G V   (Q_ii = 0).
// s_i + s_j = 0 (a reflection-degenerate spectrum, where the projection itself is not unique) contributes nothing.
// Thread per matrix; 144 B in, 36 B out per matrix: HBM-bound and tiny.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSvdThreads)
polar_backward_kernel(const float* __restrict__ U, const float* __restrict__ V, const float* __restrict__ S,
                      const float* __restrict__ G, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * kSvdThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kSvdThreads) {
        float u[9], v[9], g[9], s[3], t[9], p[9], q[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { u[k] = __ldg(U + i * 9 + k); v[k] = __ldg(V + i * 9 + k); g[k] = __ldg(G + i * 9 + k); }
        s[0] = __ldg(S + i * 3); s[1] = __ldg(S + i * 3 + 1); s[2] = __ldg(S + i * 3 + 2);
        // t = G V, p = U^T t
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                t[3 * r + c] = fmaf(g[3 * r], v[c], fmaf(g[3 * r + 1], v[3 + c], g[3 * r + 2] * v[6 + c]));
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                p[3 * r + c] = fmaf(u[r], t[c], fmaf(u[3 + r], t[3 + c], u[6 + r] * t[6 + c]));
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float den = s[r] + s[c];
                q[3 * r + c] = (r == c || den == 0.0f) ? 0.0f : div_rn(p[3 * r + c] - p[3 * c + r], den);
            }
        // out = U q V^T
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                t[3 * r + c] = fmaf(u[3 * r], q[c], fmaf(u[3 * r + 1], q[3 + c], u[3 * r + 2] * q[6 + c]));
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                out[i * 9 + 3 * r + c] = fmaf(t[3 * r], v[3 * c], fmaf(t[3 * r + 1], v[3 * c + 1], t[3 * r + 2] * v[3 * c + 2]));
    }
}

// ---------------------------------------------------------------------------
// fisher_CE closing kernel: thread per pair.  The two quadratures (target and
// prediction) are K2 launches; this step re-derives both proper SVDs (registers, same routine),
// builds the two quaternion frames and evaluates fisher_ce_close (so3_math.cuh).  HBM-bound:
// 72 B of parameters + 40 B of statistics in, 4 + 36 B out per pair.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSvdThreads)
fisher_ce_close_kernel(FisherCeArgs p, bool vec_ok) {
    __shared__ SvdScratch scratch[kSvdWarps];        // a: A1 in -> gradient out, r: A2 in
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    SvdScratch& ws = scratch[warp];
    const long long warps_total = (long long)gridDim.x * kSvdWarps;
    const long long tiles = (p.n + 31) / 32;
    int flags = 0;
    for (long long tile = (long long)blockIdx.x * kSvdWarps + warp; tile < tiles; tile += warps_total) {
        const long long base = tile * 32;
        const int count = (int)min(32LL, p.n - base);
        load_tile(ws.a, p.A1 + base * 9, count, vec_ok, lane);
        load_tile(ws.r, p.A2 + base * 9, count, vec_ok, lane);
        __syncwarp();
        float grad[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) grad[k] = 0.0f;
        const bool kept_row = lane < count && (p.keep == nullptr || p.keep[base + lane] != 0);
        if (lane < count && !kept_row) p.ce[base + lane] = 0.0f;      // filtered row: the reference never evaluates it
        if (kept_row) {
            const long long i = base + lane;
            float A[9], U1[9], V1[9], U2[9], V2[9], s1[3], s2[3];
#pragma unroll
            for (int k = 0; k < 9; ++k) A[k] = ws.a[lane * 9 + k];
            if (!proper_svd3(A, U1, V1, s1)) flags |= kStatusNonFinite;
#pragma unroll
            for (int k = 0; k < 9; ++k) A[k] = ws.r[lane * 9 + k];
            if (!proper_svd3(A, U2, V2, s2)) flags |= kStatusNonFinite;
            const float g1[3] = {__ldg(p.G1 + 3 * i), __ldg(p.G1 + 3 * i + 1), __ldg(p.G1 + 3 * i + 2)};
            const float g2[3] = {__ldg(p.G2 + 3 * i), __ldg(p.G2 + 3 * i + 1), __ldg(p.G2 + 3 * i + 2)};
            // the singular values the statistics were computed for (same routine, same input; read
            // back rather than assumed bit-identical across the two compilations)
            s2[0] = __ldg(p.S2 + 3 * i); s2[1] = __ldg(p.S2 + 3 * i + 1); s2[2] = __ldg(p.S2 + 3 * i + 2);
            // log f(s2) from K2's entropy H = log f + sum_j s_j (1 - g_j): the correction is O(1), so
            // nothing cancels (logC - sum(s) would lose the leading digits for concentrated densities)
            const float logf2 = __ldg(p.H2 + i) - fmaf(s2[0], 1.0f - g2[0], fmaf(s2[1], 1.0f - g2[1], s2[2] * (1.0f - g2[2])));
            const float ce = fisher_ce_close(U1, V1, g1, U2, V2, s2, g2, logf2, p.grad ? grad : nullptr);
            if (!(fabsf(ce) <= 3.402823466e38f)) flags |= kStatusNonFiniteCE;
            p.ce[i] = ce;
        }
        __syncwarp();
        if (p.grad) {
            if (lane < count) {
#pragma unroll
                for (int k = 0; k < 9; ++k) ws.a[lane * 9 + k] = grad[k];
            }
            __syncwarp();
            store_tile(p.grad + base * 9, ws.a, count, vec_ok, lane);
        }
        __syncwarp();
    }
    if (flags && p.status) atomicOr(p.status, flags);
}

// ---------------------------------------------------------------------------
// Rotate-augmentation adjustment of the teacher's parameter matrices (src/agent.py:110-119, SURVEY
// 8f-2): the unlabeled image was rotated in-plane for the strong view, so the weak-view prediction
// is moved into the strong view's frame before it becomes the pseudo label.
//   mode 0 (train_labeled == "DAD3DHeads"):  out = Raug P
//   mode 1 (train_labeled == "300WLP"):      out = (D Raug D P^T)^T = P D Raug^T D,  D = diag(1,-1,-1)
// Thread per matrix, tiles staged like K1.  HBM-bound, 108 B per matrix.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSvdThreads)
rotate_adjust_kernel(const float* __restrict__ P, const float* __restrict__ Raug, long long n, int mode,
                     float* __restrict__ out, bool vec_ok) {
    __shared__ SvdScratch scratch[kSvdWarps];        // a: P in -> result out, r: Raug in
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    SvdScratch& ws = scratch[warp];
    const long long warps_total = (long long)gridDim.x * kSvdWarps;
    const long long tiles = (n + 31) / 32;
    for (long long tile = (long long)blockIdx.x * kSvdWarps + warp; tile < tiles; tile += warps_total) {
        const long long base = tile * 32;
        const int count = (int)min(32LL, n - base);
        load_tile(ws.a, P + base * 9, count, vec_ok, lane);
        load_tile(ws.r, Raug + base * 9, count, vec_ok, lane);
        __syncwarp();
        float M[9];
        if (lane < count) {
            float p[9], r[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { p[k] = ws.a[lane * 9 + k]; r[k] = ws.r[lane * 9 + k]; }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (mode == 0) {
                        M[3 * i + j] = fmaf(r[3 * i], p[j], fmaf(r[3 * i + 1], p[3 + j], r[3 * i + 2] * p[6 + j]));
                    } else {
                        // (P D Raug^T D)_ij = sum_k P_ik d_k d_j Raug_jk
                        const float dj = (j == 0) ? 1.0f : -1.0f;
                        M[3 * i + j] = fmaf(p[3 * i], dj * r[3 * j], fmaf(p[3 * i + 1], -dj * r[3 * j + 1], p[3 * i + 2] * (-dj * r[3 * j + 2])));
                    }
                }
        }
        __syncwarp();
        if (lane < count) {
#pragma unroll
            for (int k = 0; k < 9; ++k) ws.a[lane * 9 + k] = M[k];
        }
        __syncwarp();
        store_tile(out + base * 9, ws.a, count, vec_ok, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Body probe (bench/profiling aid): the W2 pass body of one run type in a tight loop, tables in
// shared memory, no per-sample glue.  variant bits: 0-1 type, 2 skip LDS, 3 skip MUFU, 4 skip mask
// ---------------------------------------------------------------------------
template <int T, bool NO_LDS, bool NO_MUFU, bool NO_MASK>
__device__ __forceinline__ void probe_pass(unsigned ta, unsigned tbb, unsigned base, unsigned lenm, const RunConsts& k,
                                           f2& accY, f2& accUY, float4 ra, float4 rb) {
    constexpr int W = 2;
    float4 a[W], b[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if (NO_LDS) { a[w] = ra; b[w] = rb; }
        else {
            a[w] = lds128(ta + 512 * w);
            if (T == kLL || T == kLS) b[w] = lds128(tbb + 512 * w);
        }
    }
    f2 pd[W], ps[W], e[W], u[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if (T == kLL) {
            u[w] = pk(b[w].x, b[w].y);
            pd[w] = mul2(dup(k.ifd), pk(a[w].x, a[w].y));
            ps[w] = mul2(dup(k.ifs), pk(a[w].z, a[w].w));
            e[w] = fma2(dup(k.k1L), u[w], pk(b[w].z, b[w].w));
        } else {
            u[w] = pk(a[w].x, a[w].y);
            const f2 v = pk(a[w].z, a[w].w);
            pd[w] = mul2(dup(k.fd), u[w]);
            ps[w] = mul2(dup(k.fs), v);
            e[w] = fma2(dup(k.k2), v, mul2(dup(k.k1S), u[w]));
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        pd[w] = (T == kLL) ? large2(pd[w]) : small2(pd[w]);
        ps[w] = (T == kLL) ? large2(ps[w]) : small2(ps[w]);
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        float ylo, yhi;
        upk(mul2(mul2(pd[w], ps[w]), NO_MUFU ? e[w] : ex22(e[w])), ylo, yhi);
        if (!NO_MASK) {
            ylo = (base + 64u * w < lenm) ? ylo : 0.0f;
            yhi = (base + 64u * w + 1u < lenm) ? yhi : 0.0f;
        }
        const f2 y = pk(ylo, yhi);
        acc_add2(accY, y);
        acc_fma2(accUY, u[w], y);
    }
}

template <int T, bool NO_LDS, bool NO_MUFU, bool NO_MASK>
__global__ void __launch_bounds__(kThreads, 1) body_probe_kernel(float* sink, int iters, float seed) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QuadTables& tb = *reinterpret_cast<QuadTables*>(smem_raw);
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * kTabPairs; i += kThreads) {
        const NodeVals n = node_vals(i);
        const int m = i >> 1, h = i & 1;
        float* q;
        q = reinterpret_cast<float*>(&tb.LLa[m]); q[h] = n.iu; q[2 + h] = n.iv;
        q = reinterpret_cast<float*>(&tb.LLb[m]); q[h] = n.u;  q[2 + h] = n.Lu + n.Lv;
        q = reinterpret_cast<float*>(&tb.SSa[m]); q[h] = n.u;  q[2 + h] = n.v;
    }
    __syncthreads();
    unsigned tab_s = (unsigned)__cvta_generic_to_shared(&tb) + 16u * lane;
    asm volatile("" : "+r"(tab_s));
    RunConsts k;
    k.fd = seed; k.fs = seed * 0.5f; k.ifd = 1.0f / (seed * 40.f); k.ifs = 1.0f / (seed * 80.f); k.k1L = -seed; k.k1S = -seed * 1.5f; k.k2 = -0.7f * seed;
    f2 accY = pk(0.f, 0.f), accUY = pk(0.f, 0.f);
    const float4 ra = make_float4(1.5f * seed, 1.25f * seed, 1.1f * seed, 1.3f * seed), rb = make_float4(0.5f, 0.25f, -0.5f * seed, -0.7f);
    const unsigned offA = (T == kLL) ? offsetof(QuadTables, LLa) : offsetof(QuadTables, SSa);
    const unsigned offB = offsetof(QuadTables, LLb);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned o = ((unsigned)it & 1u) * 1024u + 64u * 16u;
        probe_pass<T, NO_LDS, NO_MUFU, NO_MASK>(tab_s + offA + o, tab_s + offB + o, 2u * lane, 128u - ((unsigned)it & 3u), k, accY, accUY, ra, rb);
    }
    float a0, a1, b0, b1;
    upk(accY, a0, a1); upk(accUY, b0, b1);
    if (a0 + a1 + b0 + b1 == 123.456f) sink[0] = a0;
}

cudaError_t launch_body_probe(float* sink, int variant, int iters, int blocks, cudaStream_t stream) {
    constexpr size_t kSmem = sizeof(QuadTables);
#define SUHPE_BP(T, A, B, C) { cudaFuncSetAttribute(body_probe_kernel<T, A, B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); \
      body_probe_kernel<T, A, B, C><<<blocks, kThreads, kSmem, stream>>>(sink, iters, 1.0f); }
    const int t = variant & 3;
    const bool nl = variant & 4, nm = variant & 8, nk = variant & 16;
    if (t == kLL) {
        if (!nl && !nm && !nk) SUHPE_BP(kLL, false, false, false)
        else if (nl && !nm && !nk) SUHPE_BP(kLL, true, false, false)
        else if (!nl && nm && !nk) SUHPE_BP(kLL, false, true, false)
        else if (!nl && !nm && nk) SUHPE_BP(kLL, false, false, true)
        else SUHPE_BP(kLL, true, true, true)
    } else {
        if (!nl && !nm && !nk) SUHPE_BP(kSS, false, false, false)
        else SUHPE_BP(kSS, true, true, true)
    }
#undef SUHPE_BP
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
// per-device launch facts, looked up once per device (the library holds no other state)
int device_sm_count() {
    constexpr int kMaxDevices = 64;
    static int sms[kMaxDevices] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    int v = sms[dev];
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;                         // benign race: every writer stores the same value
    }
    return v;
}

static int sm_count() { return device_sm_count(); }

cudaError_t launch_fisher_fused(FisherArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int sms = sm_count();
    constexpr size_t kSmem = sizeof(QuadTables) + sizeof(WarpScratch) * kWarpsPerBlock;
    static_assert(kSmem <= 227 * 1024, "fisher_fused_kernel shared memory exceeds one SM");
    const bool forward_only = !(p.grad || p.entropy || p.G || p.hist);
    auto kernel = forward_only ? fisher_fused_kernel<1> : fisher_fused_kernel<3>;
    static unsigned long long attr_done[2] = {0ull, 0ull};
    cudaError_t err = allow_dynamic_smem(kernel, kSmem, attr_done[forward_only ? 0 : 1]);
    if (err != cudaSuccess) return err;
    // one persistent CTA of 24 warps per SM (the node tables are shared by the whole CTA); small
    // batches get one sample per warp on as many SMs as that takes (latency)
    long long blocks = (p.n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > sms) blocks = sms;
    const long long warps_total = blocks * kWarpsPerBlock;
    // full rounds of 32-sample tiles (float4-coalesced tile I/O), then the leftover spread evenly
    p.full_rounds = p.n / (warps_total * 32);
    const long long rest = p.n - p.full_rounds * warps_total * 32;
    const long long spw = (rest + warps_total - 1) / warps_total;    // 0..32
    p.samples_per_warp = (int)spw;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.vec_ok = aligned(p.A) && aligned(p.Rgt) && aligned(p.grad) && aligned(p.Rout);
    kernel<<<(unsigned)blocks, kThreads, kSmem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_rotate_adjust(const float* P, const float* Raug, long long n, int mode, float* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int sms = sm_count();
    const long long tiles = (n + 31) / 32;
    long long blocks = (tiles + kSvdWarps - 1) / kSvdWarps;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    rotate_adjust_kernel<<<(unsigned)blocks, kSvdThreads, 0, stream>>>(P, Raug, n, mode, out, aligned(P) && aligned(Raug) && aligned(out));
    return cudaGetLastError();
}

cudaError_t launch_fisher_ce_close(FisherCeArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int sms = sm_count();
    const long long tiles = (p.n + 31) / 32;
    long long blocks = (tiles + kSvdWarps - 1) / kSvdWarps;
    const long long max_blocks = (long long)sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec_ok = aligned(p.A1) && aligned(p.A2) && aligned(p.grad);
    fisher_ce_close_kernel<<<(unsigned)blocks, kSvdThreads, 0, stream>>>(p, vec_ok);
    return cudaGetLastError();
}

cudaError_t launch_polar_backward(const float* U, const float* V, const float* S, const float* G, long long n, float* out,
                                  cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + kSvdThreads - 1) / kSvdThreads;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    polar_backward_kernel<<<(unsigned)blocks, kSvdThreads, 0, stream>>>(U, V, S, G, n, out);
    return cudaGetLastError();
}

cudaError_t launch_proper_svd(SvdArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int sms = sm_count();
    const long long tiles = (p.n + 31) / 32;
    long long blocks = (tiles + kSvdWarps - 1) / kSvdWarps;
    const long long max_blocks = (long long)sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.vec_ok = aligned(p.A) && aligned(p.R);
    proper_svd_kernel<<<(unsigned)blocks, kSvdThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace suhpe
