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
//
// Layout of the work: every warp is its own two-stage pipeline.  A tile is 32 consecutive
// pairs = 1152 + 1152 (+ 384) contiguous, 16-byte aligned bytes; lane 0 fetches the warp's NEXT
// tile with 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on a per-warp mbarrier) while
// the 32 lanes work on the current one out of shared memory (stride-9 word reads: no bank
// conflicts).  No block-wide barrier anywhere; a thread spends ~250 instructions per pair
// (branch-free atan2, Markstein division: so3_math.cuh), so the copy engine, not instruction
// issue, sets the pace.  Ragged tail tiles and unaligned base pointers take plain loads.
#include "kernels.cuh"
#include "so3_math.cuh"

namespace suhpe {

namespace {

constexpr int kMetWarps = 8;
constexpr int kMetThreads = kMetWarps * 32;
constexpr int kMetCtasPerSm = 4;
constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) MetStage {
    float rp[288];
    float rg[288];
    float ge[96];
};
struct __align__(16) MetWarp {
    MetStage st[2];
    float eul[96];           // euler out, staged for float4 stores
    float err[96];           // abs err out
    unsigned long long bar[2];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        :: "r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void plain_tile(float* dst, const float* __restrict__ src, int floats, int lane) {
    for (int i = lane; i < floats; i += 32) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(kMetThreads, kMetCtasPerSm)
metrics_kernel(MetricsArgs p, bool vec_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    MetWarp& ws = reinterpret_cast<MetWarp*>(smem_raw)[warp];

    const long long tiles = (p.n + 31) / 32;
    const long long full_tiles = vec_ok ? p.n / 32 : 0;          // tiles the bulk path may fetch
    const long long stride = (long long)gridDim.x * kMetWarps;
    const bool want_euler = p.euler || p.abs_err || p.mae || (p.sums && p.gt_euler);
    const unsigned tx_bytes = 1152u + (p.Rg ? 1152u : 0u) + (p.gt_euler ? 384u : 0u);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    bool bad = false;

    const unsigned bar0 = smem_u32(&ws.bar[0]);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto fetch = [&](long long tile, int s) {                    // lane 0 only
        const unsigned bar = bar0 + 8u * s;
        mbar_expect_tx(bar, tx_bytes);
        bulk_g2s(smem_u32(ws.st[s].rp), p.Rp + tile * 288, 1152u, bar);
        if (p.Rg) bulk_g2s(smem_u32(ws.st[s].rg), p.Rg + tile * 288, 1152u, bar);
        if (p.gt_euler) bulk_g2s(smem_u32(ws.st[s].ge), p.gt_euler + tile * 96, 384u, bar);
    };

    long long tile = (long long)blockIdx.x * kMetWarps + warp;
    if (lane == 0 && tile < full_tiles) fetch(tile, 0);
    unsigned it = 0;
    for (; tile < tiles; tile += stride, ++it) {
        const int s = it & 1;
        const long long next = tile + stride;
        if (lane == 0 && next < full_tiles) fetch(next, s ^ 1);
        MetStage& st = ws.st[s];
        const long long base = tile * 32;
        const int count = (int)min(32LL, p.n - base);
        if (tile < full_tiles) {
            mbar_wait(bar0 + 8u * s, (it >> 1) & 1u);
        } else {
            plain_tile(st.rp, p.Rp + base * 9, count * 9, lane);
            if (p.Rg) plain_tile(st.rg, p.Rg + base * 9, count * 9, lane);
            if (p.gt_euler) plain_tile(st.ge, p.gt_euler + base * 3, count * 3, lane);
            __syncwarp();
        }
        if (lane < count) {
            float Rp[9], Rg[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) Rp[k] = st.rp[lane * 9 + k];
            const long long i = base + lane;
            if (p.Rg) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Rg[k] = st.rg[lane * 9 + k];
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
                const bool dad = p.full_range == 2;          // DAD convention: e is already in degrees
                if (dad) euler_dad_degrees(Rp, e);
                else euler_from_rotation(Rp, p.full_range != 0, e);
                ws.eul[lane * 3] = e[0]; ws.eul[lane * 3 + 1] = e[1]; ws.eul[lane * 3 + 2] = e[2];
                if (p.gt_euler) {
                    float d[3], sum = 0.0f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        d[k] = fabsf((dad ? e[k] : rad_to_deg_ref(e[k])) - st.ge[lane * 3 + k]);
                        sum = add_rn(sum, d[k]);
                        acc[2 + k] += (double)d[k];
                    }
                    const float m = div_by_const(sum, 3.0f, (float)(1.0 / 3.0));      // == euler_mae_degrees
                    if (p.mae) p.mae[i] = m;
                    acc[5] += (double)m;
                    ws.err[lane * 3] = d[0]; ws.err[lane * 3 + 1] = d[1]; ws.err[lane * 3 + 2] = d[2];
                }
            }
        }
        __syncwarp();                      // stage s fully read (it is refilled next iteration); eul/err complete
        if (p.euler || (p.abs_err && p.gt_euler)) {
            if (vec_ok && count == 32) {
                if (lane < 24) {
                    if (p.euler) reinterpret_cast<float4*>(p.euler + base * 3)[lane] = reinterpret_cast<const float4*>(ws.eul)[lane];
                    if (p.abs_err && p.gt_euler) reinterpret_cast<float4*>(p.abs_err + base * 3)[lane] = reinterpret_cast<const float4*>(ws.err)[lane];
                }
            } else {
                for (int j = lane; j < count * 3; j += 32) {
                    if (p.euler) p.euler[base * 3 + j] = ws.eul[j];
                    if (p.abs_err && p.gt_euler) p.abs_err[base * 3 + j] = ws.err[j];
                }
            }
            __syncwarp();
        }
    }
    if (bad && p.status) atomicOr(p.status, kStatusTraceRange);
    if (p.sums) {
        // warp sums, then one atomic per warp and quantity (148 x 32 warps: negligible)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = acc[k];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
            if (lane == 0 && v != 0.0) atomicAdd(p.sums + k, v);
        }
    }
}

}  // namespace

cudaError_t launch_metrics(MetricsArgs p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int sms = device_sm_count();
    constexpr size_t kSmem = sizeof(MetWarp) * kMetWarps;
    static_assert(kSmem * kMetCtasPerSm <= 227 * 1024, "metrics_kernel: shared memory of the resident CTAs exceeds one SM");
    static unsigned long long attr_done = 0ull;
    cudaError_t err = allow_dynamic_smem(metrics_kernel, kSmem, attr_done);
    if (err != cudaSuccess) return err;
    const long long tiles = (p.n + 31) / 32;
    long long blocks = (tiles + kMetWarps - 1) / kMetWarps;
    const long long cap = (long long)sms * kMetCtasPerSm;        // persistent: every warp streams its share of tiles
    if (blocks > cap) blocks = cap;
    auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec_ok = al(p.Rp) && al(p.Rg) && al(p.gt_euler) && al(p.euler) && al(p.abs_err);
    metrics_kernel<<<(unsigned)blocks, kMetThreads, kSmem, stream>>>(p, vec_ok);
    return cudaGetLastError();
}

}  // namespace suhpe
