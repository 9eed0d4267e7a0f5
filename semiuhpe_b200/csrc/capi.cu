// capi.cu -- the extern "C" boundary declared in include/semiuhpe_b200.h.
// No torch types, no exceptions, no hidden synchronisation (except the _host
// pipeline and suhpe_select_read, which return host values).
#include "../../include/semiuhpe_b200.h"
#include "../../include/semiuhpe_b200_probe.h"
#include "kernels.cuh"

#include <new>
#include <stdio.h>
#include <string.h>

using namespace suhpe;

namespace {

static_assert(sizeof(SelectState) == SUHPE_SELECT_STATE_BYTES, "SelectState layout is part of the ABI");
static_assert(kHistBinsMax == SUHPE_HIST_BINS, "histogram width is part of the ABI");
static_assert(kStatusNonFinite == SUHPE_STATUS_NONFINITE && kStatusTraceRange == SUHPE_STATUS_TRACE_RANGE &&
              kStatusNonFiniteCE == SUHPE_STATUS_NONFINITE_CE, "status bits");

inline int clamp_cut_bits(int32_t bits) { return bits < 0 ? 0 : (bits > 60 ? 60 : (int)bits); }

inline int rc(cudaError_t e) { return e == cudaSuccess ? 0 : -(int)e; }
inline cudaStream_t st(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- FP32 pipe probe ---------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* sink, int iters) {
    const float seed = 1.0f + 1e-7f * (float)(threadIdx.x + blockIdx.x);
    if (VARIANT >= 10 && VARIANT <= 13) {
        // issue-mix probes: 8 packed chains; what does one instruction of another pipe cost next to the FFMA2 stream?
        // 10: every FFMA2 followed by a select (FSEL, ALU pipe)         11: every FFMA2 followed by FMNMX (ALU pipe)
        // 12: one LDS.128 per 8 FFMA2 (K2's table loads: 4 per 46)      13: one MUFU.EX2 per 8 FFMA2 (K2: 4 per 46)
        __shared__ float4 tab[256];
        tab[threadIdx.x] = make_float4(seed, seed + 1.f, seed + 2.f, seed + 3.f);
        __syncthreads();
        unsigned long long x[8], a, b;
        float z[8];
        const float af = 0.9999f, bf = 1e-4f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af));
        asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float v = seed + (float)c; asm("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v)); z[c] = v; }
        const float thr = 1.0f + 1e-7f * (float)blockIdx.x;
        unsigned sink_bits = threadIdx.x * 2654435761u;
        unsigned sa = (unsigned)__cvta_generic_to_shared(tab) + 16u * (threadIdx.x & 31);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(a), "l"(b));
                    if (VARIANT == 10) asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, %2; selp.f32 %0, %0, %1, p; }" : "+f"(z[c]) : "f"(thr), "f"(z[(c + 1) & 7]));
                    if (VARIANT == 11) asm volatile("min.f32 %0, %0, %1;" : "+f"(z[c]) : "f"(thr));
                }
                if (VARIANT == 12) {
                    unsigned v0, v1, v2, v3;                  // consumed on the ALU pipe (LOP3), not by an FMA
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(sa + 512u * (unsigned)((r + i) & 7)));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(sink_bits) : "r"(v0 ^ v2), "r"(v1 ^ v3));
                }
                if (VARIANT == 13) {
                    float y;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float((sink_bits & 0x007fffffu) | 0x3f800000u)));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(sink_bits) : "r"(__float_as_uint(y)));
                }
            }
        }
        z[0] += __uint_as_float(sink_bits & 0x3fffffffu);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c])); acc += lo + hi + z[c]; }
        if (acc == 123.456f) sink[0] = acc;
    } else if (VARIANT == 8 || VARIANT == 9) {
        // operand-bandwidth probes: every packed FMA reads DISTINCT 64-bit registers, none shared between consecutive
        // instructions (no operand-reuse-cache hits).  8: x[c] = y[c] * z[c] + x[c] (three register pairs, the form of
        // K2L's sums);  9: x[c] = x[c] * y[c] + immediate (two register pairs, K2's Horner step)
        unsigned long long x[8], y[8], z[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float v = seed + (float)c;
            asm("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v));
            // run-time values, different in the two halves: neither an immediate nor a broadcast scalar operand
            const float t = 1e-7f * (float)(threadIdx.x + 1);
            const float y0 = 0.9999f + t + 1e-6f * (float)c, y1 = 0.9998f - t + 1e-6f * (float)c;
            const float z0 = 1.0001f - t - 1e-6f * (float)c, z1 = 1.0002f + t - 1e-6f * (float)c;
            asm("mov.b64 %0, {%1, %2};" : "=l"(y[c]) : "f"(y0), "f"(y1));
            asm("mov.b64 %0, {%1, %2};" : "=l"(z[c]) : "f"(z0), "f"(z1));
        }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (VARIANT == 8) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x[c]) : "l"(y[c]), "l"(z[c]));
                    else asm volatile("{ .reg .b64 k; mov.b64 k, {0f38D1B717, 0f38D1B717}; fma.rn.f32x2 %0, %0, %1, k; }" : "+l"(x[c]) : "l"(y[c]));
                }
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c])); acc += lo + hi; }
        if (acc == 123.456f) sink[0] = acc;
    } else if (VARIANT == 6 || VARIANT == 7) {
        // packed FMA whose addend is an immediate (6) or whose multiplier is a broadcast scalar register (7):
        // the operand forms the Horner bodies of K2 use
        unsigned long long x[8], a;
        const float af = 0.9999f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af));
        float sc = 0.9999f + 1e-9f * (float)blockIdx.x;
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float v = seed + (float)c; asm("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v)); }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (VARIANT == 6) {
                        asm volatile("{ .reg .b64 k; mov.b64 k, {0f38D1B717, 0f38D1B717}; fma.rn.f32x2 %0, %0, %1, k; }" : "+l"(x[c]) : "l"(a));
                    } else {
                        asm volatile("{ .reg .b64 k; mov.b64 k, {%1, %1}; fma.rn.f32x2 %0, %0, k, %2; }" : "+l"(x[c]) : "f"(sc), "l"(a));
                    }
                }
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c])); acc += lo + hi; }
        if (acc == 123.456f) sink[0] = acc;
    } else if (VARIANT == 4 || VARIANT == 5) {
        // 8 packed chains, each FFMA2 followed by one ALU (4: LOP3) or one MUFU-free FSEL-like (5: IADD) op:
        // does a 2-cycle FFMA2 leave an issue slot for another pipe?
        unsigned long long x[8], a, b;
        unsigned int z[8];
        const float af = 0.9999f, bf = 1e-4f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af));
        asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float v = seed + (float)c; asm("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v)); z[c] = threadIdx.x + c; }
        const unsigned int m = blockIdx.x | 0x55u;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(a), "l"(b));
                    if (VARIANT == 4) asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[c]) : "r"(m));
                    else asm volatile("add.u32 %0, %0, %1;" : "+r"(z[c]) : "r"(m));
                }
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c])); acc += lo + hi + (float)z[c]; }
        if (acc == 123.456f) sink[0] = acc;
    } else if (VARIANT == 3) {
        // 8 packed chains and 8 scalar chains interleaved: does the scalar FFMA find a free sub-pipe
        // while FFMA2 occupies the other?
        unsigned long long x[8], a, b;
        float z[8];
        const float af = 0.9999f, bf = 1e-4f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af));
        asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float v = seed + (float)c; asm("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v)); z[c] = v; }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(a), "l"(b));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(z[c]) : "f"(af), "f"(bf));
                }
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c])); acc += lo + hi + z[c]; }
        if (acc == 123.456f) sink[0] = acc;
    } else if (VARIANT == 1) {
        // 8 independent packed chains: x = x * a + b on (lo,hi) pairs
        unsigned long long x[8], a, b;
        const float af = 0.9999f, bf = 1e-4f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af));
        asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float v = seed + (float)c; asm("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v)); }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(a), "l"(b));
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c])); acc += lo + hi; }
        if (acc == 123.456f) sink[0] = acc;
    } else {
        float x[8];
        const float a = 0.9999f, b = 1e-4f;
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = seed + (float)c;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = fmaf(x[c], a, b);
                if (VARIANT == 2) {
                    // one MUFU.EX2 per 8 FMA, on its own dependency chain slot
                    float y;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[r & 7]));
                    x[(r + 1) & 7] += y * 1e-30f;
                }
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += x[c];
        if (acc == 123.456f) sink[0] = acc;
    }
}

}  // namespace

namespace suhpe {
cudaError_t launch_fp32_probe(float* sink, int variant, int iters, int blocks, cudaStream_t stream) {
    if (variant == 0) fp32_probe_kernel<0><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 1) fp32_probe_kernel<1><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 2) fp32_probe_kernel<2><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 3) fp32_probe_kernel<3><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 4) fp32_probe_kernel<4><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 5) fp32_probe_kernel<5><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 6) fp32_probe_kernel<6><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 7) fp32_probe_kernel<7><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 8) fp32_probe_kernel<8><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 9) fp32_probe_kernel<9><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 10) fp32_probe_kernel<10><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 11) fp32_probe_kernel<11><<<blocks, 256, 0, stream>>>(sink, iters);
    else if (variant == 12) fp32_probe_kernel<12><<<blocks, 256, 0, stream>>>(sink, iters);
    else fp32_probe_kernel<13><<<blocks, 256, 0, stream>>>(sink, iters);
    return cudaGetLastError();
}
}  // namespace suhpe

// ---- host-buffer pipeline ------------------------------------------------------
struct suhpe_pipeline {
    // Three in-order queues -- host->device copies, kernels, device->host copies -- over a ring of
    // kBuffers chunk buffers, so the two copy engines and the SMs all stay busy at once.
    static constexpr int kBuffers = 4;
    long long max_n, chunk;
    int cut_bits;
    bool rec_k[kBuffers], rec_out[kBuffers];   // ev_k / ev_out of this buffer have been recorded (in any call)
    cudaStream_t s_in, s_k, s_out;
    cudaEvent_t ev_in[kBuffers], ev_k[kBuffers], ev_out[kBuffers], ev_user;
    float *dA[kBuffers], *dR[kBuffers], *dGrad[kBuffers], *dNll[kBuffers];
    float* dEnt;             // (max_n) entropies of the whole pool stay resident for the select
    uint8_t* dMask;          // (max_n)
    unsigned long long* dHist;   // 2 x SUHPE_HIST_BINS: [0] fused first pass, [1] scratch
    SelectState* dState;
    int* dStatus;
    SelectState* hState;     // pinned
    int* hStatus;            // pinned
};

// ---- SSL loss head -------------------------------------------------------------
// the handle owns only the two forked streams and their events; every byte of scratch comes from the caller
struct suhpe_ssl_step {
    cudaStream_t s_a, s_b;                  // forked branches (the caller's stream carries the teacher branch)
    cudaEvent_t ev_fork, ev_a, ev_b;
};

extern "C" {

int suhpe_abi_version(void) { return SUHPE_ABI_VERSION; }

const char* suhpe_error_string(int code) {
    if (code == 0) return "ok";
    if (code == SUHPE_EINVAL) return "invalid argument";
    if (code < 0) return cudaGetErrorString((cudaError_t)(-code));
    return "unknown";
}

int suhpe_proper_svd_f32(const float* A, int64_t n, float* R, float* S, float* U, float* V,
                         int* status, void* stream) {
    if (n < 0 || (n > 0 && !A)) return SUHPE_EINVAL;
    SvdArgs p{A, (long long)n, R, S, U, V, status, false};
    return rc(launch_proper_svd(p, st(stream)));
}

int suhpe_proper_svd_backward_f32(const float* U, const float* V, const float* S, const float* gradR, int64_t n,
                                  float* gradA, void* stream) {
    if (n < 0 || (n > 0 && (!U || !V || !S || !gradR || !gradA))) return SUHPE_EINVAL;
    return rc(launch_polar_backward(U, V, S, gradR, (long long)n, gradA, st(stream)));
}

int suhpe_fisher_fused_f32(const float* A, const float* Rgt, int64_t n, float overreg, int32_t cut_bits,
                           float* nll, float* grad, float* Rout, float* entropy, float* logC,
                           float* S, float* G, uint64_t* hist, int* status, void* stream) {
    if (n < 0 || (n > 0 && !A)) return SUHPE_EINVAL;
    FisherArgs p{};
    p.A = A; p.Rgt = Rgt; p.n = (long long)n; p.overreg = overreg;
    p.nll = nll; p.grad = grad; p.Rout = Rout; p.entropy = entropy; p.logC = logC; p.S = S; p.G = G;
    p.hist = reinterpret_cast<unsigned long long*>(hist); p.status = status;
    p.cut_bits = clamp_cut_bits(cut_bits);
    return rc(launch_fisher_fused(p, st(stream)));
}

int suhpe_fisher_from_s_f32(const float* S, int64_t n, int32_t cut_bits, float* logC, float* G, float* entropy,
                            int* status, void* stream) {
    if (n < 0 || (n > 0 && !S)) return SUHPE_EINVAL;
    FisherArgs p{};
    p.Sin = S; p.n = (long long)n; p.overreg = 1.0f;
    p.logC = logC; p.G = G; p.entropy = entropy; p.status = status;
    p.cut_bits = clamp_cut_bits(cut_bits);
    return rc(launch_fisher_fused(p, st(stream)));
}

static int fisher_ce_impl(const float* A1, const float* G1_given, const float* A2, int64_t n, int cut_bits,
                          const uint8_t* keep, float* ce, float* gradA2, float* workspace, int* status, void* stream) {
    if (n < 0 || (n > 0 && (!A1 || !A2 || !ce || !workspace))) return SUHPE_EINVAL;
    if (n == 0) return 0;
    float* G1 = workspace;                     // (n,3)
    float* S2 = workspace + 3 * n;             // (n,3)
    float* G2 = workspace + 6 * n;             // (n,3)
    float* H2 = workspace + 9 * n;             // (n)
    cudaError_t e = cudaSuccess;
    if (!G1_given) {
        FisherArgs t{};
        t.A = A1; t.n = (long long)n; t.overreg = 1.0f; t.G = G1; t.status = status; t.keep = keep; t.cut_bits = cut_bits;
        e = launch_fisher_fused(t, st(stream));
        if (e != cudaSuccess) return rc(e);
    }
    FisherArgs q{};
    q.A = A2; q.n = (long long)n; q.overreg = 1.0f; q.S = S2; q.G = G2; q.entropy = H2; q.status = status; q.keep = keep;
    q.cut_bits = cut_bits;
    e = launch_fisher_fused(q, st(stream));
    if (e != cudaSuccess) return rc(e);
    FisherCeArgs c{A1, A2, (long long)n, G1_given ? G1_given : G1, S2, G2, H2, ce, gradA2, status, keep};
    return rc(launch_fisher_ce_close(c, st(stream)));
}

int suhpe_fisher_ce_f32(const float* A1, const float* A2, int64_t n, int32_t cut_bits, const uint8_t* keep,
                        float* ce, float* gradA2, float* workspace, int* status, void* stream) {
    return fisher_ce_impl(A1, nullptr, A2, n, clamp_cut_bits(cut_bits), keep, ce, gradA2, workspace, status, stream);
}

int suhpe_fisher_ce_with_g1_f32(const float* A1, const float* G1, const float* A2, int64_t n, int32_t cut_bits,
                                const uint8_t* keep, float* ce, float* gradA2, float* workspace, int* status,
                                void* stream) {
    if (n > 0 && !G1) return SUHPE_EINVAL;
    return fisher_ce_impl(A1, G1, A2, n, clamp_cut_bits(cut_bits), keep, ce, gradA2, workspace, status, stream);
}

int suhpe_scale_rows_f32(const float* in, int64_t n, int32_t width, const float* row_weight,
                         const float* scalar_weight, const uint8_t* keep, float* out, void* stream) {
    if (n < 0 || width <= 0 || (n > 0 && (!in || !out))) return SUHPE_EINVAL;
    return rc(launch_scale_rows(in, (long long)n, (int)width, row_weight, scalar_weight, keep, out, st(stream)));
}

int suhpe_rotate_adjust_f32(const float* pred, const float* aug_rot, int64_t n, int32_t mode, float* out, void* stream) {
    if (n < 0 || mode < 0 || mode > 1 || (n > 0 && (!pred || !aug_rot || !out))) return SUHPE_EINVAL;
    return rc(launch_rotate_adjust(pred, aug_rot, (long long)n, (int)mode, out, st(stream)));
}

int suhpe_ema_update_f32(float* const* ema, const float* const* src, const int64_t* numel, int32_t count,
                         float alpha, float one_minus_alpha, int32_t mode, void* stream) {
    if (count < 0 || mode < 0 || mode > 1 || (count > 0 && (!ema || !src || !numel))) return SUHPE_EINVAL;
    for (int i = 0; i < count; ++i)
        if (numel[i] < 0 || (numel[i] > 0 && (!ema[i] || !src[i]))) return SUHPE_EINVAL;
    static_assert(sizeof(long long) == sizeof(int64_t), "numel table layout");
    return rc(launch_ema_update(ema, src, reinterpret_cast<const long long*>(numel), (int)count, alpha, one_minus_alpha,
                                (int)mode, st(stream)));
}

int suhpe_laplace_nll_f32(const float* A, const float* Rgt, int64_t n, const float* grid, int32_t N,
                          float* nll, float* grad, float* mode, float* logF, int* status, void* stream) {
    if (n < 0 || N <= 0 || (n > 0 && (!A || !Rgt || !grid || !nll))) return SUHPE_EINVAL;
    LaplaceArgs p{A, Rgt, (long long)n, grid, (int)N, nll, grad, mode, logF, status};
    return rc(launch_laplace(p, st(stream)));
}

int suhpe_select_init(void* state, uint64_t k, void* stream) {
    if (!state) return SUHPE_EINVAL;
    return rc(launch_select_init(static_cast<SelectState*>(state), k, st(stream)));
}

int suhpe_select_hist_f32(const float* entropy, int64_t n, int32_t pass, const void* state,
                          uint64_t* hist, void* stream) {
    if (n < 0 || pass < 1 || pass > 3 || !state || !hist || (n > 0 && !entropy)) return SUHPE_EINVAL;
    return rc(launch_select_hist(entropy, (long long)n, pass, static_cast<const SelectState*>(state),
                                 reinterpret_cast<unsigned long long*>(hist), st(stream)));
}

int suhpe_select_scan(const uint64_t* hist_parts, int32_t parts, int32_t pass, void* state, void* stream) {
    if (!hist_parts || parts < 1 || pass < 1 || pass > 3 || !state) return SUHPE_EINVAL;
    return rc(launch_select_scan(reinterpret_cast<const unsigned long long*>(hist_parts), parts, pass,
                                 static_cast<SelectState*>(state), st(stream)));
}

int suhpe_entropy_threshold_f32(const float* entropy, int64_t n, uint64_t k, void* state,
                                uint64_t* hist_scratch, const uint64_t* first_pass_hist, void* stream) {
    if (n <= 0 || !entropy || !state || !hist_scratch || k >= (uint64_t)n) return SUHPE_EINVAL;
    SelectState* s = static_cast<SelectState*>(state);
    unsigned long long* h = reinterpret_cast<unsigned long long*>(hist_scratch);
    cudaError_t e = launch_select_init(s, k, st(stream));
    for (int pass = 1; pass <= 3 && e == cudaSuccess; ++pass) {
        if (pass == 1 && first_pass_hist) {
            e = launch_select_scan(reinterpret_cast<const unsigned long long*>(first_pass_hist), 1, 1, s, st(stream));
            continue;
        }
        e = launch_select_hist(entropy, (long long)n, pass, s, h, st(stream));
        if (e == cudaSuccess) e = launch_select_scan(h, 1, pass, s, st(stream));
    }
    return rc(e);
}

const float* suhpe_select_threshold_ptr(const void* state) {
    return state ? &static_cast<const SelectState*>(state)->threshold : nullptr;
}

int suhpe_select_read(const void* state, float* threshold, uint32_t* key, uint64_t* kept, void* stream) {
    if (!state) return SUHPE_EINVAL;
    SelectState h;
    cudaError_t e = cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(st(stream));
    if (e != cudaSuccess) return rc(e);
    if (threshold) *threshold = h.threshold;
    if (key) *key = h.threshold_key;
    if (kept) *kept = h.kept;
    return 0;
}

int suhpe_entropy_mask_f32(const float* entropy, int64_t n, const float* thr_dev, float thr_host,
                           uint8_t* mask, uint64_t* kept, void* stream) {
    if (n < 0 || (n > 0 && !entropy)) return SUHPE_EINVAL;
    return rc(launch_mask(entropy, (long long)n, thr_dev, thr_host, mask,
                          reinterpret_cast<unsigned long long*>(kept), st(stream)));
}

int suhpe_so3_metrics_f32(const float* Rp, const float* Rg, const float* gt_euler_deg, int64_t n,
                          int32_t full_range, float* geo_deg, float* frob, float* euler,
                          float* abs_err, float* mae, double* sums, int* status, void* stream) {
    if (n < 0 || (n > 0 && !Rp)) return SUHPE_EINVAL;
    if (!Rg && (geo_deg || frob)) return SUHPE_EINVAL;
    if (!gt_euler_deg && (abs_err || mae)) return SUHPE_EINVAL;
    MetricsArgs p{Rp, Rg, gt_euler_deg, (long long)n, (int)full_range, geo_deg, frob, euler, abs_err, mae, sums, status};
    return rc(launch_metrics(p, st(stream)));
}

int suhpe_fp32_probe(float* sink, int32_t variant, int32_t iters, int32_t blocks, void* stream) {
    if (!sink || iters < 1 || blocks < 1 || variant < 0) return SUHPE_EINVAL;
    if (variant >= 100) return rc(launch_body_probe(sink, variant - 100, iters, blocks, st(stream)));
    if (variant > 13) return SUHPE_EINVAL;
    return rc(launch_fp32_probe(sink, variant, iters, blocks, st(stream)));
}

// ---- pipeline ---------------------------------------------------------------------
int suhpe_pipeline_destroy(suhpe_pipeline* p) {
    if (!p) return 0;
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_k) cudaStreamDestroy(p->s_k);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->ev_user) cudaEventDestroy(p->ev_user);
    for (int i = 0; i < suhpe_pipeline::kBuffers; ++i) {
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_k[i]) cudaEventDestroy(p->ev_k[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
        cudaFree(p->dA[i]); cudaFree(p->dR[i]); cudaFree(p->dGrad[i]); cudaFree(p->dNll[i]);
    }
    cudaFree(p->dEnt); cudaFree(p->dMask); cudaFree(p->dHist); cudaFree(p->dState); cudaFree(p->dStatus);
    if (p->hState) cudaFreeHost(p->hState);
    if (p->hStatus) cudaFreeHost(p->hStatus);
    delete p;
    return 0;
}

int suhpe_pipeline_create(suhpe_pipeline** out, int64_t max_n, int64_t chunk, int32_t cut_bits) {
    if (!out || max_n <= 0 || chunk <= 0) return SUHPE_EINVAL;
    suhpe_pipeline* p = new (std::nothrow) suhpe_pipeline();
    if (!p) return SUHPE_EINVAL;
    memset(p, 0, sizeof(*p));
    if (chunk > max_n) chunk = max_n;
    chunk = (chunk + 127) & ~127LL;               // chunk bases (and quarter chunks) stay 16-byte aligned: 32 records = 1152 B
    p->max_n = max_n; p->chunk = chunk; p->cut_bits = clamp_cut_bits(cut_bits);
    cudaError_t e = cudaSuccess;
    auto A = [&](void** q, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(q, bytes); };
    auto S = [&](cudaStream_t* s) { if (e == cudaSuccess) e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); };
    auto E = [&](cudaEvent_t* v) { if (e == cudaSuccess) e = cudaEventCreateWithFlags(v, cudaEventDisableTiming); };
    S(&p->s_in); S(&p->s_k); S(&p->s_out); E(&p->ev_user);
    for (int i = 0; i < suhpe_pipeline::kBuffers && e == cudaSuccess; ++i) {
        E(&p->ev_in[i]); E(&p->ev_k[i]); E(&p->ev_out[i]);
        A((void**)&p->dA[i], (size_t)chunk * 36); A((void**)&p->dR[i], (size_t)chunk * 36);
        A((void**)&p->dGrad[i], (size_t)chunk * 36); A((void**)&p->dNll[i], (size_t)chunk * 4);
    }
    A((void**)&p->dEnt, (size_t)max_n * 4); A((void**)&p->dMask, (size_t)max_n);
    A((void**)&p->dHist, sizeof(unsigned long long) * 2 * SUHPE_HIST_BINS);
    A((void**)&p->dState, sizeof(SelectState)); A((void**)&p->dStatus, sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->hState, sizeof(SelectState));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->hStatus, sizeof(int));
    if (e != cudaSuccess) { suhpe_pipeline_destroy(p); return rc(e); }
    *out = p;
    return 0;
}

// Phase A of the host pipeline: stream the pool through K2.  Entropies land in ent_dev (n),
// the first radix histogram is ADDED into hist_dev (2048 counters the caller zeroed) and the
// status bits are OR-ed into status_dev -- all three are caller-owned device buffers, ordered
// against `stream`: the first kernel waits for work already queued on `stream`, and on return
// `stream` waits for the last kernel, so a select (single-GPU, or the all-gather form over
// NCCL) can be queued on it straight away while the device->host copies still drain.
int suhpe_fisher_pool_host(suhpe_pipeline* p, const float* A_host, const float* Rgt_host, int64_t n,
                           float overreg, float* nll_host, float* grad_host, float* entropy_host,
                           float* ent_dev, uint64_t* hist_dev, int* status_dev, void* stream) {
    if (!p || !A_host || !ent_dev || n <= 0 || n > p->max_n) return SUHPE_EINVAL;
    constexpr int NB = suhpe_pipeline::kBuffers;
    cudaStream_t user = st(stream);
    const bool external = user != p->s_k;
    cudaError_t e = cudaSuccess;
#define CK(x) do { if (e == cudaSuccess) e = (x); } while (0)
    if (external) {
        CK(cudaEventRecord(p->ev_user, user));
        CK(cudaStreamWaitEvent(p->s_k, p->ev_user, 0));
    }
    // Chunk schedule: the pipeline's fill (first H2D copy) and drain (last D2H copy) are exposed
    // latency, so the pool starts and ends with quarter and half chunks and runs full chunks between.
    long long sched[8];
    int ramp = 0;
    const long long q = p->chunk / 4;                 // chunk is a multiple of 128: q keeps 16-byte alignment
    if (n >= 4 * p->chunk && q > 0) { sched[0] = q; sched[1] = 2 * q; ramp = 2; }
    const long long body = n - (ramp ? 6 * q : 0);    // head q + 2q, tail 2q + q
    const long long nbody = (body + p->chunk - 1) / p->chunk;
    const long long nchunks = nbody + 2 * ramp;
    long long base = 0;
    for (long long c = 0; c < nchunks && e == cudaSuccess; ++c) {
        const int b = (int)(c % NB);
        long long cnt;
        if (c < ramp) cnt = sched[c];
        else if (c < ramp + nbody) { const long long done = (c - ramp) * p->chunk; cnt = (body - done < p->chunk) ? (body - done) : p->chunk; }
        else cnt = (c == nchunks - 1) ? q : 2 * q;
        const long long this_base = base;
        base += cnt;
        // inputs: the buffer is free once the kernel of its previous use -- in this call or an earlier one
        // that was not followed by suhpe_pipeline_sync -- has run
        if (p->rec_k[b]) CK(cudaStreamWaitEvent(p->s_in, p->ev_k[b], 0));
        CK(cudaMemcpyAsync(p->dA[b], A_host + this_base * 9, (size_t)cnt * 36, cudaMemcpyHostToDevice, p->s_in));
        if (Rgt_host) CK(cudaMemcpyAsync(p->dR[b], Rgt_host + this_base * 9, (size_t)cnt * 36, cudaMemcpyHostToDevice, p->s_in));
        CK(cudaEventRecord(p->ev_in[b], p->s_in));
        // kernel: after its inputs landed and the previous outputs of this buffer were drained
        CK(cudaStreamWaitEvent(p->s_k, p->ev_in[b], 0));
        if (p->rec_out[b]) CK(cudaStreamWaitEvent(p->s_k, p->ev_out[b], 0));
        FisherArgs a{};
        a.A = p->dA[b]; a.Rgt = Rgt_host ? p->dR[b] : nullptr; a.n = cnt; a.overreg = overreg;
        a.nll = nll_host ? p->dNll[b] : nullptr;
        a.grad = grad_host ? p->dGrad[b] : nullptr;
        a.entropy = ent_dev + this_base;
        a.hist = reinterpret_cast<unsigned long long*>(hist_dev); a.status = status_dev; a.cut_bits = p->cut_bits;
        CK(launch_fisher_fused(a, p->s_k));
        CK(cudaEventRecord(p->ev_k[b], p->s_k));
        if (e == cudaSuccess) p->rec_k[b] = true;
        // outputs
        CK(cudaStreamWaitEvent(p->s_out, p->ev_k[b], 0));
        if (nll_host) CK(cudaMemcpyAsync(nll_host + this_base, p->dNll[b], (size_t)cnt * 4, cudaMemcpyDeviceToHost, p->s_out));
        if (grad_host) CK(cudaMemcpyAsync(grad_host + this_base * 9, p->dGrad[b], (size_t)cnt * 36, cudaMemcpyDeviceToHost, p->s_out));
        if (entropy_host) CK(cudaMemcpyAsync(entropy_host + this_base, ent_dev + this_base, (size_t)cnt * 4, cudaMemcpyDeviceToHost, p->s_out));
        CK(cudaEventRecord(p->ev_out[b], p->s_out));
        if (e == cudaSuccess) p->rec_out[b] = true;
    }
    if (external) CK(cudaStreamWaitEvent(user, p->ev_k[(nchunks - 1) % NB], 0));
#undef CK
    return rc(e);
}

// Blocks until every copy and kernel the pipeline queued has finished.
int suhpe_pipeline_sync(suhpe_pipeline* p) {
    if (!p) return SUHPE_EINVAL;
    cudaError_t e = cudaStreamSynchronize(p->s_k);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_in);
    return rc(e);
}

int suhpe_fisher_filter_host(suhpe_pipeline* p, const float* A_host, const float* Rgt_host, int64_t n,
                             float overreg, uint64_t k, float* nll_host, float* grad_host,
                             float* entropy_host, uint8_t* mask_host, float* threshold, uint64_t* kept) {
    if (!p || !A_host || n <= 0 || n > p->max_n || k >= (uint64_t)n) return SUHPE_EINVAL;
    cudaError_t e = cudaSuccess;
#define CK(x) do { if (e == cudaSuccess) e = (x); } while (0)
    CK(cudaMemsetAsync(p->dHist, 0, sizeof(unsigned long long) * 2 * SUHPE_HIST_BINS, p->s_k));
    CK(cudaMemsetAsync(p->dStatus, 0, sizeof(int), p->s_k));
    if (e != cudaSuccess) return rc(e);
    int r = suhpe_fisher_pool_host(p, A_host, Rgt_host, n, overreg, nll_host, grad_host, entropy_host,
                                   p->dEnt, reinterpret_cast<uint64_t*>(p->dHist), p->dStatus, p->s_k);
    if (r != 0) return r;
    // select + mask on the resident entropies, in order behind the last kernel
    r = suhpe_entropy_threshold_f32(p->dEnt, n, k, p->dState,
                                    reinterpret_cast<uint64_t*>(p->dHist + SUHPE_HIST_BINS),
                                    reinterpret_cast<const uint64_t*>(p->dHist), p->s_k);
    if (r != 0) return r;
    CK(launch_mask(p->dEnt, n, &p->dState->threshold, 0.f, mask_host ? p->dMask : nullptr, &p->dState->kept, p->s_k));
    if (mask_host) CK(cudaMemcpyAsync(mask_host, p->dMask, (size_t)n, cudaMemcpyDeviceToHost, p->s_k));
    CK(cudaMemcpyAsync(p->hState, p->dState, sizeof(SelectState), cudaMemcpyDeviceToHost, p->s_k));
    CK(cudaMemcpyAsync(p->hStatus, p->dStatus, sizeof(int), cudaMemcpyDeviceToHost, p->s_k));
#undef CK
    if (e != cudaSuccess) return rc(e);
    r = suhpe_pipeline_sync(p);
    if (r != 0) return r;
    if (threshold) *threshold = p->hState->threshold;
    if (kept) *kept = p->hState->kept;
    return (*p->hStatus & SUHPE_STATUS_NONFINITE) ? 1 : 0;   // >0: completed, non-finite input seen
}

// ---- SSL loss head -----------------------------------------------------------------
int suhpe_ssl_step_destroy(suhpe_ssl_step* c) {
    if (!c) return 0;
    if (c->s_a) cudaStreamDestroy(c->s_a);
    if (c->s_b) cudaStreamDestroy(c->s_b);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    delete c;
    return 0;
}

int suhpe_ssl_step_create(suhpe_ssl_step** out) {
    if (!out) return SUHPE_EINVAL;
    suhpe_ssl_step* c = new (std::nothrow) suhpe_ssl_step();
    if (!c) return SUHPE_EINVAL;
    memset(c, 0, sizeof(*c));
    cudaError_t e = cudaStreamCreateWithFlags(&c->s_a, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_b, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming);
    if (e != cudaSuccess) { suhpe_ssl_step_destroy(c); return rc(e); }
    *out = c;
    return 0;
}

int64_t suhpe_ssl_step_workspace_floats(int64_t b_l, int64_t b_u) {
    if (b_l < 0 || b_u < 0) return 0;
    // kept counter | nll_l | entropy | fisher_CE workspace (G1 S2 G2 H2) | adjusted | pseudo | loss_u | mask bytes,
    // every region rounded up to a multiple of 4 floats (the carving in suhpe_ssl_step_f32)
    auto up4 = [](int64_t v) { return (v + 3) & ~(int64_t)3; };
    return 4 + up4(b_l) + up4(b_u) + up4(SUHPE_FISHER_CE_WORKSPACE_FLOATS * b_u) + up4(9 * b_u) + up4(9 * b_u) + up4(b_u) +
           up4((b_u + 3) / 4);
}

int suhpe_ssl_step_f32(suhpe_ssl_step* c, const float* out_l, const float* gt_l, int64_t b_l,
                       const float* pred_weak, const float* pred_strong, int64_t b_u,
                       const float* aug_rot, int32_t aug_mode, const float* conf_thres_dev, float conf_thres_host,
                       float overreg, float ssl_lambda, int32_t unsup_kind, int32_t cut_bits,
                       float* workspace,
                       float* losses, float* grad_l, float* grad_strong,
                       float* Rest_l, float* entropy, uint8_t* mask, float* pseudo, float* losses_l, float* losses_u,
                       int* status, void* stream) {
    if (!c || !losses || !workspace || b_l <= 0 || b_u < 0 || !out_l || !gt_l) return SUHPE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return SUHPE_EINVAL;
    if (b_u > 0 && (!pred_weak || !pred_strong)) return SUHPE_EINVAL;
    if (unsup_kind < 0 || unsup_kind > 1 || (aug_rot && (aug_mode < 0 || aug_mode > 1))) return SUHPE_EINVAL;
    const int bits = clamp_cut_bits(cut_bits);
    cudaStream_t S = st(stream);
    // carve the workspace (every region starts on a multiple of 4 floats = 16 bytes)
    auto up4 = [](int64_t v) { return (v + 3) & ~(int64_t)3; };
    float* w = workspace;
    unsigned long long* kept = reinterpret_cast<unsigned long long*>(w);  w += 4;
    float* ws_nll = w;        w += up4(b_l);
    float* ws_ent = w;        w += up4(b_u);
    float* ws_work = w;       w += up4(SUHPE_FISHER_CE_WORKSPACE_FLOATS * b_u);
    float* ws_adjusted = w;   w += up4(9 * b_u);
    float* ws_pseudo = w;     w += up4(9 * b_u);
    float* ws_loss_u = w;     w += up4(b_u);
    uint8_t* ws_mask = reinterpret_cast<uint8_t*>(w);
    float* nll_l = losses_l ? losses_l : ws_nll;
    float* ent = entropy ? entropy : ws_ent;
    uint8_t* msk = mask ? mask : ws_mask;
    float* pse = pseudo ? pseudo : ws_pseudo;
    cudaError_t e = cudaSuccess;
#define CK(x) do { if (e == cudaSuccess) e = (x); } while (0)
    // fork: the supervised quadrature (and, for 'ce', the student's unlabeled quadrature) run beside the teacher branch
    CK(cudaEventRecord(c->ev_fork, S));
    CK(cudaStreamWaitEvent(c->s_a, c->ev_fork, 0));
    {
        FisherArgs a{};
        a.A = out_l; a.Rgt = gt_l; a.n = (long long)b_l; a.overreg = overreg; a.nll = nll_l; a.grad = grad_l; a.Rout = Rest_l;
        a.status = status; a.cut_bits = bits;
        CK(launch_fisher_fused(a, c->s_a));
        CK(cudaEventRecord(c->ev_a, c->s_a));
    }
    if (b_u > 0) {
        float* G1 = ws_work;
        if (unsup_kind == 0) {
            CK(cudaStreamWaitEvent(c->s_b, c->ev_fork, 0));
            FisherArgs q{};                       // student, unlabeled: the statistics fisher_CE needs of the prediction
            q.A = pred_strong; q.n = (long long)b_u; q.overreg = 1.0f;
            q.S = ws_work + 3 * b_u; q.G = ws_work + 6 * b_u; q.entropy = ws_work + 9 * b_u;
            q.status = nullptr;                   // non-finite rows are reported by the closing kernel, for kept rows only
            q.cut_bits = bits;
            CK(launch_fisher_fused(q, c->s_b));
            CK(cudaEventRecord(c->ev_b, c->s_b));
        }
        FisherArgs t{};                           // teacher: entropy (src/agent.py:139) and, for 'ce', G1 = d logC/dS
        t.A = pred_weak; t.n = (long long)b_u; t.overreg = 1.0f; t.entropy = ent; t.G = unsup_kind == 0 ? G1 : nullptr;
        t.status = status; t.cut_bits = bits;
        CK(launch_fisher_fused(t, S));
        CK(cudaMemsetAsync(kept, 0, sizeof(unsigned long long), S));
        CK(launch_mask(ent, (long long)b_u, conf_thres_dev, conf_thres_host, msk, kept, S));
        const float* adjusted = pred_weak;
        if (aug_rot) {
            CK(launch_rotate_adjust(pred_weak, aug_rot, (long long)b_u, (int)aug_mode, ws_adjusted, S));
            adjusted = ws_adjusted;
        }
        if (unsup_kind == 1 || pseudo) {
            SvdArgs sv{adjusted, (long long)b_u, pse, nullptr, nullptr, nullptr, status, false};
            CK(launch_proper_svd(sv, S));
        }
        if (unsup_kind == 0) {
            CK(cudaStreamWaitEvent(S, c->ev_b, 0));
            FisherCeArgs ce{adjusted, pred_strong, (long long)b_u, G1, ws_work + 3 * b_u, ws_work + 6 * b_u, ws_work + 9 * b_u,
                            ws_loss_u, grad_strong, status, msk};
            CK(launch_fisher_ce_close(ce, S));
        } else {
            FisherArgs q{};                       // 'nll': the student's Fisher NLL against the projected pseudo labels
            q.A = pred_strong; q.Rgt = pse; q.n = (long long)b_u; q.overreg = overreg; q.nll = ws_loss_u; q.grad = grad_strong;
            q.status = status; q.keep = msk; q.cut_bits = bits;
            CK(launch_fisher_fused(q, S));
        }
    }
    CK(cudaStreamWaitEvent(S, c->ev_a, 0));
    SslFinalizeArgs f{};
    f.nll_l = nll_l; f.b_l = (long long)b_l; f.grad_l = grad_l;
    f.loss_u = ws_loss_u; f.b_u = (long long)b_u; f.grad_u = b_u > 0 ? grad_strong : nullptr;
    f.mask = msk; f.kept = kept; f.ssl_lambda = ssl_lambda; f.losses = losses; f.losses_u_out = b_u > 0 ? losses_u : nullptr;
    CK(launch_ssl_finalize(f, S));
#undef CK
    return rc(e);
}

}  // extern "C"
