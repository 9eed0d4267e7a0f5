// kernels.cuh -- argument blocks and launchers shared by the .cu files and capi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace suhpe {

// status bits written (atomicOr) by the kernels into an optional device word
constexpr int kStatusNonFinite  = 1;   // A held NaN/Inf (reference: torch.svd raises)
constexpr int kStatusTraceRange = 2;   // trace(R1 R2^T) outside [-1-1e-4, 3+1e-4] (pytorch3d raises ValueError)
constexpr int kStatusNonFiniteCE = 4;  // a cross entropy came out NaN/Inf (reference: assert, fisher_utils.py:98)

// radix-select digit layout over the 32-bit monotone entropy key: 11 | 11 | 10 bits
constexpr int kHistBins1 = 2048, kHistShift1 = 21;
constexpr int kHistBins2 = 2048, kHistShift2 = 10;
constexpr int kHistBins3 = 1024, kHistShift3 = 0;
constexpr int kHistBinsMax = 2048;

struct FisherArgs {
    const float* A;        // (n,9) network output, row-major 3x3 records
    const float* Rgt;      // (n,9) target rotations or nullptr (entropy / projection only)
    const float* Sin;      // (n,3) singular values given directly (then A, Rgt, grad, Rout unused) | nullptr
    long long n;
    float overreg;
    float* nll;            // (n)   -<A,R> + overreg*logC          | nullptr
    float* grad;           // (n,9) d nll_i / d A_i                | nullptr
    float* Rout;           // (n,9) proper-SVD rotation            | nullptr
    float* entropy;        // (n)                                  | nullptr
    float* logC;           // (n)   log normaliser                 | nullptr
    float* S;              // (n,3) proper singular values         | nullptr
    float* G;              // (n,3) d logC / d S                   | nullptr
    unsigned long long* hist;  // (2048) += histogram of the top 11 key bits of entropy, counted in the kernel | nullptr
    int* status;           // |= kStatus* | nullptr
    const uint8_t* keep;   // (n) rows with keep == 0 never raise a status bit (their outputs are still written) | nullptr
    int cut_bits;          // negligible-node cut: skipped mass < 2^-cut_bits of the normaliser sum; <= 0 = off
    long long full_rounds; // set by the launcher: rounds of one 32-sample tile per warp
    int samples_per_warp;  // set by the launcher: samples per warp in the closing round (0..32)
    bool vec_ok;           // set by the launcher: float4 tile I/O allowed
};

struct SvdArgs {
    const float* A; long long n;
    float* R; float* S; float* U; float* V;   // each nullable
    int* status;
    bool vec_ok;
};

struct LaplaceArgs {
    const float* A;        // (n,9)
    const float* Rgt;      // (n,9)
    long long n;
    const float* grid;     // (N,9) SO(3) grid, device resident
    int N;
    float* nll;            // (n)
    float* grad;           // (n,9) | nullptr
    float* mode;           // (n,9) | nullptr
    float* logF;           // (n)   | nullptr
    int* status;
};

struct MetricsArgs {
    const float* Rp;       // (n,9) predictions
    const float* Rg;       // (n,9) ground truth | nullptr (Euler of Rp only)
    const float* gt_euler; // (n,3) degrees (pitch,yaw,roll) | nullptr
    long long n;
    int full_range;        // 0 | 1: src/utils.py:232 (radians); 2: DAD-trained scipy-xyz convention (degrees, limit_angle)
    float* geo_deg;        // (n)   | nullptr
    float* frob;           // (n)   | nullptr
    float* euler;          // (n,3) radians | nullptr
    float* abs_err;        // (n,3) |euler_deg - gt| | nullptr
    float* mae;            // (n)   mean_3 abs_err | nullptr
    double* sums;          // (8) += [geo, frob, |dp|, |dy|, |dr|, mae, 0, 0] | nullptr
    int* status;
};

// closing step of fisher_CE: per-pair frames, cross entropy and its gradient from the K2 statistics
struct FisherCeArgs {
    const float* A1;       // (n,9) target parameters (constants)
    const float* A2;       // (n,9) predicted parameters
    long long n;
    const float* G1;       // (n,3) grad logC of A1          (K2)
    const float* S2;       // (n,3) proper singular values of A2, G2 (n,3) grad logC, H2 (n) entropy   (K2)
    const float* G2;
    const float* H2;
    float* ce;             // (n)
    float* grad;           // (n,9) d ce_i / d A2_i | nullptr
    int* status;
    const uint8_t* keep;   // (n) rows with keep == 0 give ce = 0, grad = 0 and no status bit | nullptr = all rows
};

int device_sm_count();      // SM count of the current device (cached per device)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of on every launch;
// `done_mask` is a static word owned by the call site, one bit per device
template <typename K>
inline cudaError_t allow_dynamic_smem(K kernel, size_t bytes, unsigned long long& done_mask) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_mask & bit) return cudaSuccess;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err == cudaSuccess) done_mask |= bit;     // benign race: setting the attribute twice is harmless
    return err;
}
cudaError_t launch_fisher_fused(FisherArgs p, cudaStream_t stream);
cudaError_t launch_fisher_ce_close(FisherCeArgs p, cudaStream_t stream);
cudaError_t launch_proper_svd(SvdArgs p, cudaStream_t stream);
// d L / d A from G = d L / d R for R = U V^T of the proper SVD (U, V, S as K1 wrote them)
cudaError_t launch_polar_backward(const float* U, const float* V, const float* S, const float* G, long long n, float* out,
                                  cudaStream_t stream);
// out[i,:] = in[i,:] * row_weight[i] * *scalar_weight, zero rows where keep[i] == 0 (each factor nullable)
cudaError_t launch_scale_rows(const float* in, long long n, int width, const float* row_weight, const float* scalar_weight,
                              const uint8_t* keep, float* out, cudaStream_t stream);
// closing reduction of the SSL loss head (see suhpe_ssl_step_f32): means, loss_all, scaled gradients
struct SslFinalizeArgs {
    const float* nll_l; long long b_l;      // per-sample supervised losses
    float* grad_l;                          // (b_l,9) in: d nll_i/d out_i, out: d loss_all / d out_l
    const float* loss_u; long long b_u;     // per-sample unsupervised losses (any value on filtered rows)
    float* grad_u;                          // (b_u,9) in: d l_i / d pred_strong_i, out: d loss_all / d pred_strong
    const uint8_t* mask;                    // (b_u)
    const unsigned long long* kept;         // mask popcount
    float ssl_lambda;
    float* losses;                          // [4] loss_sup, unsuper_loss, mask_ratio, loss_all
    float* losses_u_out;                    // (b_u) masked per-sample losses | nullptr
};
cudaError_t launch_ssl_finalize(SslFinalizeArgs a, cudaStream_t stream);
cudaError_t launch_rotate_adjust(const float* P, const float* Raug, long long n, int mode, float* out, cudaStream_t stream);
cudaError_t launch_laplace(LaplaceArgs p, cudaStream_t stream);

// EMA / EMAN teacher update (ema_kernels.cu): one launch carries up to kEmaMaxTensors tensors and
// kEmaMaxBlocks CTAs of kEmaChunk elements each in its parameter block
constexpr int kEmaMaxTensors = 48;
constexpr int kEmaMaxBlocks = 320;
constexpr int kEmaChunk = 32768;            // elements per CTA (a multiple of 4)
struct EmaLaunch {
    float* ema[kEmaMaxTensors];
    const float* src[kEmaMaxTensors];
    long long numel[kEmaMaxTensors];
    int block_chunk[kEmaMaxBlocks];
    unsigned char block_tensor[kEmaMaxBlocks];
};
cudaError_t launch_ema_update(float* const* ema, const float* const* src, const long long* numel, int count,
                              float alpha, float one_minus_alpha, int mode, cudaStream_t stream);
cudaError_t launch_metrics(MetricsArgs p, cudaStream_t stream);

// ---- radix select over entropy keys -----------------------------------------
// state block kept on the device between passes (no host round trip on 1 GPU)
struct SelectState {
    unsigned long long k_remaining;   // rank still to locate inside the current prefix
    unsigned int prefix;              // key bits fixed so far (high bits)
    unsigned int pass;                // passes completed
    unsigned int threshold_key;       // valid after pass 3
    float threshold;                  // valid after pass 3
    unsigned long long kept;          // # entropies strictly below threshold (after mask)
};

cudaError_t launch_select_hist(const float* e, long long n, int pass, const SelectState* state,
                               unsigned long long* hist, cudaStream_t stream);
// hist_parts: (parts, bins) gathered histograms, summed on the fly
cudaError_t launch_select_scan(const unsigned long long* hist_parts, int parts, int pass,
                               SelectState* state, cudaStream_t stream);
cudaError_t launch_select_init(SelectState* state, unsigned long long k, cudaStream_t stream);
cudaError_t launch_mask(const float* e, long long n, const float* thr_dev, float thr_host,
                        uint8_t* mask, unsigned long long* kept, cudaStream_t stream);

// FP32 pipe probe (roofline denominator measured on the box): returns FMA count executed
// K2 pass-body probe (profiling aid): variant bits 0-1 run type, 2 no LDS, 3 no MUFU, 4 no mask;
// passes executed = blocks * 16 warps * iters, 128 nodes each
cudaError_t launch_body_probe(float* sink, int variant, int iters, int blocks, cudaStream_t stream);
cudaError_t launch_fp32_probe(float* sink, int variant, int iters, int blocks, cudaStream_t stream);

}  // namespace suhpe
