/* semiuhpe_b200.h -- C ABI of the B200-native rotation-distribution hot path.
 *
 * One shared library (libsemiuhpe_b200.so, built for sm_100a by
 * __graft_entry__.build()) exports exactly these symbols.  Plain pointers and
 * sizes only; device pointers unless the function name ends in `_host`.
 * Every function returns 0 on success, -(cudaError_t) for a CUDA failure and
 * SUHPE_EINVAL for a bad argument; nothing throws and nothing synchronises the
 * device except the `_host` pipeline and suhpe_select_read.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  The library keeps
 * no mutable process-wide settings: everything a call depends on is an argument
 * (or lives in a handle the caller created), so calls from several threads,
 * streams and devices do not interact.
 *
 * The reference (hnuzhy/SemiUHPE) has no FFI: the path sits behind plain Python
 * functions.  Each entry point names the reference function(s) it replaces
 * (paths relative to the reference root); the modules under semiuhpe_b200/ mirror those
 * Python signatures on top of this ABI and INTEGRATION.md shows the binding.
 *
 * Records: a rotation / parameter matrix is 9 contiguous fp32 (row-major 3x3).
 */
#ifndef SEMIUHPE_B200_H
#define SEMIUHPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUHPE_ABI_VERSION 2
#define SUHPE_EINVAL (-100000)

/* bits OR-ed into the optional device `status` word */
#define SUHPE_STATUS_NONFINITE   1  /* A held NaN/Inf: torch.svd raises there (fisher_utils.py:28) */
#define SUHPE_STATUS_TRACE_RANGE 2  /* trace(R1 R2^T) out of [-1-1e-4, 3+1e-4]: pytorch3d raises ValueError */
#define SUHPE_STATUS_NONFINITE_CE 4 /* a cross entropy is NaN/Inf: the reference asserts (fisher_utils.py:98) */

#define SUHPE_HIST_BINS 2048        /* uint64 counters per radix histogram */
#define SUHPE_SELECT_STATE_BYTES 32 /* opaque device block, see suhpe_select_read */

int suhpe_abi_version(void);
const char* suhpe_error_string(int code);

/* K1 -- proper SVD projection onto SO(3):  A = U diag(S) V^T, det U = det V = +1,
 * S0 >= S1 >= |S2| (S2 signed), R = U V^T.
 * Replaces batch_torch_A_to_R (src/fisher/fisher_utils.py:39-48), analytical_mode
 * (src/laplace/rotation_laplace.py:102-115) and proper_svd
 * (src/fisher/between_bingham_fisher.py:63-82).  R,S,U,V,status nullable. */
int suhpe_proper_svd_f32(const float* A, int64_t n, float* R, float* S, float* U, float* V,
                         int* status, void* stream);
/* Backward of R = U V^T w.r.t. A (batch_torch_A_to_R is differentiable through torch.svd in the reference,
 * src/fisher/fisher_utils.py:39-48): gradA = U Q V^T, Q_ij = (P_ij - P_ji) / (S_i + S_j), P = U^T gradR V,
 * with U, V, S exactly as suhpe_proper_svd_f32 wrote them.  S_i + S_j = 0 contributes nothing. */
int suhpe_proper_svd_backward_f32(const float* U, const float* V, const float* S, const float* gradR, int64_t n,
                                  float* gradA, void* stream);

/* K2 -- fused matrix-Fisher head: per sample
 *   nll     = -<A,Rgt> + overreg * logC(S)                 KL_Fisher   fisher_utils.py:21-36
 *   grad    = d nll / dA = -Rgt + overreg * U diag(g) V^T  class_logC_F.backward torch_norm_factor.py:80-90
 *   Rout    = U V^T                                        batch_torch_A_to_R fisher_utils.py:39-48
 *   entropy = log f(S) + sum_j S_j (1 - g_j)               fisher_entropy fisher_utils.py:70-81
 *   logC, S (n,3), G = dlogC/dS (n,3)                      logC_F torch_norm_factor.py:66-92
 *   hist   += histogram of the top 11 bits of the entropy keys: the first radix-select pass, counted
 *             inside the kernel as each entropy is produced (warp-aggregated atomic adds into the 2048
 *             counters; `entropy` itself may be NULL)
 * Rgt may be NULL (then nll = overreg*logC and grad has no -Rgt term); every output is nullable.
 *
 * cut_bits -- negligible-node cut of the quadrature.  Every integrand of
 * src/fisher/torch_norm_factor.py:33-63 decays like exp(-c(1-x)) away from x = 1, so a prefix of the 512
 * nodes contributes less than 2^-cut_bits of the normaliser sum in total; K2 skips that prefix where a
 * per-sample bound PROVES it (rigorous upper bound of the prefix mass against a rigorous lower bound of
 * the sum: cut_threshold / cut_index in csrc/so3_math.cuh, float64 check in tests/test_emul_math.py).
 * SUHPE_CUT_BITS_DEFAULT = 26 is an eighth of an fp32 ulp of the sum (below what the reference's own fp32
 * torch.sum resolves); 0 evaluates all 512 nodes of every integral. */
#define SUHPE_CUT_BITS_DEFAULT 26
int suhpe_fisher_fused_f32(const float* A, const float* Rgt, int64_t n, float overreg, int32_t cut_bits,
                           float* nll, float* grad, float* Rout, float* entropy, float* logC,
                           float* S, float* G, uint64_t* hist, int* status, void* stream);

/* fisher_CE(A1 = target, A2 = prediction) -> ce (n) and d ce_i / d A2_i (n,9, nullable): the
 * cross entropy of two matrix-Fisher densities through their Bingham forms, the reference's default
 * unsupervised loss (src/fisher/fisher_utils.py:84-99 -> between_bingham_fisher.py:107-152 ->
 * bingham_utils.py:5-32, reproduced with its row/column quirk at bingham_utils.py:27).  Two K2
 * launches (the quadratures of A1 and A2) and one closing kernel that evaluates the value and
 * the gradient in closed form (what autograd yields through torch.svd, matrix_to_quaternion and
 * the logC_F backward).  A1 is a constant (the agent detaches the teacher: src/agent.py:107).
 * keep (n, nullable): rows with keep[i] == 0 are the ones the reference's boolean gather drops before
 * the loss (src/agent.py:152-160); they get ce = 0 and a zero gradient and raise no status bit, whatever
 * their inputs hold.
 * workspace: SUHPE_FISHER_CE_WORKSPACE_FLOATS * n floats of device scratch, caller-owned. */
#define SUHPE_FISHER_CE_WORKSPACE_FLOATS 10
int suhpe_fisher_ce_f32(const float* A1, const float* A2, int64_t n, int32_t cut_bits, const uint8_t* keep,
                        float* ce, float* gradA2, float* workspace, int* status, void* stream);
/* Same, with G1 = d logC / dS of the target (n,3) supplied by the caller: the training step has
 * it already -- the entropy launch on the teacher output (suhpe_fisher_fused_f32 with G) produced
 * it, and the rotate-augmentation adjustment (a rotation applied on one side) leaves the singular
 * values, hence G1, unchanged -- so the cross entropy costs one quadrature instead of two. */
int suhpe_fisher_ce_with_g1_f32(const float* A1, const float* G1, const float* A2, int64_t n, int32_t cut_bits,
                                const uint8_t* keep, float* ce, float* gradA2, float* workspace, int* status,
                                void* stream);

/* out[i, :] = w_i * in[i, :] for rows of `width` floats; w_i = row_weight[i] (n, nullable) times
 * *scalar_weight (device scalar, nullable).  The backward of the loss mirrors: the per-sample gradient
 * the forward launch produced times the incoming d L / d loss_i, without leaving the stream.
 * keep (n, nullable): rows with keep[i] == 0 are written as zeros regardless of `in` (0 * NaN never forms).
 * in == out is allowed. */
int suhpe_scale_rows_f32(const float* in, int64_t n, int32_t width, const float* row_weight,
                         const float* scalar_weight, const uint8_t* keep, float* out, void* stream);

/* Rotate-augmentation adjustment of the teacher's parameter matrices before they become pseudo
 * labels (src/agent.py:110-119): mode 0 (train_labeled "DAD3DHeads") out = aug_rot * pred;
 * mode 1 ("300WLP") out = (D aug_rot D pred^T)^T with D = diag(1,-1,-1).  All (n,9) row-major. */
int suhpe_rotate_adjust_f32(const float* pred, const float* aug_rot, int64_t n, int32_t mode, float* out, void* stream);

/* EMA / EMAN teacher update, SSLAgent.update_ema_variables (src/agent.py:277-299), for a list of
 * `count` fp32 tensors in one call (the host arrays ema / src / numel are read before the call
 * returns; the tensors are device memory):
 *   mode 0 (config.eman, :293)  ema[i] = ema[i]*alpha + one_minus_alpha*src[i]   (products rounded, then added)
 *   mode 1 (parameters, :298)   ema[i].mul_(alpha).add_(src[i], alpha=one_minus_alpha)   (fused multiply-add)
 * alpha is the value AFTER the reference's warm-up rule (:278-283), already rounded to fp32 like ATen
 * rounds a Python scalar; ema[i] and src[i] must not overlap. */
int suhpe_ema_update_f32(float* const* ema, const float* const* src, const int64_t* numel, int32_t count,
                         float alpha, float one_minus_alpha, int32_t mode, void* stream);

/* K2 on given singular values: logC_F(S) and its backward G = dlogC/dS, entropy(S)
 * (src/fisher/torch_norm_factor.py:66-92 `logC_F`; S (n,3) sorted s1>=s2>=|s3| like every
 * reference call site).  Outputs nullable. */
int suhpe_fisher_from_s_f32(const float* S, int64_t n, int32_t cut_bits, float* logC, float* G, float* entropy,
                            int* status, void* stream);

/* K2L -- rotation-Laplace NLL forward+backward against an SO(3) grid (N,9), device resident.
 * Replaces NLL_loss("RLaplace") / log_pdf / logF_const_laplace / power_fn_sqrtL2_proper /
 * analytical_mode (src/laplace/rotation_laplace.py:24-34,58-115,140-173) and their autograd
 * backward.  grad, mode, logF, status nullable. */
int suhpe_laplace_nll_f32(const float* A, const float* Rgt, int64_t n, const float* grid, int32_t N,
                          float* nll, float* grad, float* mode, float* logF, int* status, void* stream);

/* K3 -- dynamic-entropy filter (src/agent.py:148-150,403-407).
 * Exact k-th smallest (0-based rank k in numpy.sort order: ascending, -0==+0, NaN last)
 * by three radix passes.  Single GPU: suhpe_entropy_threshold_f32 chains everything on
 * `stream`.  Multi GPU: call suhpe_select_init, then per pass p=1,2,3
 * suhpe_select_hist (local shard) -> all-gather the SUHPE_HIST_BINS counters of every rank
 * (NCCL) -> suhpe_select_scan over the gathered (parts,SUHPE_HIST_BINS) block.
 * `state` is SUHPE_SELECT_STATE_BYTES of device memory; `hist` SUHPE_HIST_BINS uint64. */
int suhpe_select_init(void* state, uint64_t k, void* stream);
int suhpe_select_hist_f32(const float* entropy, int64_t n, int32_t pass, const void* state,
                          uint64_t* hist, void* stream);
int suhpe_select_scan(const uint64_t* hist_parts, int32_t parts, int32_t pass, void* state, void* stream);
/* first_pass_hist: optional histogram already produced by suhpe_fisher_fused_f32 (skips pass 1's read) */
int suhpe_entropy_threshold_f32(const float* entropy, int64_t n, uint64_t k, void* state,
                                uint64_t* hist_scratch, const uint64_t* first_pass_hist, void* stream);
/* device pointer to the fp32 threshold inside `state` (valid after pass 3) */
const float* suhpe_select_threshold_ptr(const void* state);
/* blocking read-back of the state block: threshold, its key, kept count */
int suhpe_select_read(const void* state, float* threshold, uint32_t* key, uint64_t* kept, void* stream);
/* mask[i] = entropy[i] < thr  (thr read from thr_dev if non-NULL, else thr_host); kept += popcount.
 * mask nullable (count only); kept nullable. */
int suhpe_entropy_mask_f32(const float* entropy, int64_t n, const float* thr_dev, float thr_host,
                           uint8_t* mask, uint64_t* kept, void* stream);

/* K4 -- error metrics over rotation pairs.
 *   geo_deg  rad2deg(so3_relative_angle(Rp,Rg))       src/agent.py:449-451, eval.py:88-89
 *   frob     ||I - Rp Rg^T||_F                        eval.py:93-98
 *   euler    (pitch,yaw,roll) radians of Rp           src/utils.py:232-260   (full_range 0 | 1)
 *            full_range = 2 (SUHPE_EULER_DAD): the DAD-trained convention of eval.py:66-74 --
 *            scipy as_euler("xyz") of Rp^T, [roll,pitch,yaw] = limit_angle([a2, a0-180, a1]) -- in
 *            DEGREES; abs_err / mae then compare against gt_euler directly
 *   abs_err  |euler*180/pi - gt_euler|, mae = mean_3  src/agent.py:452-454, eval.py:76-83
 *   sums[8] += {geo, frob, |dpitch|, |dyaw|, |droll|, mae, 0, 0} in fp64 (eval.py:125-133 means)
 * Rg may be NULL when only Euler angles are wanted; all outputs nullable. */
#define SUHPE_EULER_DAD 2
int suhpe_so3_metrics_f32(const float* Rp, const float* Rg, const float* gt_euler_deg, int64_t n,
                          int32_t full_range, float* geo_deg, float* frob, float* euler,
                          float* abs_err, float* mae, double* sums, int* status, void* stream);

/* Host-buffer pipeline (what bench.py's e2e leg and a non-torch host would call):
 * the teacher-side filter step over a pool of n (A,Rgt) pairs living in HOST memory
 * (pinned for full PCIe rate).  Chunks move through three in-order queues -- host->device
 * copies, K2 (+fused first histogram), device->host copies -- over a ring of four chunk
 * buffers, so both copy engines and the SMs are busy at once; then K3 selects the k-th
 * smallest entropy of the whole pool and the mask is emitted and copied back.
 * Replaces the per-batch loop of src/agent.py:357-417 (teacher forward -> fisher_entropy ->
 * host sort).  Host outputs nullable; threshold/kept written on return (the call blocks).
 * `cut_bits` (see suhpe_fisher_fused_f32) is fixed per pipeline at creation. */
typedef struct suhpe_pipeline suhpe_pipeline;
int suhpe_pipeline_create(suhpe_pipeline** out, int64_t max_n, int64_t chunk, int32_t cut_bits);
int suhpe_pipeline_destroy(suhpe_pipeline* p);
int suhpe_fisher_filter_host(suhpe_pipeline* p, const float* A_host, const float* Rgt_host, int64_t n,
                             float overreg, uint64_t k, float* nll_host, float* grad_host,
                             float* entropy_host, uint8_t* mask_host, float* threshold, uint64_t* kept);
/* The same pipeline in two phases, for pools sharded over several GPUs: phase A streams this
 * rank's shard through K2 and leaves the entropies in ent_dev (n), ADDS the first radix
 * histogram into hist_dev (2048 counters, zeroed by the caller) and ORs status bits into
 * status_dev (all caller-owned device memory).  Returns once everything is queued; `stream`
 * is made to wait for the last kernel, so the caller queues the global select on it (K3 passes
 * + an all-gather of the histograms, semiuhpe_b200/distributed.py) while the device->host copies
 * drain.  suhpe_pipeline_sync blocks until the host buffers are complete.  Calls on one pipeline
 * may follow each other without a sync in between: a chunk buffer is reused only after the kernel
 * and the device->host copies of its previous use -- in this or an earlier call -- have finished
 * (the HOST output arrays of the earlier call must of course not be reused before a sync). */
int suhpe_fisher_pool_host(suhpe_pipeline* p, const float* A_host, const float* Rgt_host, int64_t n,
                           float overreg, float* nll_host, float* grad_host, float* entropy_host,
                           float* ent_dev, uint64_t* hist_dev, int* status_dev, void* stream);
int suhpe_pipeline_sync(suhpe_pipeline* p);

/* The loss head of one semi-supervised training step (SSLAgent.forward + the backward of
 * train_func, src/agent.py:76-83,99-166,194-210) in ONE call: a fixed sequence of launches over
 * three forked streams, no host synchronisation, no dynamic shapes -- CUDA-graph capturable.
 *   labeled     out_l (b_l,9) student output, gt_l (b_l,9) rotations:
 *               losses_l = KL_Fisher(out_l, gt_l, overreg), Rest_l = batch_torch_A_to_R(out_l)      :76-83
 *   unlabeled   pred_weak (b_u,9) teacher output (a constant), pred_strong (b_u,9) student output:
 *               entropy = fisher_entropy(pred_weak), mask = entropy < conf_thres                    :139,148
 *               adjusted = rotate-augmentation adjustment of pred_weak (aug_rot nullable)           :110-122
 *               unsup_kind 0 ('ce'):  l_u = fisher_CE(adjusted, pred_strong)                        :155
 *               unsup_kind 1 ('nll'): l_u = KL_Fisher(pred_strong, batch_torch_A_to_R(adjusted))    :157
 *   losses[0] = mean(losses_l)                       losses[1] = sum(mask ? l_u : 0) / b_u   (:163,166)
 *   losses[2] = mask ratio                           losses[3] = losses[0] + ssl_lambda * losses[1] (:203)
 *   grad_l      = d losses[3] / d out_l        (b_l,9)
 *   grad_strong = d losses[3] / d pred_strong  (b_u,9), exactly zero for filtered rows
 * conf_thres is read from conf_thres_dev if non-NULL (e.g. suhpe_select_threshold_ptr) else conf_thres_host.
 * Optional outputs (nullable): Rest_l (b_l,9), entropy (b_u), mask (b_u), pseudo (b_u,9) = the projected
 * pseudo labels of every row, losses_l (b_l), losses_u (b_u, zero for filtered rows).
 * b_u may be 0 (supervised step, train_func_s1: src/agent.py:253-270); then only the labeled part runs.
 * workspace: suhpe_ssl_step_workspace_floats(b_l, b_u) floats of caller-owned device scratch, 16-byte aligned, private to
 * the call until the work queued on `stream` has run.  The handle owns only the two forked streams and their events --
 * no data -- so one handle per device serves every caller stream (calls from different streams serialise their forked
 * branches but do not share a byte), and nothing is allocated inside the call: it can be issued during stream capture. */
typedef struct suhpe_ssl_step suhpe_ssl_step;
int suhpe_ssl_step_create(suhpe_ssl_step** out);
int suhpe_ssl_step_destroy(suhpe_ssl_step* ctx);
int64_t suhpe_ssl_step_workspace_floats(int64_t b_l, int64_t b_u);
int suhpe_ssl_step_f32(suhpe_ssl_step* ctx, const float* out_l, const float* gt_l, int64_t b_l,
                       const float* pred_weak, const float* pred_strong, int64_t b_u,
                       const float* aug_rot, int32_t aug_mode, const float* conf_thres_dev, float conf_thres_host,
                       float overreg, float ssl_lambda, int32_t unsup_kind, int32_t cut_bits,
                       float* workspace,
                       float* losses, float* grad_l, float* grad_strong,
                       float* Rest_l, float* entropy, uint8_t* mask, float* pseudo, float* losses_l, float* losses_u,
                       int* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMIUHPE_B200_H */
