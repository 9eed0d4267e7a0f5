// so3_math.cuh -- per-sample SO(3) arithmetic shared by all kernels (sm_100a).
//
// Everything here is scalar, register-resident and branch-light; the kernels in
// fisher_kernels.cu / laplace_kernels.cu / metrics_kernels.cu decide how samples
// and quadrature nodes are spread over lanes.  The functions are
// __host__ __device__ so that tests/emul (a host-only build used by the CPU test
// suite) can exercise exactly this source without a GPU; the product never runs
// the host instantiation.
//
// Reference behaviour restated (hnuzhy/SemiUHPE, paths relative to its root):
//   proper SVD / rotation      src/fisher/fisher_utils.py:27-31,39-48
//                              src/fisher/between_bingham_fisher.py:63-82
//   A&S Bessel polynomials     src/fisher/torch_norm_factor.py:4-19
//   512-node trapezoid         src/fisher/torch_norm_factor.py:21-31
//   integrands                 src/fisher/torch_norm_factor.py:33-63
//   Euler angles               src/utils.py:232-260
//   geodesic angle             pytorch3d 0.7.2 so3_relative_angle (src/agent.py:450)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define SUHPE_HD __host__ __device__ __forceinline__

namespace suhpe {

// ----------------------------------------------------------------------------
// MUFU wrappers.  On the device these are single SFU instructions; the host
// versions exist only for tests/emul.
// ----------------------------------------------------------------------------
SUHPE_HD float mufu_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}
SUHPE_HD float mufu_ex2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}
SUHPE_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
SUHPE_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
SUHPE_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
SUHPE_HD float sqrt_rn(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

// ----------------------------------------------------------------------------
// Proper SVD of a 3x3 matrix:  A = U diag(s) V^T,  det U = det V = +1,
// s0 >= s1 >= |s2|, s2 carries the sign of det A.   R = U V^T is the rotation
// the reference builds as [u1,u2,u3*det(U V^T)] V^T, and (s0,s1,s2) is its
// S_sign.
//
// Method: one-sided (Hestenes) Jacobi on the columns of A -- rotations only, so
// V stays in SO(3) -- followed by a rotation-only sort and a cross product for
// the third left vector.  No division by the smallest singular value anywhere,
// so rank-deficient input (A = 0 gives U = V = R = I like LAPACK) is safe.
// ----------------------------------------------------------------------------
constexpr int kJacobiSweeps = 5;          // fp32: converged to rounding for every tested spectrum
constexpr int kJacobiSweepsF64 = 7;       // fp64 from scratch: two more quadratic sweeps reach 1e-16
constexpr int kJacobiPolishF64 = 2;       // fp64 sweeps on top of a converged fp32 run (K2L set-up; one already reaches 5e-13 on T)
#ifndef SUHPE_K2L_SETUP_F64_ONLY
#define SUHPE_K2L_SETUP_F64_ONLY 0
#endif

// scalar-type shims so the same source instantiates in fp32 (every kernel) and fp64 (K2L's
// per-sample set-up, where the Laplace NLL cancels sum(s) against <A,R>)
SUHPE_HD float  fma_t(float a, float b, float c) { return fmaf(a, b, c); }
SUHPE_HD double fma_t(double a, double b, double c) { return fma(a, b, c); }
SUHPE_HD float  sqrt_t(float a) { return sqrt_rn(a); }
SUHPE_HD double sqrt_t(double a) { return sqrt(a); }
SUHPE_HD float  div_t(float a, float b) { return div_rn(a, b); }
SUHPE_HD double div_t(double a, double b) { return a / b; }
SUHPE_HD float  abs_t(float a) { return fabsf(a); }
SUHPE_HD double abs_t(double a) { return fabs(a); }

template <typename T>
SUHPE_HD void jacobi_pair(T* bp, T* bq, T* vp, T* vq, T skip) {
    const T alpha = fma_t(bp[0], bp[0], fma_t(bp[1], bp[1], bp[2] * bp[2]));
    const T beta  = fma_t(bq[0], bq[0], fma_t(bq[1], bq[1], bq[2] * bq[2]));
    const T gamma = fma_t(bp[0], bq[0], fma_t(bp[1], bq[1], bp[2] * bq[2]));
    const T d  = beta - alpha;
    const T g2 = gamma + gamma;
    const T h  = sqrt_t(fma_t(d, d, g2 * g2));
    const T den = abs_t(d) + h;
    // tan(theta) = sgn(d) 2 gamma / (|d| + sqrt(d^2 + 4 gamma^2)); 0 when already orthogonal
    T t = (den > T(0)) ? div_t((d < T(0)) ? -g2 : g2, den) : T(0);
    // rotations below the rounding level of the columns are skipped (also avoids
    // denormal-driven drift when gamma ~ 0)
    if (gamma * gamma <= skip * alpha * beta) t = T(0);
    const T c = div_t(T(1), sqrt_t(fma_t(t, t, T(1))));
    const T s = t * c;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T x = bp[i], y = bq[i];
        bp[i] = fma_t(c, x, -s * y);
        bq[i] = fma_t(s, x, c * y);
        const T vx = vp[i], vy = vq[i];
        vp[i] = fma_t(c, vx, -s * vy);
        vq[i] = fma_t(s, vx, c * vy);
    }
}

// swap columns p,q as the rotation (bp,bq) <- (bq,-bp): keeps det V = +1
template <typename T>
SUHPE_HD void rot_swap(bool doit, T* bp, T* bq, T* vp, T* vq, T& np_, T& nq_) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T x = bp[i], y = bq[i];
        bp[i] = doit ? y : x;
        bq[i] = doit ? -x : y;
        const T vx = vp[i], vy = vq[i];
        vp[i] = doit ? vy : vx;
        vq[i] = doit ? -vx : vy;
    }
    const T a = np_, b = nq_;
    np_ = doit ? b : a;
    nq_ = doit ? a : b;
}

// A (fp32 input), U, V row-major 3x3 in T.  Returns false when A holds a non-finite value
// (the reference's torch.svd raises in that case).
// Vstart (nullable, row-major, fp32): an orthogonal matrix to start the one-sided iteration from -- the V of a
// lower-precision run of the same routine, which leaves `sweeps` only the polishing to do.
template <typename T>
SUHPE_HD bool proper_svd3_t(const float* A, T* U, T* V, T* s, const float* Vstart = nullptr,
                            int sweeps = (sizeof(T) == 8 ? kJacobiSweepsF64 : kJacobiSweeps)) {
    constexpr bool kF64 = sizeof(T) == 8;
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) amax = fmaxf(amax, fabsf(A[i]));
    bool finite = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) finite = finite && (fabsf(A[i]) <= 3.402823466e38f);
    // exact power-of-two prescale: columns land in [1,2) * O(1); undone exactly on s
#if defined(__CUDA_ARCH__)
    const int ebits = (__float_as_int(amax) >> 23) & 0xff;
#else
    union { float f; int i; } cv; cv.f = amax;
    const int ebits = (cv.i >> 23) & 0xff;
#endif
    float down = 1.0f, up = 1.0f;
    if (ebits >= 1 && ebits <= 253) {
#if defined(__CUDA_ARCH__)
        down = __int_as_float((254 - ebits) << 23);
        up   = __int_as_float(ebits << 23);
#else
        union { float f; int i; } a, b; a.i = (254 - ebits) << 23; b.i = ebits << 23;
        down = a.f; up = b.f;
#endif
    }
    // column vectors of the scaled matrix (the scaling is exact in either precision)
    T b0[3] = {T(A[0] * down), T(A[3] * down), T(A[6] * down)};
    T b1[3] = {T(A[1] * down), T(A[4] * down), T(A[7] * down)};
    T b2[3] = {T(A[2] * down), T(A[5] * down), T(A[8] * down)};
    T v0[3] = {T(1), T(0), T(0)}, v1[3] = {T(0), T(1), T(0)}, v2[3] = {T(0), T(0), T(1)};
    if (Vstart) {
        // start from A V0 instead of A: V0 = Vstart made orthogonal to the working precision by one Newton-Schulz
        // step, V0 = Vs (3 I - Vs^T Vs) / 2 (Vs is orthogonal to fp32 rounding: the step squares that error)
        const T c0[3] = {T(Vstart[0]), T(Vstart[3]), T(Vstart[6])}, c1[3] = {T(Vstart[1]), T(Vstart[4]), T(Vstart[7])},
                c2[3] = {T(Vstart[2]), T(Vstart[5]), T(Vstart[8])};
        auto dot = [](const T* x, const T* y) { return fma_t(x[0], y[0], fma_t(x[1], y[1], x[2] * y[2])); };
        const T g00 = dot(c0, c0), g11 = dot(c1, c1), g22 = dot(c2, c2), g01 = dot(c0, c1), g02 = dot(c0, c2), g12 = dot(c1, c2);
        const T m00 = T(1.5) - T(0.5) * g00, m11 = T(1.5) - T(0.5) * g11, m22 = T(1.5) - T(0.5) * g22;
        const T m01 = T(-0.5) * g01, m02 = T(-0.5) * g02, m12 = T(-0.5) * g12;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            v0[i] = fma_t(c0[i], m00, fma_t(c1[i], m01, c2[i] * m02));
            v1[i] = fma_t(c0[i], m01, fma_t(c1[i], m11, c2[i] * m12));
            v2[i] = fma_t(c0[i], m02, fma_t(c1[i], m12, c2[i] * m22));
        }
        const T a0[3] = {b0[0], b1[0], b2[0]}, a1[3] = {b0[1], b1[1], b2[1]}, a2[3] = {b0[2], b1[2], b2[2]};   // rows of A
        b0[0] = dot(a0, v0); b0[1] = dot(a1, v0); b0[2] = dot(a2, v0);
        b1[0] = dot(a0, v1); b1[1] = dot(a1, v1); b1[2] = dot(a2, v1);
        b2[0] = dot(a0, v2); b2[1] = dot(a1, v2); b2[2] = dot(a2, v2);
    }
    const T skip = kF64 ? T(1e-32) : T(1e-16f);
#pragma unroll 1
    for (int sweep = 0; sweep < sweeps; ++sweep) {
        jacobi_pair(b0, b1, v0, v1, skip);
        jacobi_pair(b0, b2, v0, v2, skip);
        jacobi_pair(b1, b2, v1, v2, skip);
    }
    T n0 = fma_t(b0[0], b0[0], fma_t(b0[1], b0[1], b0[2] * b0[2]));
    T n1 = fma_t(b1[0], b1[0], fma_t(b1[1], b1[1], b1[2] * b1[2]));
    T n2 = fma_t(b2[0], b2[0], fma_t(b2[1], b2[1], b2[2] * b2[2]));
    rot_swap(n0 < n1, b0, b1, v0, v1, n0, n1);
    rot_swap(n0 < n2, b0, b2, v0, v2, n0, n2);
    rot_swap(n1 < n2, b1, b2, v1, v2, n1, n2);

    // left vectors: u0 = b0/|b0|, u1 = normalised (b1 - (b1.u0)u0), u2 = u0 x u1
    T u0[3], u1[3], u2[3];
    const T sig0 = sqrt_t(n0);
    if (sig0 > T(0)) {
        const T inv = div_t(T(1), sig0);
        u0[0] = b0[0] * inv; u0[1] = b0[1] * inv; u0[2] = b0[2] * inv;
    } else {
        u0[0] = T(1); u0[1] = T(0); u0[2] = T(0);
    }
    const T p = fma_t(b1[0], u0[0], fma_t(b1[1], u0[1], b1[2] * u0[2]));
    T w[3] = {fma_t(-p, u0[0], b1[0]), fma_t(-p, u0[1], b1[1]), fma_t(-p, u0[2], b1[2])};
    T nw = fma_t(w[0], w[0], fma_t(w[1], w[1], w[2] * w[2]));
    if (!(nw > T(1e-30))) {
        // rank <= 1: any unit vector orthogonal to u0 (axis least aligned with u0)
        const T a0 = abs_t(u0[0]), a1 = abs_t(u0[1]), a2 = abs_t(u0[2]);
        const int k = (a1 <= a0 && a1 <= a2) ? 1 : ((a2 <= a0 && a2 <= a1) ? 2 : 0);
        const T uk = (k == 0) ? u0[0] : ((k == 1) ? u0[1] : u0[2]);
        w[0] = ((k == 0) ? T(1) : T(0)) - uk * u0[0];
        w[1] = ((k == 1) ? T(1) : T(0)) - uk * u0[1];
        w[2] = ((k == 2) ? T(1) : T(0)) - uk * u0[2];
        nw = fma_t(w[0], w[0], fma_t(w[1], w[1], w[2] * w[2]));
    }
    {
        const T inv = div_t(T(1), sqrt_t(nw));
        u1[0] = w[0] * inv; u1[1] = w[1] * inv; u1[2] = w[2] * inv;
    }
    u2[0] = fma_t(u0[1], u1[2], -u0[2] * u1[1]);
    u2[1] = fma_t(u0[2], u1[0], -u0[0] * u1[2]);
    u2[2] = fma_t(u0[0], u1[1], -u0[1] * u1[0]);

    s[0] = sig0 * T(up);
    s[1] = fma_t(b1[0], u1[0], fma_t(b1[1], u1[1], b1[2] * u1[2])) * T(up);
    s[2] = fma_t(b2[0], u2[0], fma_t(b2[1], u2[1], b2[2] * u2[2])) * T(up);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        U[3 * i + 0] = u0[i]; U[3 * i + 1] = u1[i]; U[3 * i + 2] = u2[i];
        V[3 * i + 0] = v0[i]; V[3 * i + 1] = v1[i]; V[3 * i + 2] = v2[i];
    }
    return finite;
}

SUHPE_HD bool proper_svd3(const float* A, float* U, float* V, float* s) { return proper_svd3_t<float>(A, U, V, s); }

// R = U diag(d0,d1,d2) V^T  (row-major)
SUHPE_HD void u_diag_vt(const float* U, const float* V, float d0, float d1, float d2, float* R) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = U[3 * i + 0] * d0, b = U[3 * i + 1] * d1, c = U[3 * i + 2] * d2;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = fmaf(a, V[3 * j + 0], fmaf(b, V[3 * j + 1], c * V[3 * j + 2]));
    }
}

// K2L per-sample set-up in fp64: mode R* = U V^T (rounded to fp32) and T = s1 + s2 + s3 (proper,
// signed) as a double.  The Laplace NLL subtracts <A,R> from T (src/laplace/rotation_laplace.py:
// 161-170); carrying T and the ground-truth term in fp64 keeps that cancellation exact to fp32
// rounding for the cost of one 3x3 Jacobi per sample (the grid loop is ~5,000x more work).
SUHPE_HD bool laplace_setup(const float* A, float* Rs, double* T) {
    double U[9], V[9], s[3];
#if SUHPE_K2L_SETUP_F64_ONLY
    const bool ok = proper_svd3_t<double>(A, U, V, s);
#else
    // the fp32 iteration (K1's) finds V to fp32 rounding at a fraction of the cost of fp64 sweeps (two sqrt and two
    // divisions per rotation); kJacobiPolishF64 fp64 sweeps from there converge quadratically: 1e-7 -> 1e-14 -> 1e-28
    float Uf[9], Vf[9], sf[3];
    const bool ok = proper_svd3_t<float>(A, Uf, Vf, sf);
    proper_svd3_t<double>(A, U, V, s, Vf, kJacobiPolishF64);
#endif
    *T = s[0] + s[1] + s[2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Rs[3 * i + j] = (float)fma(U[3 * i], V[3 * j], fma(U[3 * i + 1], V[3 * j + 1], U[3 * i + 2] * V[3 * j + 2]));
    return ok;
}

// ----------------------------------------------------------------------------
// Matrix-Fisher quadrature.
//
// I0e(a) = exp(-a) I0(a), a >= 0, by the reference's two polynomials:
//   a <= 3.75 : P6((a/3.75)^2) * exp(-a)          (A&S 9.8.1)
//   a >  3.75 : P8(3.75/a) / sqrt(a)              (A&S 9.8.2)
// The powers of 3.75 are folded into the coefficients (poly in a^2 resp. 1/a),
// and exp(-a) of the small branch is NOT applied here: the caller merges it
// into the one exponential per node.
// ----------------------------------------------------------------------------
constexpr float kBesselSwitch = 3.75f;

constexpr double cpow(double b, int n) { double r = 1.0; for (int i = 0; i < n; ++i) r *= b; return r; }
// reference coefficient (a Python float stored into an fp32 tensor) folded with 3.75^k in double
constexpr float fold_small(double coef, int k) { return (float)((double)(float)coef / cpow(3.75, k)); }
constexpr float fold_large(double coef, int k) { return (float)((double)(float)coef * cpow(3.75, k)); }

constexpr float kSm6 = fold_small(0.45813e-2, 12), kSm5 = fold_small(0.360768e-1, 10),
                kSm4 = fold_small(0.2659732, 8),   kSm3 = fold_small(1.2067492, 6),
                kSm2 = fold_small(3.0899424, 4),   kSm1 = fold_small(3.5156229, 2);
constexpr float kLg8 = fold_large(0.392377e-2, 8),   kLg7 = fold_large(-0.1647633e-1, 7),
                kLg6 = fold_large(0.2635537e-1, 6),  kLg5 = fold_large(-0.2057706e-1, 5),
                kLg4 = fold_large(0.916281e-2, 4),   kLg3 = fold_large(-0.157565e-2, 3),
                kLg2 = fold_large(0.225319e-2, 2),   kLg1 = fold_large(0.1328592e-1, 1),
                kLg0 = fold_large(0.39894228, 0);

SUHPE_HD float i0_small_poly(float a) {
    // sum_i A_i (a/3.75)^(2i) as a polynomial in q = a^2
    const float q = a * a;
    float p = kSm6;
    p = fmaf(p, q, kSm5);
    p = fmaf(p, q, kSm4);
    p = fmaf(p, q, kSm3);
    p = fmaf(p, q, kSm2);
    p = fmaf(p, q, kSm1);
    p = fmaf(p, q, 1.0f);
    return p;
}

// returns P8(3.75/a)/sqrt(a) as a polynomial in r = 1/a; one MUFU (rsqrt), r = rsqrt^2
SUHPE_HD float i0e_large(float a) {
    const float rs = mufu_rsqrt(a);
    const float r = rs * rs;
    float p = kLg8;
    p = fmaf(p, r, kLg7);
    p = fmaf(p, r, kLg6);
    p = fmaf(p, r, kLg5);
    p = fmaf(p, r, kLg4);
    p = fmaf(p, r, kLg3);
    p = fmaf(p, r, kLg2);
    p = fmaf(p, r, kLg1);
    p = fmaf(p, r, kLg0);
    return p * rs;
}

constexpr float kLog2e = 1.4426950408889634f;

// quadrature node i of 512 on [-1,1], rounded exactly like the reference:
// fl(fl(i * fl(2/511)) - 1)   (src/fisher/torch_norm_factor.py:25-26)
SUHPE_HD float quad_node(float i_as_float) {
    return add_rn(mul_rn(i_as_float, (float)(2.0 / 511.0)), -1.0f);
}

// ----------------------------------------------------------------------------
// Quadrature nodes and per-node constants.  x_i is rounded exactly like the
// reference; u = 1-x, v = 1+x (both exact roundings the reference also performs,
// and u + v == 2 holds exactly for all 512 nodes).  The reciprocal and the
// half-log columns let the large-argument Bessel branch run without MUFU:
//   P8(3.75/a)/sqrt(a), a = f*u  ==  P8'( (1/f)*(1/u) ) * rsqrt(f) * 2^(-1/2 log2 u)
// 1/u comes from the table, 1/f and rsqrt(f) are per-sample constants, and the
// node's 2^(-1/2 log2 u) is folded into the one exponential every node needs.
// ----------------------------------------------------------------------------
constexpr int kQuadNodes = 512;

struct NodeVals { float u, v, iu, iv, Lu, Lv; };

SUHPE_HD NodeVals node_vals(int i) {
    NodeVals n;
    if (i >= kQuadNodes) {            // padding behind the last node (results are discarded)
        n.u = n.v = n.iu = n.iv = 1.0f; n.Lu = n.Lv = 0.0f;
        return n;
    }
    const float x = quad_node((float)i);
    n.u = add_rn(1.0f, -x);
    n.v = add_rn(1.0f, x);
    n.iu = div_rn(1.0f, n.u);         // +inf at node 511 (u == 0): that node is never large-d
    n.iv = div_rn(1.0f, n.v);         // +inf at node 0   (v == 0): that node is never large-s
    n.Lu = -0.5f * log2f(n.u);
    n.Lv = -0.5f * log2f(n.v);
    return n;
}

// ----------------------------------------------------------------------------
// One family of the three Bessel-product integrands:
//   y(x) = I0e(fd (1-x)) * I0e(fs (1+x)) * exp(c (x-1))
// fam 0 (normaliser & d/ds1): fd=(s2-s3)/2 fs=(s2+s3)/2 c=s1+s3
// fam 1 (d/ds2)             : fd=(s1-s3)/2 fs=(s1+s3)/2 c=s2+s3
// fam 2 (d/ds3)             : fd=(s1-s2)/2 fs=(s1+s2)/2 c=s2+s3
// (src/fisher/torch_norm_factor.py:33-63 with the cyclic shifts of :85-87).
//
// a_d = fd*u falls and a_s = fs*v rises with the node index, so each Bessel factor
// switches polynomial (at 3.75) at most once: nodes [0,id) are d-LARGE, [id,512)
// d-small; nodes [0,js) are s-small, [js,512) s-LARGE.  With b0 = min(id,js) and
// b1 = max(id,js) the 512 nodes split into three runs of uniform type
//   [0,b0)  LS  (d large, s small)
//   [b0,b1) LL if js < id, SS if id < js
//   [b1,512) SL (d small, s large)
// and inside a run every node executes the same instruction sequence:
//   LL  y = P8(ifd*iu) P8(ifs*iv) 2^(k1L*u + Lu+Lv)                 * rsqrt(fd fs)
//   LS  y = P8(ifd*iu) P6((fs*v)^2) 2^(k1L*u + Lu + k2*v)           * rsqrt(fd)
//   SL  y = P6((fd*u)^2) P8(ifs*iv) 2^(k1S*u + Lv)                  * rsqrt(fs)
//   SS  y = P6((fd*u)^2) P6((fs*v)^2) 2^(k1S*u + k2*v)
// k1L = -c log2e, k1S = -(c+fd) log2e, k2 = -fs log2e: the exp(-a) of the small
// branch (reference: poly/exp(|a|)) and exp(-c u) share ONE ex2 per node, and the
// run's constant factor multiplies the run's partial sums once.
// ----------------------------------------------------------------------------
enum NodeType { kLS = 0, kLL = 1, kSS = 2, kSL = 3 };

struct FamilyDesc {
    float fd, fs, ifd, ifs;     // Bessel argument scales and their reciprocals
    float k1L, k1S, k2;         // log2-domain exponent slopes
    float scLS, scMid, scSL;    // constant factor of each run
    int b0, b1;                 // run boundaries (node indices in [0,512])
    int mid;                    // kLL or kSS
    int cut;                    // nodes [0,cut) are provably negligible and skipped (0 = evaluate all)
};

// P8 in r = 1/a (3.75^k folded into the coefficients): sqrt(a) * I0e(a) for a > 3.75
SUHPE_HD float large_poly(float r) {
    float p = kLg8;
    p = fmaf(p, r, kLg7); p = fmaf(p, r, kLg6); p = fmaf(p, r, kLg5); p = fmaf(p, r, kLg4);
    p = fmaf(p, r, kLg3); p = fmaf(p, r, kLg2); p = fmaf(p, r, kLg1); p = fmaf(p, r, kLg0);
    return p;
}

// first node index whose argument f*t[i] is on the far side of the 3.75 switch.
//   falling=true : t decreasing (u table); returns first i with f*t[i] <= 3.75
//   falling=false: t increasing (v table); returns first i with f*t[i] >  3.75
// The estimate from the node spacing is corrected against the exact fp32 products the
// reference compares (src/fisher/torch_norm_factor.py:15-18), so the split is bit-faithful.
// (the node columns are read from the kernel's pair-interleaved table: u at t4[4*(i/2) + i%2], v two floats on)
SUHPE_HD float tab_at(const float* t4, int i) { return t4[((i >> 1) << 2) + (i & 1)]; }

SUHPE_HD int switch_index(float f, const float* t, bool falling) {
    const float q = div_rn(kBesselSwitch, f);                     // f = 0 -> +inf
    const float est = falling ? (2.0f - q) * 255.5f : q * 255.5f;
    int i = (int)fminf(fmaxf(ceilf(est), 0.0f), (float)kQuadNodes);   // NaN -> 0
    if (falling) {
        while (i > 0 && mul_rn(f, tab_at(t, i - 1)) <= kBesselSwitch) --i;
        while (i < kQuadNodes && mul_rn(f, tab_at(t, i)) > kBesselSwitch) ++i;
    } else {
        while (i > 0 && mul_rn(f, tab_at(t, i - 1)) > kBesselSwitch) --i;
        while (i < kQuadNodes && mul_rn(f, tab_at(t, i)) <= kBesselSwitch) ++i;
    }
    return i;
}

// Negligible-node cut.  A prefix [0,cut) of a family's nodes is skipped when its whole trapezoid
// mass is provably below 2^-bits of the normaliser sum F (bits = 26: an eighth of an fp32 ulp of F,
// i.e. below what the reference's own fp32 torch.sum resolves).  u falls with the node index, so
// "negligible" is a prefix.  Both sides of the comparison are bounded rigorously:
//
//   lower bound of F (family 0: fd0 = (s2-s3)/2, 2 fs0 = s2+s3, c0 = s1+s3).  I0e falls with its
//   argument, so on the K+1 nodes next to x = 1 (u_k = k h, h = 2/511, k <= K)
//       y0(u_k) >= I0e(s2+s3) I0e(fd0 K h) e^(-c0 h k)
//   and with the trapezoid's half weight at k = 0
//       F >= I0e_lo(s2+s3) I0e_lo(fd0 K h) (G_K - 1/2),  G_K = sum_{k<=K} e^(-c0 h k),
//   I0e_lo(a) = max(e^-a, 0.39 / sqrt(max(a,1))) <= I0e(a);  K = floor(1/(c0 h)) (one e-folding), capped at 510;
//   a second group K < k <= K2 = 3K+2 adds I0e_lo(s2+s3) I0e_lo(fd0 K2 h) (G_K2 - G_K) the same way.
//
//   upper bound of a prefix {i : u_i >= u*} of a family (fd, fs, c):  with I0e(a) <= min(1, kI0eUp/sqrt(a))
//   (sup_a sqrt(a) I0e(a) = 0.4688), the d factor is <= Bd = I0e_up(fd u*) on the whole prefix; the s
//   factor is <= 1 where v < v0 and <= Bs = I0e_up(fs v0) where v >= v0, for any split v0 of the
//   prefix's v range [0, v* = 2-u*]; and sum e^(-c u_i) over nodes with u_i >= w is a geometric series
//   <= e^(-c w) / (1 - e^(-c h)).  Hence
//       mass(prefix) <= Bd (Bs + e^(-c (v*-v0))) e^(-c u*) / (1 - e^(-c h)).
//
// A candidate u* is found by a few fixed-point steps and then CHECKED against the inequality (the
// bound holds for any u*, however it was found); if the check fails the plain e^(-c u) bound
// (Bd = Bs-term = 1) is used.  tests/test_emul_math.py verifies the claim in float64 over random
// spectra.  bits <= 0 disables the cut.
// The search runs in the log2 domain on single-MUFU approximations (relative error ~1e-7 against the
// 0.03-bit slack added to the threshold); the whole thing costs a thread ~50 instructions per family.
constexpr float kI0eUp = 0.4690f;
constexpr float kLog2I0eUp = -1.0923400f;          // log2(0.4690)
constexpr float kQuadStep = (float)(2.0 / 511.0);

SUHPE_HD float mufu_lg2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return log2f(x);
#endif
}
SUHPE_HD float mufu_rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

// I0e(a) >= max(e^-a, 0.39/sqrt(max(a,1)))   (I0 >= 1 covers small a; sqrt(a) I0e(a) >= 0.3989 for a >= 1)
SUHPE_HD float i0e_lower(float a) { return fmaxf(mufu_ex2(-kLog2e * a), 0.39f * mufu_rsqrt(fmaxf(a, 1.0f))); }
// log2 of min(1, kI0eUp / sqrt(a)) >= log2 I0e(a)
SUHPE_HD float lg2_i0e_upper(float a) { return fminf(0.0f, fmaf(-0.5f, mufu_lg2(fmaxf(a, 1e-30f)), kLog2I0eUp)); }

// log2(2^bits / F_lower) plus slack for the approximate evaluation of the bound itself; +inf = no cut
SUHPE_HD float cut_threshold(const float* s, int bits) {
    if (bits <= 0) return INFINITY;
    const float c0 = s[0] + s[2], fd0 = fabsf(0.5f * (s[1] - s[2])), as0 = fabsf(s[1] + s[2]);
    const float t = c0 * kQuadStep;
    if (!(t >= 0.0f) || !(fabsf(s[0]) <= 3.0e38f) || !(as0 <= 3.0e38f) || !(fd0 <= 3.0e38f)) return INFINITY;   // NaN / inf spectra: evaluate everything
    // two node groups next to x = 1: k <= K (one e-folding) and K < k <= K2 (two more)
    float sum = 0.5f;                                                  // t >= 1: the half-weighted end node alone
    if (t < 1.0f) {
        const float K = fminf(floorf(mufu_rcp_approx(fmaxf(t, 1e-9f)) * 0.999f), 510.0f);
        const float K2 = fminf(3.0f * K + 2.0f, 510.0f);
        const float tl = -kLog2e * t;
        float G1, G2;                                                  // sum_{k<=K} e^(-t k),  sum_{K<k<=K2} e^(-t k)
        if (t > 1e-3f) {                                               // 1 - e^-t well conditioned: exact geometric sums
            const float q = mufu_rcp_approx(1.0f - mufu_ex2(tl)), eK = mufu_ex2(tl * (K + 1.0f));
            G1 = (1.0f - eK) * q;
            G2 = eK * (1.0f - mufu_ex2(tl * (K2 - K))) * q;
        } else {                                                       // else the smallest term times the count
            G1 = (K + 1.0f) * mufu_ex2(tl * K);
            G2 = (K2 - K) * mufu_ex2(tl * K2);
        }
        sum = i0e_lower(fd0 * K * kQuadStep) * (G1 - 0.5f) + i0e_lower(fd0 * K2 * kQuadStep) * G2;
    }
    const float Flow = i0e_lower(as0) * sum;
    return (float)bits - mufu_lg2(Flow) + 0.03f;
}

// log2 of the prefix bound without its e^(-c u) / (1 - e^(-c h)) part: log2(Bd (Bs + e^(-c v0))), capped at 0
SUHPE_HD float cut_prefix_factor(float fd, float fs, float cl, float icl, float u) {
    // split of the s factor: v0 = the later of half the prefix and four e-foldings (5.77 octaves of the
    // log2-domain slope cl) before its end -- the prefix mass sits at its end
    const float v = fmaxf(2.0f - u, 0.0f);
    const float v0 = fmaxf(0.5f * v, v - 5.77f * icl);
    const float bs = fminf(1.0f, kI0eUp * mufu_rsqrt(fmaxf(fs * v0, 1e-30f))) + mufu_ex2(-cl * (v - v0));
    return lg2_i0e_upper(fd * u) + fminf(0.0f, mufu_lg2(bs));
}

// cl = c log2(e)
SUHPE_HD int cut_index(float fd, float fs, float cl, float thr) {
    if (!(cl > 0.0f) || !(thr < INFINITY)) return 0;
    const float rhs = thr - mufu_lg2(1.0f - mufu_ex2(-cl * kQuadStep));          // c h -> 0: +inf -> no cut
    if (!(rhs < INFINITY)) return 0;
    const float ic = mufu_rcp_approx(cl);
    float u = rhs * ic;                                               // plain bound: always valid
    if (u < 2.0f) {
        float w = u;
#pragma unroll 1
        for (int it = 0; it < 3; ++it) w = fmaxf((rhs + cut_prefix_factor(fd, fs, cl, ic, w)) * ic, 0.0f);
        w += 0.5f * kQuadStep;                                        // safety margin before the check
        if (fmaf(cl, w, -cut_prefix_factor(fd, fs, cl, ic, w)) >= rhs) u = w;
    }
    // nodes with u_i >= u:  i <= (2 - u) * 255.5 ; one more node of slack
    const float lim = (2.0f - u) * 255.5f - 1.0f;
    return (int)fminf(fmaxf(floorf(lim), 0.0f), (float)(kQuadNodes - 1));   // NaN -> 0
}

// (lo, hi, c) -> descriptor;  utab/vtab: the 512 node values u_i, v_i in the interleaved layout of tab_at()
SUHPE_HD FamilyDesc make_family(float lo, float hi, float c, const float* utab, const float* vtab, float cut_thr) {
    FamilyDesc d;
    d.fd = fabsf(0.5f * (hi - lo));
    d.fs = fabsf(0.5f * (hi + lo));
    d.ifd = div_rn(1.0f, d.fd);
    d.ifs = div_rn(1.0f, d.fs);
    const float cl = c * kLog2e;
    d.k1L = -cl;
    d.k1S = -(cl + d.fd * kLog2e);
    d.k2 = -(d.fs * kLog2e);
    const int id = switch_index(d.fd, utab, true);
    const int js = switch_index(d.fs, vtab, false);
    d.b0 = id < js ? id : js;
    d.b1 = id < js ? js : id;
    d.mid = (js < id) ? kLL : kSS;
    d.scLS = div_rn(1.0f, sqrt_rn(d.fd));
    d.scSL = div_rn(1.0f, sqrt_rn(d.fs));
    d.scMid = (d.mid == kLL) ? d.scLS * d.scSL : 1.0f;
    d.cut = cut_index(d.fd, d.fs, cl, cut_thr) & ~1;     // even: runs start on a node pair
    return d;
}

// family k of a sample with proper singular values s (cut_thr = cut_threshold(s, bits))
SUHPE_HD FamilyDesc family_of(const float* s, int k, const float* utab, const float* vtab, float cut_thr) {
    const float lo = (k == 2) ? s[1] : s[2];
    const float hi = (k == 0) ? s[1] : s[0];
    const float c = (k == 0) ? s[0] + s[2] : s[1] + s[2];
    return make_family(lo, hi, c, utab, vtab, cut_thr);
}
SUHPE_HD void fisher_families(const float* s, const float* utab, const float* vtab, int cut_bits, FamilyDesc* f) {
    const float thr = cut_threshold(s, cut_bits);
#pragma unroll
    for (int k = 0; k < 3; ++k) f[k] = family_of(s, k, utab, vtab, thr);
}

// ----------------------------------------------------------------------------
// Run plan.  The warp that replays the quadrature works on node PAIRS (2m, 2m+1) in the halves of
// packed registers, so the thread that owns a sample turns each family into
//   * up to three pair-aligned runs [start, end), start and end even, every pair inside a run is
//     fully valid and of the run's type -- so the replay needs one predicate per pair and no
//     per-node masks;
//   * up to four single "edge" nodes: where a type boundary b0 / b1 is odd, the pair (b-1, b)
//     straddles two types; the owning thread evaluates its two nodes itself (scalar, node_typed)
//     and adds them to the family's sums.  edge[0] = b0-1 as LS, edge[1] = b0 as mid,
//     edge[2] = b1-1 as mid, edge[3] = b1 as SL; -1 = none (boundary even, node cut away, or the
//     run it belongs to is empty).
// A run is packed into one word that the warp decodes with one mask and one shift:
//   bit  1     the middle run is LL (else SS)
//   bits 4-12  16*m0: byte offset of the run's first node pair (start >> 1) in a float4 table column
//   bits 16-24 the number of node pairs in the run (0 = empty run); passes of 64 pairs while more
//              than 32 remain, then one of 32
// ----------------------------------------------------------------------------
struct FamilyPlan { uint32_t word[3]; int edge[4]; };

SUHPE_HD uint32_t run_word(int start, int end, bool mid_ll) {      // start, end even
    const uint32_t flag = mid_ll ? 2u : 0u;
    if (end <= start) return flag;
    return flag | ((uint32_t)(start >> 1) << 4) | ((uint32_t)((end - start) >> 1) << 16);
}
SUHPE_HD bool run_word_mid_ll(uint32_t w) { return (w & 2u) != 0; }
SUHPE_HD unsigned run_word_m0(uint32_t w) { return (w >> 4) & 511u; }
SUHPE_HD int run_word_pairs(uint32_t w) { return (int)(w >> 16); }

SUHPE_HD void family_plan(const FamilyDesc& d, FamilyPlan& p) {
    const int cut = d.cut, b0 = d.b0, b1 = d.b1;                   // cut even, 0 <= b0 <= b1 <= 512
    const int up0 = (b0 + 1) & ~1, up1 = (b1 + 1) & ~1;
    const int s_mid = up0 > cut ? up0 : cut, s_sl = up1 > cut ? up1 : cut;
    p.word[0] = run_word(cut, b0 & ~1, false);
    p.word[1] = run_word(s_mid, b1 & ~1, d.mid == kLL);
    p.word[2] = run_word(s_sl, kQuadNodes, false);
    const bool odd0 = (b0 & 1) != 0, odd1 = (b1 & 1) != 0;
    const int lo_mid = b0 > cut ? b0 : cut;
    p.edge[0] = (odd0 && b0 - 1 >= cut) ? b0 - 1 : -1;             // last node of the LS run
    p.edge[1] = (odd0 && b0 >= cut && b0 < b1) ? b0 : -1;          // first node of the middle run
    p.edge[2] = (odd1 && b1 - 1 >= lo_mid) ? b1 - 1 : -1;          // last node of the middle run
    p.edge[3] = (odd1 && b1 >= cut) ? b1 : -1;                     // first node of the SL run (b1 odd => b1 <= 511)
}
SUHPE_HD int edge_type(const FamilyDesc& d, int k) { return k == 0 ? kLS : (k == 3 ? kSL : d.mid); }

SUHPE_HD int node_type(const FamilyDesc& d, int i) {
    return i < d.b0 ? kLS : (i < d.b1 ? d.mid : kSL);
}
SUHPE_HD float type_scale(const FamilyDesc& d, int type) {
    return type == kLS ? d.scLS : (type == kSL ? d.scSL : (type == kLL ? d.scLS * d.scSL : 1.0f));
}

// Scalar node evaluation WITHOUT the run's constant factor: the same operations, in the
// same order, as the packed f32x2 run bodies of the kernel (trapezoid end-point
// corrections and the host emulation use this).
SUHPE_HD float node_typed(const FamilyDesc& d, int type, const NodeVals& n) {
    float pd, ps, e;
    if (type == kLL) {
        pd = large_poly(d.ifd * n.iu);
        ps = large_poly(d.ifs * n.iv);
        e = fmaf(d.k1L, n.u, n.Lu + n.Lv);
    } else if (type == kLS) {
        pd = large_poly(d.ifd * n.iu);
        ps = i0_small_poly(d.fs * n.v);
        e = fmaf(d.k2, n.v, fmaf(d.k1L, n.u, n.Lu));
    } else if (type == kSL) {
        pd = i0_small_poly(d.fd * n.u);
        ps = large_poly(d.ifs * n.iv);
        e = fmaf(d.k1S, n.u, n.Lv);
    } else {
        pd = i0_small_poly(d.fd * n.u);
        ps = i0_small_poly(d.fs * n.v);
        e = fmaf(d.k2, n.v, d.k1S * n.u);
    }
    return (pd * ps) * mufu_ex2(e);
}

// Closing arithmetic once the four trapezoid sums are known.
//   F  = sum_i w_i y0(x_i)        N0 = sum_i w_i x_i y0(x_i)
//   N1 = sum_i w_i x_i y1(x_i)    N2 = sum_i w_i x_i y2(x_i)
// f = 1/2 * (F * 2/511),  g_j = (1/2 * N_j * 2/511) / f   (torch_norm_factor.py:31,73,87-88)
struct FisherStats { float logf, logC, g[3], entropy; };

SUHPE_HD FisherStats fisher_finish(const float* s, float F, float N0, float N1, float N2) {
    FisherStats o;
    const float f = div_rn(F, 511.0f);
    o.g[0] = div_rn(div_rn(N0, 511.0f), f);
    o.g[1] = div_rn(div_rn(N1, 511.0f), f);
    o.g[2] = div_rn(div_rn(N2, 511.0f), f);
    o.logf = logf(f);
    o.logC = o.logf + (s[0] + s[1] + s[2]);
    // H = log f + sum_j s_j (1 - g_j): algebraic collapse of the reference's
    // Fisher->Bingham->autograd chain (src/fisher/fisher_utils.py:70-81)
    o.entropy = o.logf + fmaf(s[0], 1.0f - o.g[0], fmaf(s[1], 1.0f - o.g[1], s[2] * (1.0f - o.g[2])));
    return o;
}

// ----------------------------------------------------------------------------
// Matrix-Fisher cross entropy through the Bingham form (src/fisher/fisher_utils.py:84-99,
// src/fisher/bingham_utils.py:5-32, src/fisher/between_bingham_fisher.py:107-152), target A1,
// prediction A2, value and gradient w.r.t. A2 in closed form.
//
// With the proper SVDs, g = grad logC (K2), the quaternion frames
//   VB = [q(U V^T), q(U E_1 V^T), q(U E_2 V^T), q(U E_3 V^T)],  q(U E_k V^T) = (0,u_k) (x) q(U V^T)
// (the Bingham-convention column order for s1 >= s2 >= |s3|), W = VB1^T VB2, the expected squared
// projections of the target  gamma = 1/4 (1+g1+g2+g3, 1+g1-g2-g3, 1-g1+g2-g3, 1-g1-g2+g3)  and
// LamB2 = -2 (s2+s3, s1+s3, s1+s2) of the prediction, the reference evaluates
//   CE = log f(s_2) - sum_{i=1..3} LamB2_i [ gamma_0 W_0i^2 + sum_{j=1..3} gamma_j W_ij^2 ]
// (W_ij, not W_ji: bingham_utils.py:27 takes a ROW of V1^T V2 where the expectation needs a
// column; reproduced as is).  x^T K(A) x = <A, R(x)> is linear in A and (lam_c, v_c) are the
// eigenpairs of K(A2), so by first-order perturbation
//   dCE/dA2 = sum_c L_lam_c R~(v_c,v_c) + sum_{k != c} (v_k . L_v_c)/(lam_c - lam_k) R~(v_k,v_c)
// with R~ the symmetric bilinear form of the quaternion -> rotation map.  This is what autograd
// yields through torch.svd, matrix_to_quaternion and the quadrature's backward (oracle:
// fisher_ce vs fisher_ce_closed_form agree to 6e-15 in fp64); quaternion signs cancel everywhere.
// ----------------------------------------------------------------------------
// unit quaternion (w,x,y,z) of a rotation matrix, largest-component branch
SUHPE_HD void quat_of_rotation(const float* R, float* q) {
    const float t0 = 1.0f + R[0] + R[4] + R[8], t1 = 1.0f + R[0] - R[4] - R[8];
    const float t2 = 1.0f - R[0] + R[4] - R[8], t3 = 1.0f - R[0] - R[4] + R[8];
    const float tm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
    if (tm == t0)      { q[0] = t0; q[1] = R[7] - R[5]; q[2] = R[2] - R[6]; q[3] = R[3] - R[1]; }
    else if (tm == t1) { q[0] = R[7] - R[5]; q[1] = t1; q[2] = R[3] + R[1]; q[3] = R[2] + R[6]; }
    else if (tm == t2) { q[0] = R[2] - R[6]; q[1] = R[3] + R[1]; q[2] = t2; q[3] = R[5] + R[7]; }
    else               { q[0] = R[3] - R[1]; q[1] = R[6] + R[2]; q[2] = R[7] + R[5]; q[3] = t3; }
    const float inv = div_rn(1.0f, sqrt_rn(fmaf(q[0], q[0], fmaf(q[1], q[1], fmaf(q[2], q[2], q[3] * q[3])))));
    q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}

// frame[c][0..3]: quaternion c of the Bingham frame of (U, V)
SUHPE_HD void bingham_frame(const float* U, const float* V, float (*frame)[4]) {
    float R0[9];
    u_diag_vt(U, V, 1.f, 1.f, 1.f, R0);
    quat_of_rotation(R0, frame[0]);
    const float w = frame[0][0], x = frame[0][1], y = frame[0][2], z = frame[0][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a0 = U[k], a1 = U[3 + k], a2 = U[6 + k];          // column k of U
        frame[k + 1][0] = -(a0 * x + a1 * y + a2 * z);                 // (0,a) (x) (w,v) = (-a.v, w a + a x v)
        frame[k + 1][1] = fmaf(w, a0, a1 * z - a2 * y);
        frame[k + 1][2] = fmaf(w, a1, a2 * x - a0 * z);
        frame[k + 1][3] = fmaf(w, a2, a0 * y - a1 * x);
    }
}

// acc += coef * R~(p, q)
SUHPE_HD void polar_rotation_axpy(float coef, const float* p, const float* q, float* acc) {
    const float ww = p[0] * q[0], xx = p[1] * q[1], yy = p[2] * q[2], zz = p[3] * q[3];
    const float xy = fmaf(p[1], q[2], p[2] * q[1]), xz = fmaf(p[1], q[3], p[3] * q[1]), yz = fmaf(p[2], q[3], p[3] * q[2]);
    const float wx = fmaf(p[0], q[1], p[1] * q[0]), wy = fmaf(p[0], q[2], p[2] * q[0]), wz = fmaf(p[0], q[3], p[3] * q[0]);
    acc[0] = fmaf(coef, (ww + xx) - (yy + zz), acc[0]);
    acc[1] = fmaf(coef, xy - wz, acc[1]);
    acc[2] = fmaf(coef, xz + wy, acc[2]);
    acc[3] = fmaf(coef, xy + wz, acc[3]);
    acc[4] = fmaf(coef, (ww + yy) - (xx + zz), acc[4]);
    acc[5] = fmaf(coef, yz - wx, acc[5]);
    acc[6] = fmaf(coef, xz - wy, acc[6]);
    acc[7] = fmaf(coef, yz + wx, acc[7]);
    acc[8] = fmaf(coef, (ww + zz) - (xx + yy), acc[8]);
}

// g1: grad logC of the target; s2, g2, logf2 = log f(s2): of the prediction.  grad may be null.
SUHPE_HD float fisher_ce_close(const float* U1, const float* V1, const float* g1,
                               const float* U2, const float* V2, const float* s2, const float* g2, float logf2,
                               float* grad) {
    float f1[4][4], f2[4][4], W[4][4];
    bingham_frame(U1, V1, f1);
    bingham_frame(U2, V2, f2);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            W[r][c] = fmaf(f1[r][0], f2[c][0], fmaf(f1[r][1], f2[c][1], fmaf(f1[r][2], f2[c][2], f1[r][3] * f2[c][3])));
    const float gam[4] = {0.25f * (1.0f + g1[0] + g1[1] + g1[2]), 0.25f * (1.0f + g1[0] - g1[1] - g1[2]),
                          0.25f * (1.0f - g1[0] + g1[1] - g1[2]), 0.25f * (1.0f - g1[0] - g1[1] + g1[2])};
    const float lamB[4] = {0.0f, -2.0f * (s2[1] + s2[2]), -2.0f * (s2[0] + s2[2]), -2.0f * (s2[0] + s2[1])};
    float phi[4] = {0.f, 0.f, 0.f, 0.f};
    float second = 0.0f;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        phi[i] = fmaf(gam[0], W[0][i] * W[0][i],
                      fmaf(gam[1], W[i][1] * W[i][1], fmaf(gam[2], W[i][2] * W[i][2], gam[3] * W[i][3] * W[i][3])));
        second = fmaf(lamB[i], phi[i], second);
    }
    const float ce = logf2 - second;
    if (grad) {
        // eigenvalues of K(A2) in frame order; d s_m / d lam_c = 1/4 [[1,1,-1,-1],[1,-1,1,-1],[1,-1,-1,1]]
        const float lam[4] = {s2[0] + s2[1] + s2[2], s2[0] - s2[1] - s2[2], -s2[0] + s2[1] - s2[2], -s2[0] - s2[1] + s2[2]};
        const float h0 = 0.25f * (g2[0] - 1.0f), h1 = 0.25f * (g2[1] - 1.0f), h2 = 0.25f * (g2[2] - 1.0f);
        const float Ll[4] = {(h0 + h1 + h2) + (phi[1] + phi[2] + phi[3]), (h0 - h1 - h2) - phi[1],
                             (-h0 + h1 - h2) - phi[2], (-h0 - h1 + h2) - phi[3]};
#pragma unroll
        for (int k = 0; k < 9; ++k) grad[k] = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) polar_rotation_axpy(Ll[c], f2[c], f2[c], grad);
        // off-diagonal terms: G_kc = -D_kc / (lam_c - lam_k), D_kc = v2_k . d(second)/d v2_c (zero for c = 0);
        // R~ is symmetric in (k,c), so each unordered pair gets G_kc + G_ck
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = k + 1; c < 4; ++c) {
                float Dkc = 2.0f * gam[0] * lamB[c] * W[0][c] * W[0][k];
                float Dck = (k == 0) ? 0.0f : 2.0f * gam[0] * lamB[k] * W[0][k] * W[0][c];
                float a = 0.0f;
#pragma unroll
                for (int i = 1; i < 4; ++i) a = fmaf(lamB[i], W[i][c] * W[i][k], a);
                Dkc = fmaf(2.0f * gam[c], a, Dkc);
                if (k != 0) Dck = fmaf(2.0f * gam[k], a, Dck);
                const float den = lam[c] - lam[k];
                // exact degeneracy (s_i = +-s_j): the reference's svd backward is 0/0 there; contribute nothing
                const float G = (den != 0.0f) ? div_rn(Dck - Dkc, den) : 0.0f;   // -Dkc/den - Dck/(-den)
                polar_rotation_axpy(G, f2[k], f2[c], grad);
            }
    }
    return ce;
}

// ----------------------------------------------------------------------------
// Rotation-Laplace grid sum (src/laplace/rotation_laplace.py:58-72,140-173 and its autograd
// backward).  Per (sample, grid point k):
//   d_k = T - <A,R_k>,  q_k = sqrt(max(d_k, 1e-8)),  w_k = exp(-q_k)/q_k
//   Z = sum w_k,  C = sum w_k (1/q_k + 1/q_k^2) [live points],  M = sum (same weight) R_k
// kept relative to the running minimum of q (the reference subtracts c = max_k(-q_k) first).
// ----------------------------------------------------------------------------
constexpr float kLapEps = 1e-8f;           // rotation_laplace.py:11

// Two-level sums: points add into the block sums (z, c, m); laplace_accum_flush() folds a block
// into the totals (Z, C, M).  The thread-per-sample decomposition flushes every 128 points, so its
// fp32 summation error matches the warp-per-sample one (32 lane partials) instead of growing with
// the 4608-point grid -- the gradient subtracts sum(chat_k) R* from sum(chat_k R_k) and amplifies it.
struct LaplaceAccum { float qmin, Z, C, M[9], z, c, m[9]; };

SUHPE_HD void laplace_accum_init(LaplaceAccum& a) {
    a.qmin = INFINITY; a.Z = a.C = a.z = a.c = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) a.M[i] = a.m[i] = 0.f;
}

// sqrt with one Newton step on top of MUFU.RSQ: q ~ sqrt(d), rs ~ 1/sqrt(d)
SUHPE_HD void sqrt_pair(float d, float& q, float& rs) {
    rs = mufu_rsqrt(d);
    q = d * rs;
    const float err = fmaf(-q, q, d);
    q = fmaf(0.5f * rs, err, q);
}

SUHPE_HD void laplace_accum_scale(LaplaceAccum& a, float sc) {
    a.Z *= sc; a.C *= sc; a.z *= sc; a.c *= sc;
#pragma unroll
    for (int i = 0; i < 9; ++i) { a.M[i] *= sc; a.m[i] *= sc; }
}

SUHPE_HD void laplace_accum_point(LaplaceAccum& a, const float* A, float T, const float* r) {
    float t = fmaf(A[0], r[0], -T);          // -T rides in the first FMA: t = <A,R_k> - T
#pragma unroll
    for (int i = 1; i < 9; ++i) t = fmaf(A[i], r[i], t);
    const float d = -t;
    const bool live = d >= kLapEps;          // clamp_min passes the gradient where input >= min
    float q, rs;
    sqrt_pair(fmaxf(d, kLapEps), q, rs);
    if (q < a.qmin) {                         // new running maximum of p = -q: rescale (rare)
        laplace_accum_scale(a, mufu_ex2((q - a.qmin) * kLog2e));
        a.qmin = q;
    }
    const float w = mufu_ex2((a.qmin - q) * kLog2e) * rs;    // exp(p - c) / (-p)
    a.z += w;
    const float cw = live ? w * fmaf(rs, rs, rs) : 0.0f;      // w (1/q + 1/q^2); the 1/2 is applied once at the end
    a.c += cw;
#pragma unroll
    for (int i = 0; i < 9; ++i) a.m[i] = fmaf(cw, r[i], a.m[i]);
}

SUHPE_HD void laplace_accum_flush(LaplaceAccum& a) {
    a.Z += a.z; a.C += a.c; a.z = a.c = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { a.M[i] += a.m[i]; a.m[i] = 0.f; }
}

// re-express a (flushed) partial accumulator relative to a smaller or equal minimum qg before merging partials
SUHPE_HD void laplace_accum_rebase(LaplaceAccum& a, float qg) {
    laplace_accum_scale(a, (a.qmin == INFINITY) ? 0.f : mufu_ex2((qg - a.qmin) * kLog2e));
    a.qmin = qg;
}

// nll = logF + q_x + log q_x,  d nll/dA = -sum_k chat_k (R* - R_k) + c_x (R* - R_gt);
// dx = T - <A,R_gt> comes in already cancelled in fp64 (laplace_setup)
SUHPE_HD void laplace_finish(const LaplaceAccum& a, float dx, const float* Rs, const float* Rg, int N,
                             float* nll, float* logF_out, float* grad) {
    // logF = c + log(sum * (1/N)), c = -qmin   (rotation_laplace.py:69-71)
    const float logF = -a.qmin + logf(a.Z * (1.0f / (float)N));
    const float qx = sqrt_rn(fmaxf(dx, kLapEps));
    *nll = logF + qx + logf(qx);
    *logF_out = logF;
    const float invZ = 0.5f / a.Z;
    const float cs = a.C * invZ;                                   // sum_k chat_k
    const float cx = (dx >= kLapEps) ? 0.5f * (1.0f + 1.0f / qx) / qx : 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) grad[i] = fmaf(a.M[i], invZ, fmaf(cx - cs, Rs[i], -cx * Rg[i]));
}

SUHPE_HD float laplace_gt_gap(const float* A, const float* Rg, double T) {
    double tx = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) tx = fma((double)A[i], (double)Rg[i], tx);
    return (float)(T - tx);
}

// ----------------------------------------------------------------------------
// Entropy -> radix key with numpy.sort's total order: ascending, -0 == +0,
// NaN last (src/agent.py:405).
// ----------------------------------------------------------------------------
SUHPE_HD uint32_t entropy_key(float e) {
    if (e != e) return 0xFFFFFFFFu;
    if (e == 0.0f) e = 0.0f;  // -0 -> +0
#if defined(__CUDA_ARCH__)
    uint32_t u = __float_as_uint(e);
#else
    union { float f; uint32_t u; } c; c.f = e; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
SUHPE_HD float key_entropy(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    if (k == 0xFFFFFFFFu) u = 0x7FC00000u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// ----------------------------------------------------------------------------
// Error metrics.
// ----------------------------------------------------------------------------
SUHPE_HD float mufu_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

// a / b for a compile-time constant b with y = RN(1/b): q = a y, r = a - b q (exact in the FMA),
// q' = q + r y (Markstein).  Checked exhaustively on the host against IEEE division for EVERY
// float with 1e-30 <= |a| <= 1e30, for b = fl(pi) and b = 3 (the two divisors of
// src/agent.py:452-454): bit-identical.  Three FMA-pipe instructions instead of the ~10 + slow
// path of __fdiv_rn.
SUHPE_HD float div_by_const(float a, float b, float y) {
    const float q = mul_rn(a, y);
    const float r = fmaf(-b, q, a);
    return fmaf(r, y, q);
}
constexpr float kPiF = 3.14159265358979323846f;
SUHPE_HD float rad_to_deg_ref(float rad) {          // euler * 180 / np.pi in fp32, two roundings like torch
    return div_by_const(mul_rn(rad, 180.0f), kPiF, (float)(1.0 / (double)kPiF));
}

// atan2(y, x), branch-free: t = min/max in [0,1] by MUFU.RCP + one Newton step, atan(t) =
// t + t s P7(s), s = t^2 (weighted minimax fit, |rel err| < 0.3 * 2^-24 before rounding), then
// the octant is unfolded with two-term pi/2 and pi.  Measured against double atan2 on 10^7 random
// rotation entries and the axis / signed-zero cases: <= 2 ulp (tests/test_emul_math.py).  IEEE
// behaviour kept for signed zeros and x = y = 0; infinities are not handled (rotation entries).
SUHPE_HD float atan2_so3(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    // keep 1/mx a normal number over the whole exponent range
    const float sc = (mx < 1e-30f) ? 1.8446744e19f : ((mx > 1e30f) ? 5.4210109e-20f : 1.0f);   // 2^64, 2^-64
    mx *= sc; mn *= sc;
    const float r = mufu_rcp(mx);
    float t = mn * r;
    t = fmaf(fmaf(-t, mx, mn), r, t);
    t = (mx > 0.0f) ? t : 0.0f;                                  // atan2(+-0, +-0)
    const float s = t * t;
    float p = 0.0029504755854872678f;
    p = fmaf(p, s, -0.01648717905513075f);
    p = fmaf(p, s, 0.043405680097467356f);
    p = fmaf(p, s, -0.07568590190914992f);
    p = fmaf(p, s, 0.10673642075124151f);
    p = fmaf(p, s, -0.1421297326642285f);
    p = fmaf(p, s, 0.19994003814484457f);
    p = fmaf(p, s, -0.33333162357779944f);
    float a = fmaf(p * s, t, t);
    if (ay > ax) a = (1.57079637050628662109375f - a) + -4.37113882867379e-8f;       // pi/2 = hi + lo
#if defined(__CUDA_ARCH__)
    const bool xneg = __float_as_int(x) < 0;
#else
    const bool xneg = signbit(x);
#endif
    if (xneg) a = (3.1415927410125732421875f - a) + -8.74227765734758e-8f;           // pi = hi + lo
    a = copysignf(a, y);
    return (x != x || y != y) ? x + y : a;
}

// (pitch, yaw, roll) in radians, src/utils.py:232-260 incl. the arithmetic blend
// with the singular flag taken before the full-range flip.  The singular branch (xs) is only
// evaluated where the flag is set: x*1 + xs*0 == x exactly for every finite xs.
SUHPE_HD void euler_from_rotation(const float* R, bool full_range, float* out) {
    const float r00 = R[0], r10 = R[3];
    float sy = sqrt_rn(add_rn(mul_rn(r00, r00), mul_rn(r10, r10)));
    const bool singular = sy < 1e-6f;
    if (full_range && r00 < 0.0f) sy = -sy;
    const float x = atan2_so3(R[7], R[8]);
    const float y = atan2_so3(-R[6], sy);
    const float z = atan2_so3(r10, r00);
    out[0] = x; out[1] = y; out[2] = z;
    if (singular) {
        const float xs = atan2_so3(-R[5], R[4]);
        const float zs = r10 * 0.0f;
        out[0] = add_rn(mul_rn(x, 0.0f), mul_rn(xs, 1.0f));
        out[1] = add_rn(mul_rn(y, 0.0f), mul_rn(y, 1.0f));
        out[2] = add_rn(mul_rn(z, 0.0f), mul_rn(zs, 1.0f));
    }
}

// limit_angle of DAD-3DHeads (src/utils.py:289-300): wrap degrees into [-180, 180] with the
// reference's int() (truncation) and // (floor) arithmetic.
SUHPE_HD float limit_angle_deg(float a) {
    if (a < -180.0f) a = fmaf(-2.0f * floorf(truncf(a / 180.0f) * 0.5f), 180.0f, a);
    if (a > 180.0f) a = fmaf(-2.0f * floorf((truncf(a / 180.0f) + 1.0f) * 0.5f), 180.0f, a);
    return a;
}

// DAD-trained models (eval.py:66-74, predict.py:84-87, image.py:218-221): the angles are
// scipy's Rotation.from_matrix(R^T).as_euler("xyz", degrees=True) = (a0, a1, a2) with
// R^T = Rz(a2) Ry(a1) Rx(a0), then [roll, pitch, yaw] = limit_angle([a2, a0 - 180, a1]); returned
// in the reference's list order (pitch, yaw, roll), DEGREES.  At gimbal lock (|R02| = 1) scipy
// zeroes the third angle and puts the whole in-plane rotation into the first.
SUHPE_HD void euler_dad_degrees(const float* R, float* out) {
    const float kDeg = 57.29577951308232f;
    const float cy = sqrt_rn(fmaf(R[0], R[0], R[1] * R[1]));         // cos(a1) >= 0
    const float a1 = atan2_so3(-R[2], cy);
    float a0, a2;
    if (cy > 1e-6f) {
        a0 = atan2_so3(R[5], R[8]);
        a2 = atan2_so3(R[1], R[0]);
    } else {                                                          // gimbal lock: a2 := 0
        a0 = atan2_so3(-R[7], R[4]);
        a2 = 0.0f;
    }
    out[0] = limit_angle_deg(fmaf(a0, kDeg, -180.0f));
    out[1] = limit_angle_deg(a1 * kDeg);
    out[2] = limit_angle_deg(a2 * kDeg);
}

// mean_3 | euler * 180 / pi - gt_euler_deg |   (src/agent.py:452-454)
SUHPE_HD float euler_mae_degrees(const float* euler_rad, const float* gt_deg) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc = add_rn(acc, fabsf(rad_to_deg_ref(euler_rad[k]) - gt_deg[k]));
    return div_by_const(acc, 3.0f, (float)(1.0 / 3.0));
}

// trace(Rp Rg^T) = sum_ij Rp_ij Rg_ij
SUHPE_HD float relative_trace(const float* Rp, const float* Rg) {
    float d0 = fmaf(Rp[0], Rg[0], fmaf(Rp[1], Rg[1], Rp[2] * Rg[2]));
    float d1 = fmaf(Rp[3], Rg[3], fmaf(Rp[4], Rg[4], Rp[5] * Rg[5]));
    float d2 = fmaf(Rp[6], Rg[6], fmaf(Rp[7], Rg[7], Rp[8] * Rg[8]));
    return d0 + d1 + d2;
}

// pytorch3d acos_linear_extrapolation with bounds +-(1-1e-4), then degrees.
// ok=false when the trace leaves [-1-1e-4, 3+1e-4] (pytorch3d raises ValueError).
SUHPE_HD float geodesic_degrees(float trace, bool* ok) {
    *ok = !(trace < -1.0f - 1e-4f) && !(trace > 3.0f + 1e-4f);
    const float c = (trace - 1.0f) * 0.5f;
    const float bound = (float)(1.0 - 1e-4);
    const float slope = (float)(-1.0 / 0.014141782065918747);   // -1/sqrt(1-bound^2)
    float ang;
    if (c >= bound)        ang = fmaf(c - bound, slope, (float)0.014142253285269182);   // acos(bound)
    else if (c <= -bound)  ang = fmaf(c + bound, slope, (float)3.1274504003045240);     // acos(-bound)
    else                   ang = acosf(c);
    return ang * (float)(180.0 / 3.14159265358979323846);
}

// ||I - Rp Rg^T||_F   (eval.py:93-98)
SUHPE_HD float frobenius_to_identity(const float* Rp, const float* Rg) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float d = fmaf(Rp[3 * i], Rg[3 * j], fmaf(Rp[3 * i + 1], Rg[3 * j + 1], Rp[3 * i + 2] * Rg[3 * j + 2]));
            const float e = ((i == j) ? 1.0f : 0.0f) - d;
            acc = fmaf(e, e, acc);
        }
    return sqrt_rn(acc);
}

}  // namespace suhpe
