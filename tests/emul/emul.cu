// Host-only instantiation of semiuhpe_b200/csrc/so3_math.cuh for the CPU test suite.
// TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` exercise the per-sample device
// arithmetic (SVD, quadrature nodes, closing formulas, metrics, key transform)
// against the oracle without a GPU.  It is never loaded by the product package.
#include "../../semiuhpe_b200/csrc/so3_math.cuh"
#include <stdint.h>
using namespace suhpe;

extern "C" {

void emul_proper_svd(const float* A, long n, float* R, float* S, float* U, float* V, int* ok) {
    for (long i = 0; i < n; ++i) {
        float u[9], v[9], s[3];
        ok[i] = proper_svd3(A + 9 * i, u, v, s) ? 1 : 0;
        u_diag_vt(u, v, 1.f, 1.f, 1.f, R + 9 * i);
        for (int k = 0; k < 9; ++k) { U[9 * i + k] = u[k]; V[9 * i + k] = v[k]; }
        for (int k = 0; k < 3; ++k) S[3 * i + k] = s[k];
    }
}

// Mirrors the warp decomposition of the kernel: lane l owns nodes l, l+32, ...,
// partial sums per lane, then an xor-butterfly.
static void quadrature(const float* s, float* F, float* N0, float* N1, float* N2) {
    Family fam[3];
    fisher_families(s, fam);
    float pf[32], p0[32], p1[32], p2[32];
    for (int lane = 0; lane < 32; ++lane) {
        float af = 0, a0 = 0, a1 = 0, a2 = 0;
        for (int it = 0; it < 16; ++it) {
            const int node = 32 * it + lane;
            const float x = quad_node((float)node);
            const float u = 1.0f - x, v = 1.0f + x;
            const float w = (node == 0 || node == 511) ? 0.5f : 1.0f;
            const float y0 = fisher_node(fam[0], u, v) * w;
            const float y1 = fisher_node(fam[1], u, v) * w;
            const float y2 = fisher_node(fam[2], u, v) * w;
            af += y0;
            a0 = fmaf(x, y0, a0);
            a1 = fmaf(x, y1, a1);
            a2 = fmaf(x, y2, a2);
        }
        pf[lane] = af; p0[lane] = a0; p1[lane] = a1; p2[lane] = a2;
    }
    for (int off = 16; off >= 1; off >>= 1)
        for (int lane = 0; lane < 32; ++lane) {
            if ((lane & off) == 0) {
                pf[lane] += pf[lane ^ off]; p0[lane] += p0[lane ^ off];
                p1[lane] += p1[lane ^ off]; p2[lane] += p2[lane ^ off];
            }
        }
    *F = pf[0]; *N0 = p0[0]; *N1 = p1[0]; *N2 = p2[0];
}

void emul_fisher(const float* A, const float* Rgt, long n, float overreg,
                 float* nll, float* grad, float* Rout, float* entropy, float* logC, float* S, float* G) {
    for (long i = 0; i < n; ++i) {
        const float* a = A + 9 * i;
        float u[9], v[9], s[3];
        proper_svd3(a, u, v, s);
        float F, N0, N1, N2;
        quadrature(s, &F, &N0, &N1, &N2);
        FisherStats st = fisher_finish(s, F, N0, N1, N2);
        u_diag_vt(u, v, 1.f, 1.f, 1.f, Rout + 9 * i);
        float gm[9];
        u_diag_vt(u, v, st.g[0], st.g[1], st.g[2], gm);
        float dot = 0.f;
        if (Rgt) for (int k = 0; k < 9; ++k) dot = fmaf(a[k], Rgt[9 * i + k], dot);
        nll[i] = fmaf(overreg, st.logC, -dot);
        for (int k = 0; k < 9; ++k) grad[9 * i + k] = fmaf(overreg, gm[k], Rgt ? -Rgt[9 * i + k] : 0.f);
        entropy[i] = st.entropy;
        logC[i] = st.logC;
        for (int k = 0; k < 3; ++k) { S[3 * i + k] = s[k]; G[3 * i + k] = st.g[k]; }
    }
}

void emul_quad_nodes(float* x) { for (int i = 0; i < 512; ++i) x[i] = quad_node((float)i); }

void emul_i0e(const float* a, long n, float* out) {
    for (long i = 0; i < n; ++i) {
        const float t = fabsf(a[i]);
        out[i] = (t <= kBesselSwitch) ? i0_small_poly(t) * expf(-t) : i0e_large(t);
    }
}

void emul_keys(const float* e, long n, uint32_t* k, float* back) {
    for (long i = 0; i < n; ++i) { k[i] = entropy_key(e[i]); back[i] = key_entropy(k[i]); }
}

void emul_metrics(const float* Rp, const float* Rg, const float* gt_euler, long n, int full_range,
                  float* geo, float* frob, float* euler, float* mae, int* ok) {
    for (long i = 0; i < n; ++i) {
        bool good;
        geo[i] = geodesic_degrees(relative_trace(Rp + 9 * i, Rg + 9 * i), &good);
        ok[i] = good;
        frob[i] = frobenius_to_identity(Rp + 9 * i, Rg + 9 * i);
        float e[3];
        euler_from_rotation(Rp + 9 * i, full_range != 0, e);
        for (int k = 0; k < 3; ++k) euler[3 * i + k] = e[k];
        if (gt_euler) mae[i] = euler_mae_degrees(e, gt_euler + 3 * i);
    }
}

}  // extern "C"
