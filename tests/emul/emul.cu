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

}  // extern "C"

// Mirrors the warp decomposition of the kernel: lane l owns, in pair p of 8, the nodes
// 64p+l (lo half) and 64p+32+l (hi half); per family it keeps packed sums of y and u*y,
// adds the halves, and the warp butterfly-reduces {Y0, UY0, N1, N2}.  The trapezoid
// end-point halves are subtracted afterwards.
static float butterfly(float* v) {
    for (int off = 16; off >= 1; off >>= 1)
        for (int lane = 0; lane < 32; ++lane)
            if ((lane & off) == 0) v[lane] += v[lane ^ off];
    return v[0];
}

static void quadrature(const float* s, float* F, float* N0, float* N1, float* N2) {
    Family fam[3];
    fisher_families(s, fam);
    float pY0[32], pUY0[32], pN1[32], pN2[32];
    float x_nodes[512];
    for (int i = 0; i < 512; ++i) x_nodes[i] = quad_node((float)i);
    for (int lane = 0; lane < 32; ++lane) {
        for (int f = 0; f < 3; ++f) {
            float ylo = 0, yhi = 0, ulo = 0, uhi = 0;
            for (int p = 0; p < 8; ++p) {
                const float x0 = x_nodes[64 * p + lane], x1 = x_nodes[64 * p + 32 + lane];
                const float u0 = add_rn(1.0f, -x0), v0 = add_rn(1.0f, x0);
                const float u1 = add_rn(1.0f, -x1), v1 = add_rn(1.0f, x1);
                const float y0 = fisher_node(fam[f], u0, v0), y1 = fisher_node(fam[f], u1, v1);
                ylo += y0; yhi += y1;
                ulo = fmaf(u0, y0, ulo); uhi = fmaf(u1, y1, uhi);
            }
            const float Y = ylo + yhi, UY = ulo + uhi;
            if (f == 0) { pY0[lane] = Y; pUY0[lane] = UY; }
            else if (f == 1) pN1[lane] = Y - UY;
            else pN2[lane] = Y - UY;
        }
    }
    const float Y0 = butterfly(pY0), UY0 = butterfly(pUY0), n1 = butterfly(pN1), n2 = butterfly(pN2);
    const float uf = add_rn(1.0f, -x_nodes[0]), vf = add_rn(1.0f, x_nodes[0]);
    const float ul = add_rn(1.0f, -x_nodes[511]), vl = add_rn(1.0f, x_nodes[511]);
    const float f0 = fisher_node(fam[0], uf, vf), l0 = fisher_node(fam[0], ul, vl);
    const float f1 = fisher_node(fam[1], uf, vf), l1 = fisher_node(fam[1], ul, vl);
    const float f2 = fisher_node(fam[2], uf, vf), l2 = fisher_node(fam[2], ul, vl);
    const float cY0 = 0.5f * (f0 + l0);
    const float cUY0 = 0.5f * fmaf(uf, f0, ul * l0);
    const float cN1 = 0.5f * ((f1 + l1) - fmaf(uf, f1, ul * l1));
    const float cN2 = 0.5f * ((f2 + l2) - fmaf(uf, f2, ul * l2));
    *F = Y0 - cY0;
    *N0 = *F - (UY0 - cUY0);
    *N1 = n1 - cN1;
    *N2 = n2 - cN2;
}

// run descriptors must classify conservatively: every pair marked uniform really is
extern "C" int emul_check_runs(const float* S, long n) {
    float x_nodes[512];
    for (int i = 0; i < 512; ++i) x_nodes[i] = quad_node((float)i);
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        Family fam[3];
        fisher_families(S + 3 * i, fam);
        for (int f = 0; f < 3; ++f) {
            const unsigned r = family_runs(fam[f]);
            const int b1 = r & 15, m0 = (r >> 4) & 15, m1 = (r >> 8) & 15, b4 = (r >> 12) & 15;
            const bool mid_ss = (r >> 16) & 1;
            if (!(0 <= b1 && b1 <= m0 && m0 <= m1 && m1 <= b4 && b4 <= 8)) { ++bad; continue; }
            for (int p = 0; p < 8; ++p) {
                int want_d, want_s;   // 0 large, 1 small, -1 any
                if (p < b1) { want_d = 0; want_s = 1; }
                else if (p >= b4) { want_d = 1; want_s = 0; }
                else if (p >= m0 && p < m1) { want_d = want_s = mid_ss ? 1 : 0; }
                else continue;
                for (int k = 64 * p; k < 64 * p + 64; ++k) {
                    const float u = add_rn(1.0f, -x_nodes[k]), v = add_rn(1.0f, x_nodes[k]);
                    const int sd = fam[f].fd * u <= kBesselSwitch, ss = fam[f].fs * v <= kBesselSwitch;
                    if (sd != want_d || ss != want_s) ++bad;
                }
            }
        }
    }
    return bad;
}

extern "C" {
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
