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

// Mirrors the warp decomposition of the kernel: per family the nodes [cut,512) split into up to
// three pair-aligned runs of uniform type plus up to four single edge nodes (family_plan); a run is
// walked in passes of 32 node pairs, lane l taking the pair (2m, 2m+1), pairs past the run's end
// skipped; each lane keeps (lo,hi) partial sums of y and u*y per run, scales them by the run's
// constant factor, and the warp butterfly-reduces {Y0, UY0, N1, N2}.  The edge nodes and the
// trapezoid end-point halves are the owning thread's corrections.
static float butterfly(float* v) {
    for (int off = 16; off >= 1; off >>= 1)
        for (int lane = 0; lane < 32; ++lane)
            if ((lane & off) == 0) v[lane] += v[lane ^ off];
    return v[0];
}

struct Tables {
    NodeVals n[576];
    float uv[1024 + 2];      // pair-interleaved (u0,u1,v0,v1) like the kernel's SSa column; v = u + 2
    const float* u;
    const float* v;
    Tables() {
        for (int i = 0; i < 576; ++i) n[i] = node_vals(i);
        for (int i = 0; i < 512; ++i) { uv[((i >> 1) << 2) + (i & 1)] = n[i].u; uv[((i >> 1) << 2) + 2 + (i & 1)] = n[i].v; }
        u = uv; v = uv + 2;
    }
};
static const Tables& tables() { static Tables t; return t; }

// one run for one lane, walked exactly like quad_run/quad_pass: passes of 64 pairs while more than
// 32 remain, then one of 32; the lane's pair counts while `left` = pairs from its pair to the run's
// end is positive; the run's factor applied at the end
static void run_lane(const FamilyDesc& d, int type, uint32_t word, int lane, float* Y, float* UY) {
    const Tables& tb = tables();
    int pairs = run_word_pairs(word);
    if (pairs == 0) return;
    unsigned m = run_word_m0(word);
    int left = pairs - lane;
    float ylo = 0, yhi = 0, ulo = 0, uhi = 0;
    while (pairs > 0) {
        const int W = pairs > 32 ? 2 : 1;
        for (int w = 0; w < W; ++w) {
            if (!(left > 32 * w)) continue;
            const int i0 = 2 * (int)(m + lane + 32 * w), i1 = i0 + 1;
            const float y0 = node_typed(d, type, tb.n[i0]), y1 = node_typed(d, type, tb.n[i1]);
            ylo += y0; yhi += y1;
            ulo = fmaf(tb.n[i0].u, y0, ulo); uhi = fmaf(tb.n[i1].u, y1, uhi);
        }
        pairs -= 32 * W; m += 32 * W; left -= 32 * W;
    }
    const float sc = type_scale(d, type);
    *Y = fmaf(sc, ylo + yhi, *Y);
    *UY = fmaf(sc, ulo + uhi, *UY);
}

static float end_node(const FamilyDesc& d, int i) {
    const int t = node_type(d, i);
    return node_typed(d, t, tables().n[i]) * type_scale(d, t);
}

static void quadrature(const float* s, int cut_bits, float* F, float* N0, float* N1, float* N2) {
    const Tables& tb = tables();
    FamilyDesc fam[3];
    fisher_families(s, tb.u, tb.v, cut_bits, fam);
    float pY0[32], pUY0[32], pN1[32], pN2[32];
    for (int lane = 0; lane < 32; ++lane) {
        for (int f = 0; f < 3; ++f) {
            const FamilyDesc& d = fam[f];
            FamilyPlan pl;
            family_plan(d, pl);
            float Y = 0, UY = 0;
            run_lane(d, kLS, pl.word[0], lane, &Y, &UY);
            run_lane(d, run_word_mid_ll(pl.word[1]) ? kLL : kSS, pl.word[1], lane, &Y, &UY);
            run_lane(d, kSL, pl.word[2], lane, &Y, &UY);
            if (f == 0) { pY0[lane] = Y; pUY0[lane] = UY; }
            else if (f == 1) pN1[lane] = Y - UY;
            else pN2[lane] = Y - UY;
        }
    }
    const float Y0 = butterfly(pY0), UY0 = butterfly(pUY0), n1 = butterfly(pN1), n2 = butterfly(pN2);
    const float uf = tb.n[0].u, ul = tb.n[511].u;
    const float f0 = fam[0].cut ? 0.f : end_node(fam[0], 0), l0 = end_node(fam[0], 511);
    const float f1 = fam[1].cut ? 0.f : end_node(fam[1], 0), l1 = end_node(fam[1], 511);
    const float f2 = fam[2].cut ? 0.f : end_node(fam[2], 0), l2 = end_node(fam[2], 511);
    // edge nodes of the owning thread (same order as the kernel: family, then edge slot)
    float eY[3], eUY[3];
    for (int f = 0; f < 3; ++f) {
        FamilyPlan pl;
        family_plan(fam[f], pl);
        eY[f] = eUY[f] = 0.f;
        for (int k = 0; k < 4; ++k) {
            if (pl.edge[k] < 0) continue;
            const float y = end_node(fam[f], pl.edge[k]);
            eY[f] += y;
            eUY[f] = fmaf(tb.n[pl.edge[k]].u, y, eUY[f]);
        }
    }
    const float cY0 = 0.5f * (f0 + l0) - eY[0];
    const float cUY0 = 0.5f * fmaf(uf, f0, ul * l0) - eUY[0];
    const float cN1 = (0.5f * (f1 + l1) - eY[1]) - (0.5f * fmaf(uf, f1, ul * l1) - eUY[1]);
    const float cN2 = (0.5f * (f2 + l2) - eY[2]) - (0.5f * fmaf(uf, f2, ul * l2) - eUY[2]);
    *F = Y0 - cY0;
    *N0 = *F - (UY0 - cUY0);
    *N1 = n1 - cN1;
    *N2 = n2 - cN2;
}

// every node of [cut,512) must be covered exactly once -- by a pair of a run whose type is the one the
// reference's branch choice dictates for BOTH nodes of the pair, or by an edge node of the right
// type -- and no pass may reach past the padded tables
extern "C" int emul_check_items(const float* S, long n, int cut_bits) {
    const Tables& tb = tables();
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        FamilyDesc fam[3];
        fisher_families(S + 3 * i, tb.u, tb.v, cut_bits, fam);
        for (int f = 0; f < 3; ++f) {
            FamilyPlan pl;
            family_plan(fam[f], pl);
            if (fam[f].cut & 1) ++bad;
            int seen[512] = {};
            for (int r = 0; r < 3; ++r) {
                const int type = r == 0 ? kLS : (r == 2 ? kSL : (run_word_mid_ll(pl.word[r]) ? kLL : kSS));
                if (r == 1 && (run_word_mid_ll(pl.word[r]) ? kLL : kSS) != fam[f].mid) ++bad;
                const unsigned m0 = run_word_m0(pl.word[r]);
                const int pairs = run_word_pairs(pl.word[r]);
                if (pairs == 0) continue;
                unsigned m = m0;
                int done = 0;
                while (pairs - done > 0) {
                    const int W = (pairs - done) > 32 ? 2 : 1;
                    if (m + 32u * W > 288u) ++bad;
                    m += 32 * W; done += 32 * W;
                }
                for (int t = 0; t < 2 * pairs; ++t) {
                    const int node = 2 * (int)m0 + t;
                    if (node >= 512 || node < fam[f].cut) { ++bad; continue; }
                    ++seen[node];
                    if (node_type(fam[f], node) != type) ++bad;
                }
            }
            for (int k = 0; k < 4; ++k) {
                const int node = pl.edge[k];
                if (node < 0) continue;
                if (node >= 512 || node < fam[f].cut) { ++bad; continue; }
                ++seen[node];
                if (node_type(fam[f], node) != edge_type(fam[f], k)) ++bad;
            }
            for (int node = 0; node < 512; ++node)
                if (seen[node] != (node >= fam[f].cut ? 1 : 0)) ++bad;
        }
    }
    return bad;
}

// per sample: the three families' cut indices, the valid nodes and the node slots the kernel's passes
// execute (a run is rounded up to whole 64-slot passes) -- for the float64 check of the cut bound and
// for the executed-work accounting in DESIGN.md
extern "C" void emul_cut_info(const float* S, long n, int cut_bits, int* cut, int* valid, int* slots_exec) {
    const Tables& tb = tables();
    for (long i = 0; i < n; ++i) {
        FamilyDesc fam[3];
        fisher_families(S + 3 * i, tb.u, tb.v, cut_bits, fam);
        int nv = 0, ns = 0;
        for (int f = 0; f < 3; ++f) {
            cut[3 * i + f] = fam[f].cut;
            nv += 512 - fam[f].cut;
            FamilyPlan pl;
            family_plan(fam[f], pl);
            for (int r = 0; r < 3; ++r) ns += (run_word_pairs(pl.word[r]) + 31) / 32 * 64;
        }
        valid[i] = nv; slots_exec[i] = ns;
    }
}

// the run boundaries must reproduce the reference's per-node branch choice exactly:
// node i is d-small iff fl(fd*u_i) <= 3.75 and s-small iff fl(fs*v_i) <= 3.75
extern "C" int emul_check_runs(const float* S, long n) {
    const Tables& tb = tables();
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        FamilyDesc fam[3];
        fisher_families(S + 3 * i, tb.u, tb.v, 0, fam);
        for (int f = 0; f < 3; ++f) {
            const FamilyDesc& d = fam[f];
            if (!(0 <= d.b0 && d.b0 <= d.b1 && d.b1 <= 512)) { ++bad; continue; }
            for (int k = 0; k < 512; ++k) {
                const bool sd = d.fd * tb.n[k].u <= kBesselSwitch, ss = d.fs * tb.n[k].v <= kBesselSwitch;
                const int want = sd ? (ss ? kSS : kSL) : (ss ? kLS : kLL);
                if (node_type(d, k) != want) ++bad;
            }
        }
    }
    return bad;
}

extern "C" {
void emul_fisher(const float* A, const float* Rgt, long n, float overreg, int cut_bits,
                 float* nll, float* grad, float* Rout, float* entropy, float* logC, float* S, float* G) {
    for (long i = 0; i < n; ++i) {
        const float* a = A + 9 * i;
        float u[9], v[9], s[3];
        proper_svd3(a, u, v, s);
        float F, N0, N1, N2;
        quadrature(s, cut_bits, &F, &N0, &N1, &N2);
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

// fisher_CE(A1 target, A2 prediction): value and gradient w.r.t. A2, with the kernels' per-sample math
void emul_fisher_ce(const float* A1, const float* A2, long n, int cut_bits, float* ce, float* grad) {
    for (long i = 0; i < n; ++i) {
        float u1[9], v1[9], s1[3], u2[9], v2[9], s2[3], F, N0, N1, N2;
        proper_svd3(A1 + 9 * i, u1, v1, s1);
        quadrature(s1, cut_bits, &F, &N0, &N1, &N2);
        const FisherStats t = fisher_finish(s1, F, N0, N1, N2);
        proper_svd3(A2 + 9 * i, u2, v2, s2);
        quadrature(s2, cut_bits, &F, &N0, &N1, &N2);
        const FisherStats p = fisher_finish(s2, F, N0, N1, N2);
        ce[i] = fisher_ce_close(u1, v1, t.g, u2, v2, s2, p.g, p.logf, grad + 9 * i);
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

void emul_euler_dad(const float* R, long n, float* out) {
    for (long i = 0; i < n; ++i) euler_dad_degrees(R + 9 * i, out + 3 * i);
}

void emul_atan2(const float* y, const float* x, long n, float* out) {
    for (long i = 0; i < n; ++i) out[i] = atan2_so3(y[i], x[i]);
}

void emul_rad_to_deg(const float* rad, long n, float* out) {
    for (long i = 0; i < n; ++i) out[i] = rad_to_deg_ref(rad[i]);
}

// K2L per-sample set-up as the kernels run it (fp32 Jacobi + fp64 polishing) next to seven fp64 sweeps from scratch
void emul_laplace_setup(const float* A, long n, float* Rs, double* T, double* Rs_f64, double* T_f64) {
    for (long i = 0; i < n; ++i) {
        laplace_setup(A + 9 * i, Rs + 9 * i, T + i);
        double U[9], V[9], s[3];
        proper_svd3_t<double>(A + 9 * i, U, V, s);
        T_f64[i] = s[0] + s[1] + s[2];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                Rs_f64[9 * i + 3 * r + c] = U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1] + U[3 * r + 2] * V[3 * c + 2];
    }
}

// K2L with the kernel's two decompositions: L = 1 (one thread walks the grid in order) and
// L = 32 (lane l takes points l, l+32, ...; partials rebased to the common minimum, butterfly-merged)
void emul_laplace(const float* A, const float* Rgt, long n, const float* grid, int N, int L,
                  float* nll, float* grad, float* mode, float* logF) {
    for (long i = 0; i < n; ++i) {
        const float* a9 = A + 9 * i;
        float Rs[9];
        double Td;
        laplace_setup(a9, Rs, &Td);
        const float T = (float)Td;
        LaplaceAccum acc[32];
        for (int l = 0; l < L; ++l) {
            laplace_accum_init(acc[l]);
            if (L == 1) {
                for (int k0 = 0; k0 < N; k0 += 128) {
                    for (int k = k0; k < N && k < k0 + 128; ++k) laplace_accum_point(acc[l], a9, T, grid + 9 * k);
                    laplace_accum_flush(acc[l]);
                }
            } else {
                for (int k = l; k < N; k += L) laplace_accum_point(acc[l], a9, T, grid + 9 * k);
                laplace_accum_flush(acc[l]);
            }
        }
        if (L > 1) {
            float qg = acc[0].qmin;
            for (int l = 1; l < L; ++l) qg = fminf(qg, acc[l].qmin);
            for (int l = 0; l < L; ++l) laplace_accum_rebase(acc[l], qg);
            for (int off = 16; off >= 1; off >>= 1)
                for (int l = 0; l < 32; ++l)
                    if ((l & off) == 0) {
                        acc[l].Z += acc[l ^ off].Z; acc[l].C += acc[l ^ off].C;
                        for (int k = 0; k < 9; ++k) acc[l].M[k] += acc[l ^ off].M[k];
                    }
        }
        laplace_finish(acc[0], laplace_gt_gap(a9, Rgt + 9 * i, Td), Rs, Rgt + 9 * i, N, nll + i, logF + i, grad + 9 * i);
        for (int k = 0; k < 9; ++k) mode[9 * i + k] = Rs[k];
    }
}

}  // extern "C"
