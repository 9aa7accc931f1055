/* C restatement (fp32) of the reference's per-step cap-hydrogen refinement -- TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithm as oracle/caph_ref.py (which is pinned against the reference's own HydrogenOptimizer and torch.optim.LBFGS),
 * written over flat arrays the way a device kernel will see the problem: all dipeptides concatenated into one position
 * buffer, every term carries global atom indices and its own parameters, gradients are needed only on the added hydrogens.
 *   energy terms   /root/reference/src/Fragmentation/hydrogen/energies.py:9-60
 *   optimiser      energies.py:211-242 -> torch.optim.LBFGS(lr 0.1, max_iter, tolerance_grad 0.1, tolerance_change 0.01),
 *                  no line search, fresh state every call (torch/optim/lbfgs.py, restated)
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off (oracle/caph_c.py).  Checked against the Python oracle and the golden
 * vectors in tests/test_caph.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n_atoms;          /* atoms in the concatenated buffer                                         */
    int64_t n_h;              /* added hydrogens (the unknowns)                                            */
    const int64_t* h_idx;     /* [n_h] their atom indices                                                  */
    int64_t n_bonds;  const int64_t* bond_ij;   const float* bond_k;  const float* bond_r0;      /* [n][2] */
    int64_t n_angles; const int64_t* angle_ijk; const float* angle_k; const float* angle_t0;     /* [n][3] */
    int64_t n_dih;    const int64_t* dih_ijkl;  const float* dih_k;   const float* dih_n; const float* dih_p;
    int64_t n_pairs;  const int64_t* pair_ij;   const float* pair_a;  const float* pair_b; const float* pair_qq;
    float scnb, scee;
} caph_problem;

static void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* total energy and dE/dx for every atom (caller reads the hydrogens' rows); g must hold 3*n_atoms floats */
static float energy_grad(const caph_problem* p, const float* x, float* g) {
    memset(g, 0, sizeof(float) * 3 * (size_t)p->n_atoms);
    float e_bond = 0.f, e_ang = 0.f, e_dih = 0.f, e_vdw = 0.f, e_el = 0.f;
    for (int64_t t = 0; t < p->n_bonds; t++) {
        const int64_t i = p->bond_ij[2 * t], j = p->bond_ij[2 * t + 1];
        float d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        const float r = sqrtf(dot3(d, d));
        const float dr = r - p->bond_r0[t];
        e_bond += p->bond_k[t] * dr * dr;
        const float f = p->bond_k[t] * dr / r;
        for (int c = 0; c < 3; c++) { g[3 * i + c] += f * d[c]; g[3 * j + c] -= f * d[c]; }
    }
    for (int64_t t = 0; t < p->n_angles; t++) {
        const int64_t i = p->angle_ijk[3 * t], j = p->angle_ijk[3 * t + 1], k = p->angle_ijk[3 * t + 2];
        float a[3], b[3], c[3], bc[3], ca[3];
        for (int q = 0; q < 3; q++) { a[q] = x[3 * i + q] - x[3 * j + q]; b[q] = x[3 * k + q] - x[3 * j + q]; }
        cross3(a, b, c);
        const float yy = sqrtf(dot3(c, c)), xx = dot3(a, b);
        const float th = atan2f(yy, xx);
        const float dth = th - p->angle_t0[t];
        e_ang += p->angle_k[t] * dth * dth;
        const float den = xx * xx + yy * yy, w = p->angle_k[t] * dth;
        cross3(b, c, bc);
        cross3(c, a, ca);
        for (int q = 0; q < 3; q++) {
            const float da = (xx * bc[q] / yy - yy * b[q]) / den, db = (xx * ca[q] / yy - yy * a[q]) / den;
            g[3 * i + q] += w * da;
            g[3 * k + q] += w * db;
            g[3 * j + q] -= w * (da + db);
        }
    }
    for (int64_t t = 0; t < p->n_dih; t++) {
        const int64_t i = p->dih_ijkl[4 * t], j = p->dih_ijkl[4 * t + 1], k = p->dih_ijkl[4 * t + 2], l = p->dih_ijkl[4 * t + 3];
        float F[3], G[3], H[3], A[3], B[3], w1[3], n1[3], n2[3], m1[3], gu[3];
        for (int q = 0; q < 3; q++) {
            F[q] = x[3 * i + q] - x[3 * j + q];
            G[q] = x[3 * j + q] - x[3 * k + q];
            H[q] = x[3 * l + q] - x[3 * k + q];
            w1[q] = -F[q];
        }
        /* phi exactly as energies.py:33-41: v0 = G, v1 = p1 - p0, v2 = H */
        cross3(w1, G, n1);
        cross3(G, H, n2);
        const float l1 = sqrtf(dot3(n1, n1)), l2 = sqrtf(dot3(n2, n2)), gn = sqrtf(dot3(G, G));
        for (int q = 0; q < 3; q++) { n1[q] /= l1; n2[q] /= l2; gu[q] = G[q] / gn; }
        cross3(n1, gu, m1);
        const float phi = atan2f(dot3(m1, n2), dot3(n1, n2));
        const float arg = p->dih_n[t] * phi - p->dih_p[t];
        e_dih += p->dih_k[t] * (1.0f + cosf(arg));
        const float de = -0.5f * p->dih_k[t] * p->dih_n[t] * sinf(arg);
        cross3(F, G, A);
        cross3(H, G, B);
        const float aa = dot3(A, A), bb = dot3(B, B), fg = dot3(F, G), hg = dot3(H, G);
        for (int q = 0; q < 3; q++) {
            const float dp0 = -gn / aa * A[q], dp3 = gn / bb * B[q];
            const float s = fg / (aa * gn) * A[q] - hg / (bb * gn) * B[q];
            g[3 * i + q] += de * dp0;
            g[3 * j + q] += de * (-dp0 + s);
            g[3 * k + q] += de * (-dp3 - s);
            g[3 * l + q] += de * dp3;
        }
    }
    for (int64_t t = 0; t < p->n_pairs; t++) {
        const int64_t i = p->pair_ij[2 * t], j = p->pair_ij[2 * t + 1];
        float d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        const float r2 = dot3(d, d), r = sqrtf(r2), r6 = r2 * r2 * r2;
        e_vdw += p->pair_a[t] / (r6 * r6) - p->pair_b[t] / r6;
        e_el += p->pair_qq[t] / r;
        const float de = (-12.0f * p->pair_a[t] / (r6 * r6 * r) + 6.0f * p->pair_b[t] / (r6 * r)) / p->scnb - p->pair_qq[t] / r2 / p->scee;
        for (int c = 0; c < 3; c++) { g[3 * i + c] += de / r * d[c]; g[3 * j + c] -= de / r * d[c]; }
    }
    return 0.5f * e_bond + 0.5f * e_ang + 0.5f * e_dih + e_vdw / p->scnb + e_el / p->scee;
}

float caph_energy_grad(const caph_problem* p, const float* x, float* g_all) { return energy_grad(p, x, g_all); }

/* One LBFGS call over the hydrogens of the whole buffer, in place.  Returns the number of energy evaluations. */
int caph_relax(const caph_problem* p, float* x, int max_iter, float lr, float tol_grad, float tol_change) {
    const int64_t n = 3 * p->n_h;
    if (n == 0) return 0;
    const int hist = max_iter > 0 ? max_iter : 1;
    float* gall = (float*)malloc(sizeof(float) * 3 * (size_t)p->n_atoms);
    float* g = (float*)malloc(sizeof(float) * n), *prev_g = (float*)malloc(sizeof(float) * n);
    float* d = (float*)malloc(sizeof(float) * n), *q = (float*)malloc(sizeof(float) * n);
    float* Y = (float*)malloc(sizeof(float) * n * hist), *S = (float*)malloc(sizeof(float) * n * hist);
    float* ro = (float*)malloc(sizeof(float) * hist), *al = (float*)malloc(sizeof(float) * hist);
#define GATHER() for (int64_t h = 0; h < p->n_h; h++) for (int c = 0; c < 3; c++) g[3 * h + c] = gall[3 * p->h_idx[h] + c]
    float loss = energy_grad(p, x, gall), prev_loss = loss, H_diag = 1.0f, t = 0.f;
    GATHER();
    int evals = 1, n_old = 0, n_iter = 0;
    const int max_eval = max_iter * 5 / 4;
    float gmax = 0.f;
    for (int64_t u = 0; u < n; u++) gmax = fmaxf(gmax, fabsf(g[u]));
    if (gmax > tol_grad) {
        while (n_iter < max_iter) {
            n_iter++;
            if (n_iter == 1) {
                for (int64_t u = 0; u < n; u++) d[u] = -g[u];
            } else {
                float ys = 0.f, yy = 0.f;
                float* y = Y + (size_t)n_old * n, *s = S + (size_t)n_old * n;     /* slot of the candidate pair */
                for (int64_t u = 0; u < n; u++) { y[u] = g[u] - prev_g[u]; s[u] = d[u] * t; }
                for (int64_t u = 0; u < n; u++) ys += y[u] * s[u];
                if (ys > 1e-10f) {
                    for (int64_t u = 0; u < n; u++) yy += y[u] * y[u];
                    ro[n_old] = 1.0f / ys;
                    H_diag = ys / yy;
                    n_old++;                  /* history never overflows: at most max_iter - 1 pairs per call */
                }
                for (int64_t u = 0; u < n; u++) q[u] = -g[u];
                for (int i = n_old - 1; i >= 0; i--) {
                    float sq = 0.f;
                    for (int64_t u = 0; u < n; u++) sq += S[(size_t)i * n + u] * q[u];
                    al[i] = sq * ro[i];
                    for (int64_t u = 0; u < n; u++) q[u] -= al[i] * Y[(size_t)i * n + u];
                }
                for (int64_t u = 0; u < n; u++) d[u] = q[u] * H_diag;
                for (int i = 0; i < n_old; i++) {
                    float yr = 0.f;
                    for (int64_t u = 0; u < n; u++) yr += Y[(size_t)i * n + u] * d[u];
                    const float be = yr * ro[i];
                    for (int64_t u = 0; u < n; u++) d[u] += (al[i] - be) * S[(size_t)i * n + u];
                }
            }
            memcpy(prev_g, g, sizeof(float) * n);
            prev_loss = loss;
            if (n_iter == 1) {
                float l1 = 0.f;
                for (int64_t u = 0; u < n; u++) l1 += fabsf(g[u]);
                t = fminf(1.0f, 1.0f / l1) * lr;
            } else {
                t = lr;
            }
            float gtd = 0.f;
            for (int64_t u = 0; u < n; u++) gtd += g[u] * d[u];
            if (gtd > -tol_change) break;
            for (int64_t h = 0; h < p->n_h; h++)
                for (int c = 0; c < 3; c++) x[3 * p->h_idx[h] + c] += t * d[3 * h + c];
            int opt_cond = 0;
            if (n_iter != max_iter) {
                loss = energy_grad(p, x, gall);
                GATHER();
                evals++;
                gmax = 0.f;
                for (int64_t u = 0; u < n; u++) gmax = fmaxf(gmax, fabsf(g[u]));
                opt_cond = gmax <= tol_grad;
            }
            if (n_iter == max_iter || evals >= max_eval || opt_cond) break;
            float dmax = 0.f;
            for (int64_t u = 0; u < n; u++) dmax = fmaxf(dmax, fabsf(d[u] * t));
            if (dmax <= tol_change) break;
            if (fabsf(loss - prev_loss) < tol_change) break;
        }
    }
#undef GATHER
    free(gall); free(g); free(prev_g); free(d); free(q); free(Y); free(S); free(ro); free(al);
    return evals;
}
