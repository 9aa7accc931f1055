// Per-step refinement of the added (cap) hydrogens -- SURVEY section 8f rank 1, second half.
//
// What the reference does every MD step between placing the cap hydrogens and the ViSNet evaluation
// (DistanceFragment.get_fragments, src/Fragmentation/distancefrag.py:56-92): ONE call of
// torch.optim.LBFGS(lr = 0.1, max_iter = 10, tolerance_grad = 0.1, tolerance_change = 0.01, no line search, fresh state)
// on the Amber energy of all dipeptides, moving only the added hydrogens (hydrogen/energies.py:211-242): bond, angle and
// dihedral terms that contain one of them, plus Lennard-Jones and Coulomb between them and the non-excluded atoms of the
// same dipeptide (energies.py:9-60; tables hydrogen/ctable.py:58-240).  CPU spec of this file: oracle/caph_ref.c (fp32,
// the same flat arrays), itself pinned against the reference's own HydrogenOptimizer and torch.optim.LBFGS.
//
// Device form: one CTA, the problem as flat term arrays over the PACKED FRAGMENT position buffer the engine reads (so the
// refinement writes straight into it, no extra copy):
//   energy + gradient : thread per term; every term leaves its energy and the gradient on each of its atoms in a scratch
//                       row; the gradient of hydrogen h is then gathered from the scratch rows listed for h (CSR built on
//                       the host) in a fixed order -- no atomics, bit-reproducible;
//   LBFGS             : two-loop recursion over <= max_iter (y, s) pairs kept in global scratch (3*n_H floats each),
//                       dot products by a fixed-shape block reduction; every thread holds the same scalars, so the
//                       optimiser's data-dependent exits are uniform branches;
//   mirrors           : ACE-NME fragments have no geometry of their own (distancefrag.py:286-307): their added hydrogens
//                       are copies of the neighbouring dipeptides' and are re-copied after the relaxation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vb {

constexpr int CAPH_THREADS = 1024;

struct CaphDev {
    int n_h;            const int* h_idx;                                     // optimised hydrogens (packed fragment atom indices)
    int n_bonds;        const int* bond_ij;  const float* bond_k;  const float* bond_r0;
    int n_angles;       const int* angle_ijk; const float* angle_k; const float* angle_t0;
    int n_dih;          const int* dih_ijkl; const float* dih_k;   const float* dih_n; const float* dih_p;
    int n_pairs;        const int* pair_ij;  const float* pair_a;  const float* pair_b; const float* pair_qq;
    int n_mirror;       const int* mirror_dst; const int* mirror_src;
    const int* gat_rowptr;   // [n_h + 1]
    const int* gat_entry;    // scratch rows (term * 4 + slot) that carry a gradient on hydrogen h
    float scnb, scee;
    int max_iter;
    float lr, tol_grad, tol_change;
    // scratch (global)
    float* tg;          // [n_terms * 4][3] per-term gradient rows
    float* vec;         // [(2 * max_iter + 4)][3 * n_h]: g, prev_g, d, q, Y[max_iter], S[max_iter]
    int* evals_out;     // [1] number of energy evaluations of the last call (diagnostic)
};

namespace caph {

__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// fixed-shape block reductions; the result is returned to every thread
template <typename Op>
__device__ __forceinline__ float block_reduce(float v, float* red, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();                                   // protects `red` against the previous call's readers
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < CAPH_THREADS / 32; w++) r = op(r, red[w]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) { return block_reduce(v, red, [](float a, float b) { return a + b; }); }
__device__ __forceinline__ float block_max(float v, float* red) { return block_reduce(v, red, [](float a, float b) { return fmaxf(a, b); }); }

// total energy (returned to every thread) and the hydrogens' gradient -> g[3 * n_h]
__device__ float energy_grad(const CaphDev& p, const float* __restrict__ x, float* __restrict__ g, float* red) {
    float e_bond = 0.f, e_ang = 0.f, e_dih = 0.f, e_vdw = 0.f, e_el = 0.f;
    float* tg = p.tg;
    int base = 0;
    for (int t = threadIdx.x; t < p.n_bonds; t += CAPH_THREADS) {
        const int i = p.bond_ij[2 * t], j = p.bond_ij[2 * t + 1];
        const float d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        const float r = sqrtf(dot3(d, d));
        const float dr = r - p.bond_r0[t];
        e_bond += p.bond_k[t] * dr * dr;
        const float f = p.bond_k[t] * dr / r;
        float* row = tg + (size_t)(base + t) * 12;
        for (int c = 0; c < 3; c++) { row[c] = f * d[c]; row[3 + c] = -(f * d[c]); }
    }
    base += p.n_bonds;
    for (int t = threadIdx.x; t < p.n_angles; t += CAPH_THREADS) {
        const int i = p.angle_ijk[3 * t], j = p.angle_ijk[3 * t + 1], k = p.angle_ijk[3 * t + 2];
        float a[3], b[3], c[3], bc[3], ca[3];
        for (int q = 0; q < 3; q++) { a[q] = x[3 * i + q] - x[3 * j + q]; b[q] = x[3 * k + q] - x[3 * j + q]; }
        cross3(a, b, c);
        const float yy = sqrtf(dot3(c, c)), xx = dot3(a, b);
        const float th = atan2f(yy, xx);
        const float dth = th - p.angle_t0[t];
        e_ang += p.angle_k[t] * dth * dth;
        const float den = xx * xx + yy * yy, w = p.angle_k[t] * dth;
        cross3(b, c, bc);
        cross3(c, a, ca);
        float* row = tg + (size_t)(base + t) * 12;
        for (int q = 0; q < 3; q++) {
            const float da = (xx * bc[q] / yy - yy * b[q]) / den, db = (xx * ca[q] / yy - yy * a[q]) / den;
            row[q] = w * da;              // atom i
            row[3 + q] = -(w * (da + db));  // atom j
            row[6 + q] = w * db;          // atom k
        }
    }
    base += p.n_angles;
    for (int t = threadIdx.x; t < p.n_dih; t += CAPH_THREADS) {
        const int i = p.dih_ijkl[4 * t], j = p.dih_ijkl[4 * t + 1], k = p.dih_ijkl[4 * t + 2], l = p.dih_ijkl[4 * t + 3];
        float F[3], G[3], H[3], A[3], B[3], w1[3], n1[3], n2[3], m1[3], gu[3];
        for (int q = 0; q < 3; q++) {
            F[q] = x[3 * i + q] - x[3 * j + q];
            G[q] = x[3 * j + q] - x[3 * k + q];
            H[q] = x[3 * l + q] - x[3 * k + q];
            w1[q] = -F[q];
        }
        cross3(w1, G, n1);
        cross3(G, H, n2);
        const float l1 = sqrtf(dot3(n1, n1)), l2 = sqrtf(dot3(n2, n2)), gn = sqrtf(dot3(G, G));
        for (int q = 0; q < 3; q++) { n1[q] /= l1; n2[q] /= l2; gu[q] = G[q] / gn; }
        cross3(n1, gu, m1);
        const float phi = atan2f(dot3(m1, n2), dot3(n1, n2));
        const float arg = p.dih_n[t] * phi - p.dih_p[t];
        e_dih += p.dih_k[t] * (1.0f + cosf(arg));
        const float de = -0.5f * p.dih_k[t] * p.dih_n[t] * sinf(arg);
        cross3(F, G, A);
        cross3(H, G, B);
        const float aa = dot3(A, A), bb = dot3(B, B), fg = dot3(F, G), hg = dot3(H, G);
        float* row = tg + (size_t)(base + t) * 12;
        for (int q = 0; q < 3; q++) {
            const float dp0 = -gn / aa * A[q], dp3 = gn / bb * B[q];
            const float s = fg / (aa * gn) * A[q] - hg / (bb * gn) * B[q];
            row[q] = de * dp0;
            row[3 + q] = de * (-dp0 + s);
            row[6 + q] = de * (-dp3 - s);
            row[9 + q] = de * dp3;
        }
    }
    base += p.n_dih;
    for (int t = threadIdx.x; t < p.n_pairs; t += CAPH_THREADS) {
        const int i = p.pair_ij[2 * t], j = p.pair_ij[2 * t + 1];
        const float d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        const float r2 = dot3(d, d), r = sqrtf(r2), r6 = r2 * r2 * r2;
        e_vdw += p.pair_a[t] / (r6 * r6) - p.pair_b[t] / r6;
        e_el += p.pair_qq[t] / r;
        const float de = (-12.0f * p.pair_a[t] / (r6 * r6 * r) + 6.0f * p.pair_b[t] / (r6 * r)) / p.scnb - p.pair_qq[t] / r2 / p.scee;
        float* row = tg + (size_t)(base + t) * 12;
        for (int c = 0; c < 3; c++) { row[c] = de / r * d[c]; row[3 + c] = -(de / r * d[c]); }
    }
    const float eb = block_sum(e_bond, red), ea = block_sum(e_ang, red), ed = block_sum(e_dih, red);
    const float ev = block_sum(e_vdw, red), ee = block_sum(e_el, red);       // (the barriers inside publish the scratch rows)
    for (int h = threadIdx.x; h < p.n_h; h += CAPH_THREADS) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int m = p.gat_rowptr[h]; m < p.gat_rowptr[h + 1]; m++) {
            const float* row = tg + (size_t)p.gat_entry[m] * 3;
            gx += row[0]; gy += row[1]; gz += row[2];
        }
        g[3 * h] = gx; g[3 * h + 1] = gy; g[3 * h + 2] = gz;
    }
    __syncthreads();
    return 0.5f * eb + 0.5f * ea + 0.5f * ed + ev / p.scnb + ee / p.scee;
}

}  // namespace caph

// One LBFGS call over all optimised hydrogens, in place on the packed fragment position buffer x.
__global__ void __launch_bounds__(CAPH_THREADS) caph_relax_kernel(CaphDev p, float* __restrict__ x) {
    using namespace caph;
    __shared__ float red[CAPH_THREADS / 32];
    __shared__ float ro[64], al[64];
    const int n = 3 * p.n_h;
    const int tid = threadIdx.x;
    float* g = p.vec;
    float* prev_g = g + n;
    float* d = prev_g + n;
    float* q = d + n;
    float* Y = q + n;
    float* S = Y + (size_t)p.max_iter * n;
    int evals = 0;
    if (n > 0 && p.max_iter > 0) {
        float loss = energy_grad(p, x, g, red), prev_loss = loss, H_diag = 1.0f, t = 0.f;
        evals = 1;
        int n_old = 0, n_iter = 0;
        const int max_eval = p.max_iter * 5 / 4;
        float gmax = 0.f;
        for (int u = tid; u < n; u += CAPH_THREADS) gmax = fmaxf(gmax, fabsf(g[u]));
        gmax = block_max(gmax, red);
        if (gmax > p.tol_grad) {
            while (n_iter < p.max_iter) {
                n_iter++;
                if (n_iter == 1) {
                    for (int u = tid; u < n; u += CAPH_THREADS) d[u] = -g[u];
                } else {
                    float* y = Y + (size_t)n_old * n;
                    float* s = S + (size_t)n_old * n;        // slot of the candidate pair
                    float ys = 0.f, yy = 0.f;
                    for (int u = tid; u < n; u += CAPH_THREADS) {
                        const float yv = g[u] - prev_g[u], sv = d[u] * t;
                        y[u] = yv; s[u] = sv;
                        ys += yv * sv; yy += yv * yv;
                    }
                    ys = block_sum(ys, red);
                    yy = block_sum(yy, red);
                    if (ys > 1e-10f) {
                        if (tid == 0) ro[n_old] = 1.0f / ys;
                        H_diag = ys / yy;
                        n_old++;                             // at most max_iter - 1 pairs per call: the history never overflows
                    }
                    for (int u = tid; u < n; u += CAPH_THREADS) q[u] = -g[u];
                    __syncthreads();
                    for (int i = n_old - 1; i >= 0; i--) {
                        float sq = 0.f;
                        for (int u = tid; u < n; u += CAPH_THREADS) sq += S[(size_t)i * n + u] * q[u];
                        sq = block_sum(sq, red);
                        const float a_i = sq * ro[i];
                        if (tid == 0) al[i] = a_i;
                        for (int u = tid; u < n; u += CAPH_THREADS) q[u] -= a_i * Y[(size_t)i * n + u];
                    }
                    for (int u = tid; u < n; u += CAPH_THREADS) d[u] = q[u] * H_diag;
                    __syncthreads();
                    for (int i = 0; i < n_old; i++) {
                        float yr = 0.f;
                        for (int u = tid; u < n; u += CAPH_THREADS) yr += Y[(size_t)i * n + u] * d[u];
                        yr = block_sum(yr, red);
                        const float be = yr * ro[i];
                        const float a_i = al[i];
                        for (int u = tid; u < n; u += CAPH_THREADS) d[u] += (a_i - be) * S[(size_t)i * n + u];
                    }
                }
                float l1 = 0.f, gtd = 0.f;
                for (int u = tid; u < n; u += CAPH_THREADS) {     // each thread only touches its own elements u = tid (mod T)
                    const float gv = g[u];
                    prev_g[u] = gv;
                    l1 += fabsf(gv);
                    gtd += gv * d[u];
                }
                prev_loss = loss;
                if (n_iter == 1) {
                    l1 = block_sum(l1, red);
                    t = fminf(1.0f, 1.0f / l1) * p.lr;
                } else {
                    t = p.lr;
                }
                gtd = block_sum(gtd, red);
                if (gtd > -p.tol_change) break;
                for (int h = tid; h < p.n_h; h += CAPH_THREADS) {
                    const int a = p.h_idx[h];
                    x[3 * a] += t * d[3 * h]; x[3 * a + 1] += t * d[3 * h + 1]; x[3 * a + 2] += t * d[3 * h + 2];
                }
                __syncthreads();
                bool opt_cond = false;
                if (n_iter != p.max_iter) {
                    loss = energy_grad(p, x, g, red);
                    evals++;
                    gmax = 0.f;
                    for (int u = tid; u < n; u += CAPH_THREADS) gmax = fmaxf(gmax, fabsf(g[u]));
                    gmax = block_max(gmax, red);
                    opt_cond = gmax <= p.tol_grad;
                }
                if (n_iter == p.max_iter || evals >= max_eval || opt_cond) break;
                float dmax = 0.f;
                for (int u = tid; u < n; u += CAPH_THREADS) dmax = fmaxf(dmax, fabsf(d[u] * t));
                dmax = block_max(dmax, red);
                if (dmax <= p.tol_change) break;
                if (fabsf(loss - prev_loss) < p.tol_change) break;
            }
        }
    }
    __syncthreads();
    // ACE-NME copies of the relaxed hydrogens
    for (int m = tid; m < p.n_mirror; m += CAPH_THREADS) {
        const int dst = p.mirror_dst[m], src = p.mirror_src[m];
        x[3 * dst] = x[3 * src]; x[3 * dst + 1] = x[3 * src + 1]; x[3 * dst + 2] = x[3 * src + 2];
    }
    if (tid == 0 && p.evals_out) *p.evals_out = evals;
}

}  // namespace vb
