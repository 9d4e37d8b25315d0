// pg_eval.cuh -- assembling and reducing the (c0+2)x(c0+2) "Pab" matrices of one likelihood evaluation.
//
// The reference rebuilds the full Gram matrices [W0,x,y]^T H^-k [W0,x,y] from scratch for every SNP
// and every lambda (pygemma_model.pyx:938,:943,:1002).  Only the x row/column depends on the SNP; the
// [W0,y] block is a function of lambda alone.  Here that block comes from tables:
//   * exact tables at the 11 fixed lambdas 10^-5..10^5 (every grid-mode evaluation, the bracket scan),
//   * Chebyshev interpolation in log10(lambda) for the SNP-specific lambdas Brent/Newton visit
//     (8 sub-intervals per decade, 12 nodes each: interpolation error < 1e-14 of the Cauchy-Schwarz
//     scale of each entry, i.e. below the rounding error of summing n terms directly).
// The x row is produced by the caller (a warp-collective pass over the rotated genotype vector).
//
// The same source runs warp-parallel on the device (32 lanes, __syncwarp between pivots) and serially on
// the host (tests/host_shim.cpp) so indexing and arithmetic are unit-tested without a GPU.
#pragma once

#include "pg_math.cuh"

namespace pg {

#if defined(__CUDA_ARCH__)
#define PG_LANE ((int)(threadIdx.x & 31))
#define PG_NLANES 32
#define PG_SYNCWARP() __syncwarp()
#else
#define PG_LANE 0
#define PG_NLANES 1
#define PG_SYNCWARP() ((void)0)
#endif

constexpr int kMaxCols = 64;  // c0 + 2 <= 64
constexpr int kMaxTri = kMaxCols * (kMaxCols + 1) / 2;

PG_HD int tri(int r, int s) { return r * (r + 1) / 2 + s; }  // r >= s

struct TriAB {
    unsigned char a, b;
};

// SNP-independent tables for one (W0, y, d): see file header.
struct Tables {
    int c0, k0, T0, NF;     // k0 = c0+1 columns [W0, y]; T0 = k0(k0+1)/2; NF = 3*T0 + 3
    const double* fixtab;   // [kNumFixed][NF]
    const double* itab;     // [kNumIntervals][kNodes][NF]
    const double* basis;    // [kNodes][kNodes]: basis[k*kNodes + j] = eps_j cos(j*pi*(k+1/2)/N) / N
    const TriAB* tri_ab;    // [kMaxTri]: idx -> (a, b) with idx = a(a+1)/2 + b
};

// function index layout inside a table row
PG_HD int fn_pair(const Tables& t, int power /*0,1,2*/, int q) { return power * t.T0 + q; }
PG_HD int fn_trP(const Tables& t) { return 3 * t.T0; }
PG_HD int fn_trPP(const Tables& t) { return 3 * t.T0 + 1; }
PG_HD int fn_logdetH(const Tables& t) { return 3 * t.T0 + 2; }

struct Level0 {
    double trP, trPP, logdetH;
};

// Fill the [W0,y] block of the packed lower triangles A, B (and C when full) for `lam`.
//   fixed_t >= 0 : exact row of the fixed table; otherwise Chebyshev interpolation.
// Full index space: 0..c0-1 = W0 columns, c0 = x, c0+1 = y.
template <bool FULL>
PG_HD_NOINLINE void assemble_w0y(const Tables& t, double lam, int fixed_t, double* A, double* B, double* C,
                                 Level0* l0)
{
    const int c0 = t.c0, NF = t.NF, T0 = t.T0;
    const int npow = FULL ? 3 : 2;
    if (fixed_t >= 0) {
        const double* row = t.fixtab + (size_t)fixed_t * NF;
        for (int q = PG_LANE; q < T0; q += PG_NLANES) {
            const TriAB ab = t.tri_ab[q];
            const int r = (ab.a == c0) ? c0 + 1 : ab.a, s = (ab.b == c0) ? c0 + 1 : ab.b;
            const int dst = tri(r, s);
            A[dst] = row[q];
            B[dst] = row[T0 + q];
            if (FULL) C[dst] = row[2 * T0 + q];
        }
        l0->trP = row[3 * T0]; l0->trPP = row[3 * T0 + 1]; l0->logdetH = row[3 * T0 + 2];
        (void)npow;
        return;
    }
    int iv;
    double xloc;
    interval_of(lam, &iv, &xloc);
    double L[kNodes];
    {
        double T[kNodes];
        T[0] = 1.0; T[1] = xloc;
        for (int j = 2; j < kNodes; ++j) T[j] = 2.0 * xloc * T[j - 1] - T[j - 2];
        for (int k = 0; k < kNodes; ++k) {
            double s = 0.0;
            for (int j = 0; j < kNodes; ++j) s += t.basis[k * kNodes + j] * T[j];
            L[k] = s;
        }
    }
    const double* base = t.itab + (size_t)iv * kNodes * NF;
    for (int q = PG_LANE; q < T0; q += PG_NLANES) {
        const TriAB ab = t.tri_ab[q];
        const int r = (ab.a == c0) ? c0 + 1 : ab.a, s = (ab.b == c0) ? c0 + 1 : ab.b;
        const int dst = tri(r, s);
        double va = 0.0, vb = 0.0, vc = 0.0;
        for (int k = 0; k < kNodes; ++k) {
            const double* row = base + (size_t)k * NF;
            va += L[k] * row[q];
            vb += L[k] * row[T0 + q];
            if (FULL) vc += L[k] * row[2 * T0 + q];
        }
        A[dst] = va;
        B[dst] = vb;
        if (FULL) C[dst] = vc;
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < kNodes; ++k) {
        const double* row = base + (size_t)k * NF + 3 * T0;
        s0 += L[k] * row[0];
        s1 += L[k] * row[1];
        s2 += L[k] * row[2];
    }
    l0->trP = s0; l0->trPP = s1; l0->logdetH = s2;
}

// The Pab recursion over covariates (pygemma_model.pyx:936-963 / :989-1036) on packed lower triangles.
// On entry A, B, (C) hold the level-0 Gram matrices; on exit they hold level c_f = c0+1.
// Order inside a level follows the reference: tr_Pi_Pi, C, tr_Pi, B, logdet, A (all read level i-1 values
// of the pivot column, which no update of level i touches).
template <bool FULL>
PG_HD_NOINLINE void pab_recursion(const Tables& t, double* A, double* B, double* C, const Level0& l0,
                                  bool need_logdet, EvalOut* out)
{
    const int c0 = t.c0, k = c0 + 2, cf = c0 + 1;
    if (PG_LANE == 0) A[0] = cy_max(A[0], kMinVal);  // pyx:939 / :993
    PG_SYNCWARP();
    double trP = l0.trP, trPP = l0.trPP, logdetWHW = 0.0;
    for (int i = 1; i <= cf; ++i) {
        const int p = i - 1;
        const int pp = tri(p, p);
        const double app = A[pp], bpp = B[pp];
        const double cpp = FULL ? C[pp] : 0.0;
        if (i == cf) {  // level c0 complete: Wald scalars (pyx:1529-1533)
            out->xPx = app;
            out->yPx = A[tri(c0 + 1, c0)];
        }
        const double inv = 1.0 / app;
        const double al2 = -inv;                  // -1/a_pp
        const double al4 = bpp / (app * app);     // b_pp/a_pp^2
        double alc = 0.0;
        if (FULL) {
            trPP = trPP + (bpp / app) * (bpp / app) - 2 * (cpp / app);          // pyx:1008-1009
            alc = (cpp / (app * app)) - ((bpp * bpp) / (app * app * app));      // pyx:1011
        }
        trP = trP - bpp / app;                                                  // pyx:948 / :1020
        if (need_logdet) logdetWHW += log(app);                                 // pyx:957 / :1029
        const int q = k - i, T = q * (q + 1) / 2;
        for (int idx = PG_LANE; idx < T; idx += PG_NLANES) {
            const TriAB ab = t.tri_ab[idx];
            const int r = i + ab.a, s = i + ab.b;
            const int rs = tri(r, s), rp = tri(r, p), sp = tri(s, p);
            const double ar = A[rp], as = A[sp], br = B[rp], bs = B[sp];
            if (FULL) {
                const double cr = C[rp], cs = C[sp];
                double v = (C[rs] + alc * ar * as) + al2 * (ar * cs + cr * as) + al2 * (br * bs)
                           + al4 * (ar * bs + br * as);                         // pyx:1011-1014
                if (idx == 0) v = cy_max(v, kMinVal);                           // pyx:1016
                C[rs] = v;
            }
            {
                double v = (B[rs] + al4 * ar * as) + al2 * (ar * bs + br * as);  // pyx:950-951 / :1022-1023
                if (idx == 0) v = cy_max(v, kMinVal);                           // pyx:953 / :1025
                B[rs] = v;
            }
            {
                double v = A[rs] + al2 * ar * as;                               // pyx:959 / :1031
                if (idx == 0) v = cy_max(v, kMinVal);                           // pyx:961 / :1034
                A[rs] = v;
            }
        }
        PG_SYNCWARP();
    }
    const int yy = tri(k - 1, k - 1);
    out->yPy = A[yy];
    out->yPPy = B[yy];
    out->yPPPy = FULL ? C[yy] : NAN;
    out->trP = trP;
    out->trPP = FULL ? trPP : NAN;
    out->logdetH = l0.logdetH;
    out->logdetWHW = logdetWHW;
    out->trH = l0.trP;
    out->trHH = l0.trPP;
    PG_SYNCWARP();
}

// ------------------------------------------------------------------------------------------------
// x-row form of the recursion.  Levels 1..c0 pivot on covariate columns; restricted to the [W0, y] block they
// do not involve the SNP at all, so that part is run ONCE per table lambda (eliminate_w0y_row) and stored as
//   per level p:  al2 = -1/a_pp, al4 = b_pp/a_pp^2, alc = c_pp/a_pp^2 - b_pp^2/a_pp^3     (pyx:1011-1031)
//                 the pivot column of level p: A/B/C[s][p] for s = p+1..c0-1 and s = y
//   finals:       A/B/C[y][y] after c0 levels, tr_Pi, tr_Pi_Pi, logdet_H, partial logdet_Wt_H_inv_W, and the level-0
//                 traces sum h, sum h^2 (h = 1/(lambda d + 1); the ML derivatives of the LRT outputs use them)
// ("table-2 row").  Per SNP only the x row (x.w_j, x.y, x.x at the three powers) is carried through the c0
// levels -- O(c0^2) fused multiply-adds without a division -- followed by the one SNP-dependent pivot (x itself).
// Arithmetic per entry is the expression of pab_recursion above, so both forms agree to rounding.
// Table-2 rows at SNP-specific lambdas are Chebyshev-interpolated like the level-0 tables: Schur complements of
// W^T H^-1 W are analytic for Re(lambda) > 0 (the Hermitian part stays positive definite), i.e. in the strip
// |Im log(lambda)| < pi/2, which still gives rho ~ 22 per 1/8-decade interval and 22^-12 ~ 1e-16.
// ------------------------------------------------------------------------------------------------
struct Tables2 {
    int c0, Tp, NF2;       // Tp = c0(c0+1)/2 pivot-column entries per power; NF2 = 3 c0 + 3 Tp + 9
    const double* fix2;    // [kNumFixed][NF2]
    const double* itab2;   // [kNumIntervals][kNodes][NF2]
    const double* basis;   // as Tables::basis
};

PG_HD int t2_pairs(int c0) { return c0 * (c0 + 1) / 2; }
PG_HD int t2_nf(int c0) { return 3 * c0 + 3 * t2_pairs(c0) + 9; }
// column entry (s, p), p < s <= c0 (s == c0 is y), power 0: index; powers 1, 2 follow at +Tp, +2Tp
PG_HD int t2_col(int c0, int p, int s) { return 3 * c0 + p * c0 - p * (p - 1) / 2 + (s - p - 1); }
PG_HD int t2_fin(int c0) { return 3 * c0 + 3 * t2_pairs(c0); }

// row0: level-0 table row (3 T0 + 3 doubles, [W0, y] packed with y at index c0); work: 3 T0 doubles; out: NF2
PG_HD_NOINLINE void eliminate_w0y_row(int c0, const double* row0, double* work, double* out)
{
    const int k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2, Tp = t2_pairs(c0);
    double* A = work;
    double* B = work + T0;
    double* C = work + 2 * T0;
    for (int q = 0; q < 3 * T0; ++q) work[q] = row0[q];
    double trP = row0[3 * T0], trPP = row0[3 * T0 + 1], logdet = 0.0;
    if (c0 >= 1) A[0] = cy_max(A[0], kMinVal);  // pyx:939 / :993
    for (int i = 1; i <= c0; ++i) {
        const int p = i - 1, pp = tri(p, p);
        const double app = A[pp], bpp = B[pp], cpp = C[pp];
        const double inv = 1.0 / app;
        const double al2 = -inv, al4 = bpp / (app * app);
        trPP = trPP + (bpp / app) * (bpp / app) - 2 * (cpp / app);
        const double alc = (cpp / (app * app)) - ((bpp * bpp) / (app * app * app));
        trP = trP - bpp / app;
        logdet += log(app);
        out[3 * p] = al2; out[3 * p + 1] = al4; out[3 * p + 2] = alc;
        for (int s = i; s <= c0; ++s) {
            const int q = t2_col(c0, p, s), sp = tri(s, p);
            out[q] = A[sp]; out[Tp + q] = B[sp]; out[2 * Tp + q] = C[sp];
        }
        for (int r = i; r <= c0; ++r)
            for (int s2 = i; s2 <= r; ++s2) {
                const int rs = tri(r, s2), rp = tri(r, p), sp = tri(s2, p);
                const double ar = A[rp], as = A[sp], br = B[rp], bs = B[sp], cr = C[rp], cs = C[sp];
                const bool clamp = (r == i && s2 == i && i < c0);  // next pivot (i,i); at i == c0 that entry is (x,x)
                double v = (C[rs] + alc * ar * as) + al2 * (ar * cs + cr * as) + al2 * (br * bs) + al4 * (ar * bs + br * as);
                if (clamp) v = cy_max(v, kMinVal);
                C[rs] = v;
                v = (B[rs] + al4 * ar * as) + al2 * (ar * bs + br * as);
                if (clamp) v = cy_max(v, kMinVal);
                B[rs] = v;
                v = A[rs] + al2 * ar * as;
                if (clamp) v = cy_max(v, kMinVal);
                A[rs] = v;
            }
    }
    const int f = t2_fin(c0), yy = tri(c0, c0);
    out[f] = A[yy]; out[f + 1] = B[yy]; out[f + 2] = C[yy];
    out[f + 3] = trP; out[f + 4] = trPP; out[f + 5] = row0[3 * T0 + 2]; out[f + 6] = logdet;
    out[f + 7] = row0[3 * T0]; out[f + 8] = row0[3 * T0 + 1];
}

// The model without any x ("null model", lmm/lmm.py:176-190): level c0 of the [W0, y] block is all there is.
PG_HD void null_eval_from_row(const double* fin, EvalOut* out)
{
    out->yPy = fin[0]; out->yPPy = fin[1]; out->yPPPy = fin[2];
    out->trP = fin[3]; out->trPP = fin[4]; out->logdetH = fin[5]; out->logdetWHW = fin[6];
    out->trH = fin[7]; out->trHH = fin[8];
    out->xPx = NAN; out->yPx = NAN;
}

// last level: pivot on x itself.  (app, bpp, cpp) = x diagonal after the covariate levels (clamped), (ar, br, cr) = the
// (y, x) entries, fin = finals of the table-2 row.
template <bool FULL>
PG_HD void xrow_final_level(const double* fin, double app, double bpp, double cpp, double ar, double br, double cr,
                            bool need_logdet, EvalOut* out)
{
    out->xPx = app;  // pyx:1529-1533
    out->yPx = ar;
    const double inv = 1.0 / app;
    const double al2 = -inv, al4 = bpp / (app * app);
    double trPP = fin[4], alc = 0.0;
    if (FULL) {
        trPP = trPP + (bpp / app) * (bpp / app) - 2 * (cpp / app);
        alc = (cpp / (app * app)) - ((bpp * bpp) / (app * app * app));
    }
    const double trP = fin[3] - bpp / app;
    double logdet = fin[6];
    if (need_logdet) logdet += log(app);
    double vc = NAN;
    if (FULL) {
        vc = (fin[2] + alc * ar * ar) + al2 * (ar * cr + cr * ar) + al2 * (br * br) + al4 * (ar * br + br * ar);
        vc = cy_max(vc, kMinVal);
    }
    double vb = (fin[1] + al4 * ar * ar) + al2 * (ar * br + br * ar);
    vb = cy_max(vb, kMinVal);
    double va = fin[0] + al2 * ar * ar;
    va = cy_max(va, kMinVal);
    out->yPy = va; out->yPPy = vb; out->yPPPy = vc;
    out->trP = trP; out->trPP = FULL ? trPP : NAN;
    out->logdetH = fin[5]; out->logdetWHW = logdet;
    out->trH = fin[7]; out->trHH = fin[8];
}

// scalar form (host tests, probes): xa/xb/xc hold the level-0 x row, entries j < c0: x.w_j, c0: x.y, c0+1: x.x
// Role swap ("de" mode, reference lmm/lmm.py:498-532 calculate_de): the per-column vector x is the PHENOTYPE and the
// fixed column y the tested regressor, i.e. the reference's W_star is [W0, y, x].  The covariate levels are the same;
// the last pivot is (y, y) -- clamped like every pivot diagonal (pyx:953,:961,:1016; only A at c0 = 0, pyx:939) -- and the
// x diagonal, which is no pivot now, is clamped only after that level (pyx:961 at i = c).  xPx / yPx of EvalOut then hold
// (y, y) and (x, y) at level c0 and yPy.. the x diagonal at level c0 + 1, so SnpSolver::finish gives the effect of y on x.
template <bool FULL>
PG_HD void xrow_final_level_swapped(int c0, const double* fin, double xxa, double xxb, double xxc, double ar, double br,
                                    double cr, bool need_logdet, EvalOut* out);

template <bool FULL>
PG_HD_NOINLINE void xrow_recursion_scalar(int c0, const double* row2, double* xa, double* xb, double* xc,
                                          bool need_logdet, EvalOut* out, bool swap = false)
{
    const int Tp = t2_pairs(c0), dg = c0 + 1;
    if (c0 == 0 && !swap) xa[dg] = cy_max(xa[dg], kMinVal);
    for (int p = 0; p < c0; ++p) {
        const double al2 = row2[3 * p], al4 = row2[3 * p + 1], alc = row2[3 * p + 2];
        const double ar = xa[p], br = xb[p], cr = xc[p];
        for (int j = p + 1; j <= dg; ++j) {
            double as = ar, bs = br, cs = cr;
            if (j <= c0) {
                const int q = t2_col(c0, p, j);
                as = row2[q]; bs = row2[Tp + q]; cs = row2[2 * Tp + q];
            }
            const bool clamp = (p == c0 - 1 && j == dg) && !swap;
            if (FULL) {
                double v = (xc[j] + alc * ar * as) + al2 * (ar * cs + cr * as) + al2 * (br * bs) + al4 * (ar * bs + br * as);
                if (clamp) v = cy_max(v, kMinVal);
                xc[j] = v;
            }
            double v = (xb[j] + al4 * ar * as) + al2 * (ar * bs + br * as);
            if (clamp) v = cy_max(v, kMinVal);
            xb[j] = v;
            v = xa[j] + al2 * ar * as;
            if (clamp) v = cy_max(v, kMinVal);
            xa[j] = v;
        }
    }
    if (swap)
        xrow_final_level_swapped<FULL>(c0, row2 + t2_fin(c0), xa[dg], xb[dg], FULL ? xc[dg] : 0.0, xa[c0], xb[c0],
                                       FULL ? xc[c0] : 0.0, need_logdet, out);
    else
        xrow_final_level<FULL>(row2 + t2_fin(c0), xa[dg], xb[dg], FULL ? xc[dg] : 0.0, xa[c0], xb[c0], FULL ? xc[c0] : 0.0,
                               need_logdet, out);
}

template <bool FULL>
PG_HD void xrow_final_level_swapped(int c0, const double* fin, double xxa, double xxb, double xxc, double ar, double br,
                                    double cr, bool need_logdet, EvalOut* out)
{
    const double ya = cy_max(fin[0], kMinVal);
    const double yb = c0 ? cy_max(fin[1], kMinVal) : fin[1];
    const double yc = c0 ? cy_max(fin[2], kMinVal) : fin[2];
    const double fin2[9] = {xxa, xxb, xxc, fin[3], fin[4], fin[5], fin[6], fin[7], fin[8]};
    xrow_final_level<FULL>(fin2, ya, yb, yc, ar, br, cr, need_logdet, out);
}

// Chebyshev weights L[k] of `lam` inside its table interval (shared by both table kinds)
PG_HD void table_weights(const double* basis, double lam, int* interval, double* L /* kNodes */)
{
    double xloc;
    interval_of(lam, interval, &xloc);
    double T[kNodes];
    T[0] = 1.0; T[1] = xloc;
    for (int j = 2; j < kNodes; ++j) T[j] = 2.0 * xloc * T[j - 1] - T[j - 2];
    for (int k = 0; k < kNodes; ++k) {
        double s = 0.0;
        for (int j = 0; j < kNodes; ++j) s += basis[k * kNodes + j] * T[j];
        L[k] = s;
    }
}

// Host-side helpers shared by the C-ABI implementation and the CPU shim -------------------------------

inline void fill_tri_ab(TriAB* tab)
{
    for (int a = 0; a < kMaxCols; ++a)
        for (int b = 0; b <= a; ++b) {
            tab[tri(a, b)].a = (unsigned char)a;
            tab[tri(a, b)].b = (unsigned char)b;
        }
}

inline void fill_basis(double* basis /* kNodes*kNodes */)
{
    for (int k = 0; k < kNodes; ++k)
        for (int j = 0; j < kNodes; ++j)
            basis[k * kNodes + j] =
                (j == 0 ? 1.0 : 2.0) * cos(j * 3.14159265358979323846 * (k + 0.5) / kNodes) / kNodes;
}

// lambda of table row: rows 0..kNumFixed-1 are the fixed lambdas, then interval-major nodes
inline double table_lambda(int row)
{
    if (row < kNumFixed) return fixed_lambda(row);
    const int r = row - kNumFixed;
    return node_lambda(r / kNodes, r % kNodes);
}
constexpr int kNumTableRows = kNumFixed + kNumIntervals * kNodes;

}  // namespace pg
