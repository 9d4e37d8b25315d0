// tests/host_shim.cpp -- compiles the product's host/device scalar headers (pygemma_b200/csrc/pg_math.cuh,
// pg_eval.cuh) with g++ so that the optimiser state machine, the table interpolation, the Pab recursion
// and the F-distribution tail are unit-tested on the CPU box.  TEST CODE: never shipped, never used by
// the product path; the genotype-dependent dot products that the GPU does warp-collectively are plain
// loops here.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../pygemma_b200/csrc/compress_plan.h"
#include "../pygemma_b200/csrc/pg_eval.cuh"

using namespace pg;

namespace {

struct HostTables {
    std::vector<double> fixtab, itab, basis;
    std::vector<TriAB> tri_ab;
    Tables t;
};

// exact table rows: Gram of [W0, y] weighted by h^p, plus sum h, sum h^2, sum log(lam d + 1)
void build_tables(HostTables& H, int n, int c0, const double* d, const double* wy /* col-major n x (c0+1) */)
{
    const int k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2, NF = 3 * T0 + 3;
    H.fixtab.assign((size_t)kNumFixed * NF, 0.0);
    H.itab.assign((size_t)kNumIntervals * kNodes * NF, 0.0);
    H.basis.resize(kNodes * kNodes);
    H.tri_ab.resize(kMaxTri);
    fill_tri_ab(H.tri_ab.data());
    fill_basis(H.basis.data());
    std::vector<long double> acc(NF);
    for (int row = 0; row < kNumTableRows; ++row) {
        const double lam = table_lambda(row);
        std::fill(acc.begin(), acc.end(), 0.0L);
        for (int l = 0; l < n; ++l) {
            const double tt = lam * d[l] + 1.0, h = 1.0 / tt;
            const long double h1 = h, h2 = (long double)h * h, h3 = h2 * h;
            for (int r = 0; r < k0; ++r)
                for (int s = 0; s <= r; ++s) {
                    const long double p = (long double)wy[(size_t)r * n + l] * wy[(size_t)s * n + l];
                    const int q = tri(r, s);
                    acc[q] += p * h1; acc[T0 + q] += p * h2; acc[2 * T0 + q] += p * h3;
                }
            acc[3 * T0] += h1; acc[3 * T0 + 1] += h2; acc[3 * T0 + 2] += logl((long double)tt);
        }
        double* dst = row < kNumFixed ? &H.fixtab[(size_t)row * NF] : &H.itab[(size_t)(row - kNumFixed) * NF];
        for (int f = 0; f < NF; ++f) dst[f] = (double)acc[f];
    }
    H.t.c0 = c0; H.t.k0 = k0; H.t.T0 = T0; H.t.NF = NF;
    H.t.fixtab = H.fixtab.data(); H.t.itab = H.itab.data(); H.t.basis = H.basis.data();
    H.t.tri_ab = H.tri_ab.data();
}

void eval_snp(const HostTables& H, int n, const double* d, const double* wy, const double* x, double lam,
              int fixed_t, int full, int need_ll, bool exact_w0y, EvalOut* out)
{
    const int c0 = H.t.c0, k = c0 + 2, k0 = c0 + 1;
    const int TT = k * (k + 1) / 2;
    std::vector<double> A(TT, 0.0), B(TT, 0.0), C(TT, 0.0);
    Level0 l0;
    if (full) assemble_w0y<true>(H.t, lam, fixed_t, A.data(), B.data(), C.data(), &l0);
    else assemble_w0y<false>(H.t, lam, fixed_t, A.data(), B.data(), C.data(), &l0);
    std::vector<long double> a1(k0 + 1, 0.0L), a2(k0 + 1, 0.0L), a3(k0 + 1, 0.0L);
    for (int l = 0; l < n; ++l) {
        const double h = 1.0 / (lam * d[l] + 1.0);
        const long double xh = (long double)x[l] * h, xh2 = xh * h, xh3 = xh2 * h;
        for (int j = 0; j < k0; ++j) {
            const double w = wy[(size_t)j * n + l];
            a1[j] += xh * w; a2[j] += xh2 * w; a3[j] += xh3 * w;
        }
        a1[k0] += xh * x[l]; a2[k0] += xh2 * x[l]; a3[k0] += xh3 * x[l];
    }
    for (int j = 0; j < c0; ++j) {
        A[tri(c0, j)] = (double)a1[j]; B[tri(c0, j)] = (double)a2[j]; C[tri(c0, j)] = (double)a3[j];
    }
    A[tri(c0, c0)] = (double)a1[k0]; B[tri(c0, c0)] = (double)a2[k0]; C[tri(c0, c0)] = (double)a3[k0];
    A[tri(c0 + 1, c0)] = (double)a1[c0]; B[tri(c0 + 1, c0)] = (double)a2[c0]; C[tri(c0 + 1, c0)] = (double)a3[c0];
    if (exact_w0y && fixed_t < 0) {  // bypass the interpolation: exact [W0,y] block at this lambda
        long double s0 = 0, s1 = 0, s2 = 0;
        for (int r = 0; r < k0; ++r)
            for (int s = 0; s <= r; ++s) {
                long double p1 = 0, p2 = 0, p3 = 0;
                for (int l = 0; l < n; ++l) {
                    const double h = 1.0 / (lam * d[l] + 1.0);
                    const long double p = (long double)wy[(size_t)r * n + l] * wy[(size_t)s * n + l];
                    p1 += p * h; p2 += p * h * h; p3 += p * h * h * h;
                }
                const int rr = r == c0 ? c0 + 1 : r, ss = s == c0 ? c0 + 1 : s;
                A[tri(rr, ss)] = (double)p1; B[tri(rr, ss)] = (double)p2; C[tri(rr, ss)] = (double)p3;
            }
        for (int l = 0; l < n; ++l) {
            const double tt = lam * d[l] + 1.0, h = 1.0 / tt;
            s0 += h; s1 += (long double)h * h; s2 += logl((long double)tt);
        }
        l0.trP = (double)s0; l0.trPP = (double)s1; l0.logdetH = (double)s2;
    }
    if (full) pab_recursion<true>(H.t, A.data(), B.data(), C.data(), l0, need_ll, out);
    else pab_recursion<false>(H.t, A.data(), B.data(), C.data(), l0, need_ll, out);
}

}  // namespace

// one evaluation from the compressed moments z[j][k] (j = 0..c0-1: x.w_j, c0: x.y, c0+1: x.x) of one SNP
namespace {
void eval_snp_compressed(const HostTables& H, const CompressPlan& P, const double* z, double lam, int fixed_t,
                                int full, int need_ll, EvalOut* out)
{
    const int c0 = H.t.c0, k = c0 + 2, k1 = c0 + 2, Kc = P.Kc;
    const int TT = k * (k + 1) / 2;
    std::vector<double> A(TT, 0.0), B(TT, 0.0), C(TT, 0.0);
    Level0 l0;
    if (full) assemble_w0y<true>(H.t, lam, fixed_t, A.data(), B.data(), C.data(), &l0);
    else assemble_w0y<false>(H.t, lam, fixed_t, A.data(), B.data(), C.data(), &l0);
    for (int j = 0; j < k1; ++j) {
        double a1 = 0, a2 = 0, a3 = 0;
        for (int q = 0; q < Kc; ++q) {
            const double h = 1.0 / (lam * P.nodes[q] + 1.0), zz = z[(size_t)j * Kc + q];
            a1 += h * zz; a2 += h * h * zz; a3 += h * h * h * zz;
        }
        const int dst = j < c0 ? tri(c0, j) : (j == c0 ? tri(c0 + 1, c0) : tri(c0, c0));
        A[dst] = a1; B[dst] = a2; C[dst] = a3;
    }
    if (full) pab_recursion<true>(H.t, A.data(), B.data(), C.data(), l0, need_ll, out);
    else pab_recursion<false>(H.t, A.data(), B.data(), C.data(), l0, need_ll, out);
}
}  // namespace

namespace {
struct HostTables2 {
    std::vector<double> fix2, itab2;
    Tables2 t;
};

void build_tables2(const HostTables& H, HostTables2& H2)
{
    const int c0 = H.t.c0, NF = H.t.NF, NF2 = t2_nf(c0), T0 = H.t.T0;
    H2.fix2.assign((size_t)kNumFixed * NF2, 0.0);
    H2.itab2.assign((size_t)kNumIntervals * kNodes * NF2, 0.0);
    std::vector<double> work(3 * T0);
    for (int r = 0; r < kNumFixed; ++r) eliminate_w0y_row(c0, &H.fixtab[(size_t)r * NF], work.data(), &H2.fix2[(size_t)r * NF2]);
    for (int r = 0; r < kNumIntervals * kNodes; ++r)
        eliminate_w0y_row(c0, &H.itab[(size_t)r * NF], work.data(), &H2.itab2[(size_t)r * NF2]);
    H2.t.c0 = c0; H2.t.Tp = t2_pairs(c0); H2.t.NF2 = NF2;
    H2.t.fix2 = H2.fix2.data(); H2.t.itab2 = H2.itab2.data(); H2.t.basis = H.basis.data();
}

// evaluation through the x-row recursion + table-2 rows (what reml_solve_kernel does)
void eval_snp_xrow(const HostTables2& H2, const CompressPlan& P, const double* z, double lam, int fixed_t, int full,
                   int need_ll, EvalOut* out, bool swap = false)
{
    const int c0 = H2.t.c0, k1 = c0 + 2, Kc = P.Kc, NF2 = H2.t.NF2;
    std::vector<double> xa(k1), xb(k1), xc(k1), row(NF2);
    for (int j = 0; j < k1; ++j) {
        double a1 = 0, a2 = 0, a3 = 0;
        for (int q = 0; q < Kc; ++q) {
            const double h = 1.0 / (lam * P.nodes[q] + 1.0), zz = z[(size_t)j * Kc + q];
            a1 += h * zz; a2 += h * h * zz; a3 += h * h * h * zz;
        }
        xa[j] = a1; xb[j] = a2; xc[j] = a3;
    }
    const double* row2;
    if (fixed_t >= 0) {
        row2 = &H2.fix2[(size_t)fixed_t * NF2];
    } else {
        int iv;
        double L[kNodes];
        table_weights(H2.t.basis, lam, &iv, L);
        for (int f = 0; f < NF2; ++f) {
            double v = 0.0;
            for (int k = 0; k < kNodes; ++k) v += L[k] * H2.itab2[((size_t)iv * kNodes + k) * NF2 + f];
            row[f] = v;
        }
        row2 = row.data();
    }
    if (full) xrow_recursion_scalar<true>(c0, row2, xa.data(), xb.data(), xc.data(), need_ll, out, swap);
    else xrow_recursion_scalar<false>(c0, row2, xa.data(), xb.data(), xc.data(), need_ll, out, swap);
}

// table-2 finals at lambda (null model: no x anywhere)
void eval_null(const HostTables2& H2, double lam, int fixed_t, EvalOut* out)
{
    const int c0 = H2.t.c0, NF2 = H2.t.NF2;
    double fin[9];
    if (fixed_t >= 0) {
        for (int f = 0; f < 9; ++f) fin[f] = H2.fix2[(size_t)fixed_t * NF2 + t2_fin(c0) + f];
    } else {
        int iv;
        double L[kNodes];
        table_weights(H2.t.basis, lam, &iv, L);
        for (int f = 0; f < 9; ++f) {
            double v = 0.0;
            for (int k = 0; k < kNodes; ++k) v += L[k] * H2.itab2[((size_t)iv * kNodes + k) * NF2 + t2_fin(c0) + f];
            fin[f] = v;
        }
    }
    null_eval_from_row(fin, out);
}
}  // namespace

extern "C" {

// number of nodes the compression plan gives for the (ascending) eigenvalues d
int pgh_plan_nodes(int n, const double* d_sorted)
{
    CompressPlan P;
    build_compress_plan(d_sorted, n, &P);
    return P.Kc;
}

// The bookkeeping of the fused rotation's x^2 moments (compress_plan.h: build_fused_plan) replayed on the host: every
// eigen-index adds L_k a_l to the piece its row names, a segment's nodes are the sums of its pieces in order, COPY rows
// write their node.  Returns max |z_fused - z_direct| / max |z_direct| against the plain definition of the compressed
// moments; *npieces, *cnodes report the plan's sizes.  Also checks the structure the kernel relies on (a piece never
// spans two piece_eig chunks or two segments; pieces of a segment are consecutive): -1 on violation.
double pgh_fused_plan_check(int n, const double* d_sorted, const double* a, int tile_eig, int piece_eig, int* npieces,
                            int* cnodes)
{
    CompressPlan P;
    build_compress_plan(d_sorted, n, &P);
    FusedPlan F;
    build_fused_plan(P, tile_eig, piece_eig, &F);
    if (npieces) *npieces = F.npieces;
    if (cnodes) *cnodes = F.cnodes;
    if ((int)F.rows.size() % tile_eig != 0 || (int)F.rows.size() < n) return -1.0;
    for (size_t l = n; l < F.rows.size(); ++l)
        if (F.rows[l].kq != -1) return -1.0;
    std::vector<double> z(P.Kc, 0.0), zf(P.Kc, 0.0);
    for (const Segment& s : P.segs)
        for (int l = s.l0; l < s.l1; ++l) {
            if (s.type == kSegCopy) z[s.kb + (l - s.l0)] += a[l];
            else for (int q = 0; q < s.kq; ++q) z[s.kb + q] += P.Lw[(size_t)l * kCq + q] * a[l];
        }
    std::vector<double> part((size_t)std::max(F.npieces, 1) * kCq, 0.0);
    std::vector<int> piece_chunk(std::max(F.npieces, 1), -1), piece_seg(std::max(F.npieces, 1), -1);
    for (int l = 0; l < n; ++l) {
        const FusedRow r = F.rows[l];
        if (r.kq > 0) {
            if (r.id < 0 || r.id >= F.npieces) return -1.0;
            if (piece_chunk[r.id] >= 0 && piece_chunk[r.id] != l / piece_eig) return -1.0;
            if (piece_seg[r.id] >= 0 && piece_seg[r.id] != P.seg_of[l]) return -1.0;
            piece_chunk[r.id] = l / piece_eig; piece_seg[r.id] = P.seg_of[l];
            for (int k = 0; k < kCq; ++k) part[(size_t)r.id * kCq + k] += P.Lw[(size_t)l * kCq + k] * a[l];
        } else if (r.kq == 0) {
            if (r.id < 0 || r.id >= P.Kc) return -1.0;
            zf[r.id] = a[l];
        } else {
            return -1.0;
        }
    }
    int expect_pb = 0, nodes = 0;
    for (const FusedSeg& sr : F.segs) {
        if (sr.pb != expect_pb || sr.pe <= sr.pb) return -1.0;
        expect_pb = sr.pe;
        nodes += sr.kq;
        for (int p = sr.pb; p < sr.pe; ++p)
            for (int k = 0; k < sr.kq; ++k) zf[sr.kb + k] += part[(size_t)p * kCq + k];
    }
    if (expect_pb != F.npieces || nodes != F.cnodes) return -1.0;
    double worst = 0.0, scale = 0.0;
    for (int q = 0; q < P.Kc; ++q) scale = std::max(scale, fabs(z[q]));
    for (int q = 0; q < P.Kc; ++q) worst = std::max(worst, fabs(zf[q] - z[q]));
    return scale > 0.0 ? worst / scale : worst;
}

// worst |compressed - direct| / sum_l |a_l| h_l^p over a lambda sweep, for a_l given (sorted order)
double pgh_compress_error(int n, const double* d_sorted, const double* a, int power)
{
    CompressPlan P;
    build_compress_plan(d_sorted, n, &P);
    std::vector<long double> z(P.Kc, 0.0L);
    for (const Segment& s : P.segs)
        for (int l = s.l0; l < s.l1; ++l) {
            if (s.type == kSegCopy) z[s.kb + (l - s.l0)] += a[l];
            else for (int q = 0; q < s.kq; ++q) z[s.kb + q] += (long double)P.Lw[(size_t)l * kCq + q] * a[l];
        }
    double worst = 0;
    for (int i = 0; i <= 400; ++i) {
        const double lam = pow(10.0, -5.0 + i / 40.0);
        long double ex = 0, ab = 0, cp = 0;
        for (int l = 0; l < n; ++l) {
            const long double h = powl(1.0L / ((long double)lam * d_sorted[l] + 1.0L), power);
            ex += a[l] * h; ab += fabsl((long double)a[l]) * h;
        }
        for (int q = 0; q < P.Kc; ++q) cp += z[q] * powl(1.0L / ((long double)lam * P.nodes[q] + 1.0L), power);
        const double e = (double)(fabsl(cp - ex) / ab);
        if (e > worst) worst = e;
    }
    return worst;
}

// the scan on compressed moments: d / wy / xr in any eigenvalue order (sorted here)
void pgh_scan_compressed_ex(int n, int c0, long m, const double* d, const double* wy, const double* xr, int grid,
                            int xrow_form, double* out6, int32_t* status, int32_t* evals, int32_t* kc_out, int swap,
                            double* null3, double* lrt2);

void pgh_scan_compressed(int n, int c0, long m, const double* d, const double* wy, const double* xr, int grid,
                         int xrow_form, double* out6, int32_t* status, int32_t* evals, int32_t* kc_out)
{
    pgh_scan_compressed_ex(n, c0, m, d, wy, xr, grid, xrow_form, out6, status, evals, kc_out, 0, nullptr, nullptr);
}

// swap != 0: "de" mode (x-row form only).  null3 / lrt2 (nullable, x-row form only): the product's MlSolver on the null
// model -> {lambda_null, tau_null, l_null} and on every alternative model -> lrt2[2 g] = lambda_ml, lrt2[2 g + 1] = loglik_ml
void pgh_scan_compressed_ex(int n, int c0, long m, const double* d, const double* wy, const double* xr, int grid,
                            int xrow_form, double* out6, int32_t* status, int32_t* evals, int32_t* kc_out, int swap,
                            double* null3, double* lrt2)
{
    std::vector<int> perm(n);
    for (int l = 0; l < n; ++l) perm[l] = l;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return d[a] < d[b]; });
    const int k0 = c0 + 1, k1 = c0 + 2;
    std::vector<double> ds(n), wys((size_t)n * k0);
    for (int l = 0; l < n; ++l) {
        ds[l] = d[perm[l]];
        for (int j = 0; j < k0; ++j) wys[(size_t)j * n + l] = wy[(size_t)j * n + perm[l]];
    }
    HostTables H;
    build_tables(H, n, c0, ds.data(), wys.data());
    HostTables2 H2;
    if (xrow_form) build_tables2(H, H2);
    CompressPlan P;
    build_compress_plan(ds.data(), n, &P);
    if (kc_out) *kc_out = P.Kc;
    if (null3 && xrow_form) {
        MlSolver ms;
        ms.init(n);
        while (ms.pending()) {
            EvalOut e;
            eval_null(H2, ms.req_lambda(), ms.req_fixed(), &e);
            ms.feed(e);
        }
        null3[0] = ms.best_lambda; null3[1] = (double)n / ms.best_yPy; null3[2] = ms.best_ll;
    }
    std::vector<double> z((size_t)k1 * P.Kc);
    for (long g = 0; g < m; ++g) {
        std::fill(z.begin(), z.end(), 0.0);
        const double* x = xr + (size_t)g * n;
        for (const Segment& s : P.segs)
            for (int l = s.l0; l < s.l1; ++l) {
                const double xv = x[perm[l]];
                for (int j = 0; j < k1; ++j) {
                    const double a = j < k0 ? xv * wys[(size_t)j * n + l] : xv * xv;
                    if (s.type == kSegCopy) z[(size_t)j * P.Kc + s.kb + (l - s.l0)] += a;
                    else for (int q = 0; q < s.kq; ++q) z[(size_t)j * P.Kc + s.kb + q] += P.Lw[(size_t)l * kCq + q] * a;
                }
            }
        SnpSolver s;
        s.init(n, c0, grid);
        while (s.pending()) {
            EvalOut e;
            if (xrow_form) eval_snp_xrow(H2, P, z.data(), s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll(), &e, swap != 0);
            else eval_snp_compressed(H, P, z.data(), s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll(), &e);
            s.feed(e);
        }
        double* o = out6 + g * 6;
        o[0] = s.beta; o[1] = s.se; o[2] = s.tau; o[3] = s.lambda; o[4] = s.F; o[5] = s.p;
        status[g] = s.status; evals[2 * g] = s.n_eval2; evals[2 * g + 1] = s.n_eval3;
        if (lrt2 && xrow_form) {
            MlSolver ms;
            ms.init(n);
            while (ms.pending()) {
                EvalOut e;
                eval_snp_xrow(H2, P, z.data(), ms.req_lambda(), ms.req_fixed(), ms.req_full(), 0, &e);
                ms.feed(e);
            }
            lrt2[2 * g] = ms.best_lambda; lrt2[2 * g + 1] = ms.best_ll;
            status[g] |= ms.status;
        }
    }
}
}

extern "C" {

// SnpSolver + table/interpolation evaluator over a block of rotated SNPs (SNP-major xr)
void pgh_scan(int n, int c0, long m, const double* d, const double* wy, const double* xr, int grid,
              int exact_w0y, double* out6 /* m x 6 */, int32_t* status, int32_t* evals /* m x 2 */)
{
    HostTables H;
    build_tables(H, n, c0, d, wy);
    for (long g = 0; g < m; ++g) {
        SnpSolver s;
        s.init(n, c0, grid);
        while (s.pending()) {
            EvalOut e;
            eval_snp(H, n, d, wy, xr + (size_t)g * n, s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll(),
                     exact_w0y != 0, &e);
            s.feed(e);
        }
        double* o = out6 + g * 6;
        o[0] = s.beta; o[1] = s.se; o[2] = s.tau; o[3] = s.lambda; o[4] = s.F; o[5] = s.p;
        status[g] = s.status; evals[2 * g] = s.n_eval2; evals[2 * g + 1] = s.n_eval3;
    }
}

double pgh_f_sf(double F, double nu) { return f_sf_1(F, nu); }

// drive pg::Brent on f(x) = p0 + p1 x + p2 x^2 + p3 x^3 + p4 exp(-x); returns root, fills the query sequence
double pgh_brent(const double* p, double a, double b, double xtol, double rtol, int maxiter, double* xs, int cap,
                 int* ncalls)
{
    auto f = [&](double x) { return p[0] + x * (p[1] + x * (p[2] + x * p[3])) + p[4] * exp(-x); };
    int nc = 0;
    auto rec = [&](double x) { if (nc < cap) xs[nc] = x; nc++; return f(x); };
    Brent br;
    const double fa = rec(a), fb = rec(b);
    br.start(a, fa, b, fb, xtol, rtol, maxiter);
    while (!br.done) br.feed(rec(br.query()));
    *ncalls = nc;
    return br.root;
}

double pgh_interp_error(int n, int c0, const double* d, const double* wy, double lam, int power)
{
    // max over [W0,y] pairs of |interp - exact| / sqrt(G_rr G_ss) at this lambda
    HostTables H;
    build_tables(H, n, c0, d, wy);
    const int k = c0 + 2, TT = k * (k + 1) / 2, k0 = c0 + 1;
    std::vector<double> A(TT), B(TT), C(TT);
    Level0 l0;
    assemble_w0y<true>(H.t, lam, -1, A.data(), B.data(), C.data(), &l0);
    const double* M = power == 1 ? A.data() : (power == 2 ? B.data() : C.data());
    std::vector<long double> ex(k0 * k0, 0.0L);
    for (int r = 0; r < k0; ++r)
        for (int s = 0; s <= r; ++s)
            for (int l = 0; l < n; ++l) {
                const double h = 1.0 / (lam * d[l] + 1.0);
                ex[r * k0 + s] += (long double)wy[(size_t)r * n + l] * wy[(size_t)s * n + l] * powl(h, power);
            }
    double worst = 0;
    for (int r = 0; r < k0; ++r)
        for (int s = 0; s <= r; ++s) {
            const int rr = r == c0 ? c0 + 1 : r, ss = s == c0 ? c0 + 1 : s;
            const double e = fabs((double)((long double)M[tri(rr, ss)] - ex[r * k0 + s])) /
                             sqrt((double)(ex[r * k0 + r] * ex[s * k0 + s]));
            if (e > worst) worst = e;
        }
    return worst;
}
}
