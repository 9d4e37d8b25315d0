// pg_math.cuh -- scalar REML math shared by the CUDA kernels and the host-side unit tests.
//
// Everything here is plain fp64 scalar code marked PG_HD so that it compiles
// both under nvcc (device) and under g++ (tests/host_shim.cpp): the control
// flow of the per-SNP lambda optimiser is a resumable state machine
// (SnpSolver) that asks an evaluator for precompute_mat-style scalars and is
// fed the answers.  On the GPU the evaluator is a warp-collective pass over the
// rotated genotype vector; in the CPU tests it is a dense loop.
//
// Reference lines followed (paths relative to the reference repository):
//   pygemma_model/pygemma_model.pyx:64-194    calc_lambda_restricted (grid + default)
//   pygemma_model/pygemma_model.pyx:1349-1416 newton
//   pygemma_model/pygemma_model.pyx:1514-1537 calc_beta_vg_ve_restricted_overload
//   pygemma_model/pygemma_model.pyx:1656-1698 first/second derivative of the REML log-likelihood
//   pygemma_model/pygemma_model.pyx:1813-1830 REML log-likelihood
//   lmm/lmm.py:461-495                        calculate (F_wald, p_wald)
//   scipy.optimize.brentq (call site pyx:176-182), scipy.stats.f.sf (call site lmm.py:482)
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#define PG_HD_NOINLINE __host__ __device__
#else
#define PG_HD inline
#define PG_HD_NOINLINE inline
#endif

namespace pg {

constexpr double kMinVal = 1e-35;  // pyx:39
constexpr int kNumFixed = 11;      // lambda = 10^-5 .. 10^5
constexpr int kSubPerDecade = 8;   // interpolation sub-intervals per decade of lambda
constexpr int kNodes = 12;         // Chebyshev nodes per sub-interval
constexpr int kNumIntervals = 10 * kSubPerDecade;

// Cython lowers max(a, b) on C doubles to (b > a) ? b : a : a NaN in `a` survives.
PG_HD double cy_max(double a, double b) { return (b > a) ? b : a; }

// np.sign
PG_HD double np_sign(double v) { return isnan(v) ? v : (double)((v > 0) - (v < 0)); }

// correctly rounded 10^k for k = -5..5 (pow(10.0, k) in the reference, pyx:101,:157-158)
PG_HD double fixed_lambda(int t)
{
    switch (t) {
    case 0: return 1e-5;
    case 1: return 1e-4;
    case 2: return 1e-3;
    case 3: return 1e-2;
    case 4: return 1e-1;
    case 5: return 1.0;
    case 6: return 1e1;
    case 7: return 1e2;
    case 8: return 1e3;
    case 9: return 1e4;
    default: return 1e5;
    }
}

// Scalars of one precompute_mat call (pyx:880-1053) that the scan consumes.
struct EvalOut {
    double yPy, yPPy, yPPPy;  // level c_f = c0 + 1
    double trP, trPP;         // level c_f
    double logdetH, logdetWHW;
    double xPx, yPx;          // level c0 (Wald numerator / denominator)
    double trH, trHH;         // level 0: sum 1/(lam d + 1), sum 1/(lam d + 1)^2 (the ML derivatives use these)
};

// pyx:1656-1669 (c = number of fixed effects including the SNP)
PG_HD double reml_d1(double lam, int n, int c, double yPy_in, double yPPy, double trP)
{
    const double yPy = cy_max(yPy_in, kMinVal);
    double r = -0.5 * (((double)(n - c) - trP) / lam);
    r = r + 0.5 * (double)(n - c) * ((yPy - cy_max(yPPy, 0.0)) / lam) / yPy;
    return r;
}

// pyx:1675-1698
PG_HD double reml_d2(double lam, int n, int c, double yPy_in, double yPPy_in, double yPPPy_in, double trP,
                     double trPP)
{
    const double yPy = cy_max(yPy_in, kMinVal);
    const double yPPy = cy_max(yPPy_in, kMinVal);
    const double yPPPy = cy_max(yPPPy_in, kMinVal);
    const double lam2 = lam * lam;
    const double g2 = (yPy + yPPPy - 2 * yPPy) / lam2;
    const double g1 = (yPy - yPPy) / lam;
    double r = 0.5 * ((double)(n - c) + trPP - 2 * trP) / lam2;
    r = r - (double)(n - c) * ((g2 * yPy) - 0.5 * g1 * g1) / (yPy * yPy);
    return r;
}

// pyx:1813-1830 (logdet_Wt_W is the constant 0.0 of pyx:970)
PG_HD double reml_loglik(int n, int c, double yPy, double logdetH, double logdetWHW)
{
    const double df = (double)(n - c);
    double r = 0.5 * df * log(0.5 * df / 3.14159265358979323846);
    r = r - 0.5 * df;
    r = r - 0.5 * logdetH;
    r = r - 0.5 * logdetWHW;
    r = r - 0.5 * df * log(yPy);
    return r;
}

// ---- unrestricted (ML) log-likelihood in lambda: the functions behind the reference's likelihood-ratio scaffolding
// (lmm/lmm.py:22-84 calc_lambda, commented call sites :177,:189,:278-282).  c = all fixed effects of the model.
// pyx:1566-1581 likelihood_derivative1_lambda
PG_HD double ml_d1(double lam, int n, double yPy_in, double yPPy_in, double trH)
{
    double r = -0.5 * (((double)n - trH) / lam);
    const double num = cy_max(yPPy_in, kMinVal), denom = cy_max(yPy_in, kMinVal);
    r = r + (0.5 * (double)n) * (1.0 - num / denom) / lam;
    return r;
}

// pyx:1586-1603 likelihood_derivative2_lambda
PG_HD double ml_d2(double lam, int n, double yPy_in, double yPPy_in, double yPPPy_in, double trH, double trHH)
{
    const double yPy = cy_max(yPy_in, kMinVal), yPPy = cy_max(yPPy_in, kMinVal);
    const double g2 = (yPy + cy_max(yPPPy_in, kMinVal) - 2 * yPPy) / (lam * lam);
    const double g1 = (yPy - yPPy) / lam;
    double r = 0.5 * ((double)n + trHH - 2 * trH) / (lam * lam);
    r = r - 0.5 * (double)n * (2 * g2 - g1 * g1 / yPy) / yPy;
    return r;
}

// pyx:1542-1560 likelihood_lambda; identical to likelihood(lam, tau = n / yPy, beta_hat, ...) (pyx:1736-1755), which is
// what the commented LRT lines evaluate (lmm/lmm.py:189,:281)
PG_HD double ml_loglik(int n, double yPy_in, double logdetH)
{
    const double hn = 0.5 * (double)n;
    double r = hn * log((double)n / (2 * 3.14159265358979323846));
    r = r - hn;
    r = r - 0.5 * logdetH;
    r = r - hn * log(cy_max(yPy_in, kMinVal));
    return r;
}

// ------------------------------------------------------------------------------------------------
// scipy.optimize.brentq as a resumable iteration (xtol=2e-12, rtol=0.1, maxiter=100 at the call site).
// start() is given both end-point values (SciPy evaluates them itself, pyx:176; they are the same
// numbers the bracket scan already holds); next() yields the abscissa to evaluate, feed() takes f there.
// ------------------------------------------------------------------------------------------------
struct Brent {
    double xpre, xcur, xblk, fpre, fcur, fblk, spre, scur;
    double xtol, rtol;
    int iter, maxiter;
    int done;  // 1: root in `root`
    double root;

    PG_HD void start(double xa, double fa, double xb, double fb, double xtol_, double rtol_, int maxiter_)
    {
        xpre = xa; xcur = xb; fpre = fa; fcur = fb;
        xblk = 0.; fblk = 0.; spre = 0.; scur = 0.;
        xtol = xtol_; rtol = rtol_; maxiter = maxiter_;
        iter = 0; done = 0; root = xb;
        if (fpre == 0) { done = 1; root = xpre; return; }
        if (fcur == 0) { done = 1; root = xcur; return; }
        step();
    }
    // runs the loop body up to (not including) the function evaluation
    PG_HD void step()
    {
        if (iter >= maxiter) { done = 1; root = xcur; return; }
        iter++;
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) { done = 1; root = xcur; return; }
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            const double a = fabs(spre), b = 3 * fabs(sbis) - delta;
            const double lim = (a < b) ? a : b;
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
    }
    PG_HD double query() const { return xcur; }
    PG_HD void feed(double f)
    {
        fcur = f;
        step();
    }
};

// ------------------------------------------------------------------------------------------------
// Regularised incomplete beta / F(1, nu) survival function (scipy.stats.f.sf(F, 1, nu), lmm.py:482)
//   sf = I_{nu/(nu+F)}(nu/2, 1/2)
// Lentz continued fraction (Numerical-Recipes form) with the usual symmetry switch; the prefactor is
// assembled in log space so tiny p-values keep their relative accuracy.
// ------------------------------------------------------------------------------------------------
PG_HD_NOINLINE double betacf(double a, double b, double x)
{
    const double kTiny = 1e-300, kEps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < kTiny) d = kTiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 2000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < kTiny) d = kTiny;
        c = 1.0 + aa / c; if (fabs(c) < kTiny) c = kTiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < kTiny) d = kTiny;
        c = 1.0 + aa / c; if (fabs(c) < kTiny) c = kTiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < kEps) break;
    }
    return h;
}

// I_x(a, b) given both x and its complement y = 1 - x (passed separately to avoid cancellation)
PG_HD_NOINLINE double ibeta_xy(double a, double b, double x, double y)
{
    if (!(x > 0.0)) return 0.0;
    if (!(y > 0.0)) return 1.0;
    const double lnpre = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log(y);
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lnpre) * betacf(a, b, x) / a;
    return 1.0 - exp(lnpre) * betacf(b, a, y) / b;
}

// the same with lnbeta = lgamma(a + b) - lgamma(a) - lgamma(b) supplied by the caller (one value per scan when a, b are
// fixed: the difference of two lgammas of ~nu/2 loses 4-5 digits in double, so the host forms it in long double)
PG_HD_NOINLINE double ibeta_xy_pre(double a, double b, double x, double y, double lnbeta)
{
    if (!(x > 0.0)) return 0.0;
    if (!(y > 0.0)) return 1.0;
    const double lnpre = lnbeta + a * log(x) + b * log(y);
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lnpre) * betacf(a, b, x) / a;
    return 1.0 - exp(lnpre) * betacf(b, a, y) / b;
}

PG_HD_NOINLINE double f_sf_1_pre(double F, double nu, double lnbeta)
{
    if (isnan(F) || isnan(nu)) return F + nu;
    if (!(nu > 0.0)) return NAN;
    if (F <= 0.0) return 1.0;
    if (isinf(F)) return 0.0;
    const double x = nu / (nu + F), y = F / (nu + F);
    return ibeta_xy_pre(0.5 * nu, 0.5, x, y, lnbeta);
}

// chi-square(1) upper tail of the LRT statistic, p = 1 - chi2.cdf(D, 1) (reference lmm/lmm.py:300) = erfc(sqrt(D/2)),
// evaluated as the survival function so that small p-values keep their relative accuracy; D <= 0 gives 1
PG_HD double chi2_sf_1(double D)
{
    if (isnan(D)) return D;
    if (D <= 0.0) return 1.0;
    return erfc(sqrt(0.5 * D));
}

PG_HD_NOINLINE double f_sf_1(double F, double nu)
{
    if (isnan(F) || isnan(nu)) return F + nu;
    if (!(nu > 0.0)) return NAN;
    if (F <= 0.0) return 1.0;
    if (isinf(F)) return 0.0;
    const double x = nu / (nu + F), y = F / (nu + F);
    return ibeta_xy(0.5 * nu, 0.5, x, y);
}

// ------------------------------------------------------------------------------------------------
// Chebyshev node set used by the lambda-interpolation tables of SNP-independent quantities.
// Interval j covers log10(lambda) in [-5 + j/kSub, -5 + (j+1)/kSub]; nodes are first-kind
// Chebyshev points.  lagrange_basis() returns L_k(x) such that f(x) ~= sum_k L_k(x) f(x_k).
// ------------------------------------------------------------------------------------------------
PG_HD double cheb_node(int k) { return cos(3.14159265358979323846 * (k + 0.5) / kNodes); }

PG_HD void interval_of(double lam, int* interval, double* xloc)
{
    double t = (log10(lam) + 5.0) * kSubPerDecade;
    int j = (int)floor(t);
    if (j < 0) j = 0;
    if (j > kNumIntervals - 1) j = kNumIntervals - 1;
    *interval = j;
    *xloc = 2.0 * (t - j) - 1.0;
}

PG_HD double node_lambda(int interval, int k)
{
    const double t = (interval + 0.5 * (cheb_node(k) + 1.0)) / kSubPerDecade - 5.0;
    return pow(10.0, t);
}

// L_k(x) = (1/N) * sum_j eps_j T_j(x_k) T_j(x), eps_0 = 1, eps_j = 2
PG_HD void lagrange_basis(double x, double* L /* kNodes */)
{
    double T[kNodes];
    T[0] = 1.0; T[1] = x;
    for (int j = 2; j < kNodes; ++j) T[j] = 2.0 * x * T[j - 1] - T[j - 2];
    for (int k = 0; k < kNodes; ++k) {
        const double th = 3.14159265358979323846 * (k + 0.5) / kNodes;
        double s = T[0];
        for (int j = 1; j < kNodes; ++j) s += 2.0 * cos(j * th) * T[j];
        L[k] = s / kNodes;
    }
}

// ------------------------------------------------------------------------------------------------
// Per-SNP optimiser as a resumable state machine.  Usage:
//   SnpSolver s; s.init(n, c0, grid);
//   while (s.pending()) { EvalOut e = evaluate(s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll()); s.feed(e); }
//   s.beta ... s.p
// The 11 fixed-lambda evaluations (index t <-> lambda = 10^(t-5)) are requested first, in order; they do
// not depend on anything the optimiser decides, so a caller can serve them from precomputed dot products
// (phase F of the GPU scan) and hand the machine over to the streaming kernel only when the first
// SNP-specific lambda is requested (Brent / Newton / root likelihood).  The reference interleaves the same
// evaluations with the bracket handling (pyx:154-192); the values, comparisons and their order are the same.
// ------------------------------------------------------------------------------------------------
struct SnpSolver {
    enum Phase { kFixed = 0, kBrent = 1, kNewton = 2, kRootLL = 3, kDone = 4 };
    int n, c0, cf, grid;
    int phase;
    int t_req;  // fixed index being requested in kFixed
    int idx;    // current bracket [10^(idx-5), 10^(idx-4)], idx = 0..9
    double d1_fixed[kNumFixed];
    // best candidate so far (pyx:144-152,:186-192) with its Wald scalars
    double best_lambda, best_ll, best_xPx, best_yPx, best_yPy;
    Brent br;
    // newton (pyx:1349-1416)
    double nt_root;
    int nt_iter;
    // pending request
    double rq_lambda;
    int rq_fixed, rq_full, rq_ll;
    // outputs
    double lambda, beta, se, tau, F, p;
    int status, n_eval2, n_eval3;

    PG_HD int pending() const { return phase != kDone; }
    PG_HD double req_lambda() const { return rq_lambda; }
    PG_HD int req_fixed() const { return rq_fixed; }  // -1: lambda is SNP-specific
    PG_HD int req_full() const { return rq_full; }
    PG_HD int req_ll() const { return rq_ll; }

    PG_HD void request_fixed(int t)
    {
        phase = kFixed; t_req = t;
        rq_fixed = t; rq_lambda = fixed_lambda(t); rq_full = 0;
        rq_ll = grid ? 1 : (t == 0 || t == kNumFixed - 1);
    }
    PG_HD void request(double lam, int full, int ll) { rq_fixed = -1; rq_lambda = lam; rq_full = full; rq_ll = ll; }

    // defer_p != 0: finish() leaves p = NaN; the caller evaluates f_sf_1(F, n - c0 - 1) itself (the GPU scan does it one
    // thread per SNP in a follow-up kernel instead of 32 lanes in lock step)
    PG_HD void init(int n_, int c0_, int grid_, int defer_p_ = 0)
    {
        defer_p = defer_p_;
        n = n_; c0 = c0_; cf = c0_ + 1; grid = grid_;
        idx = 0; status = 0; n_eval2 = 0; n_eval3 = 0;
        best_lambda = 0; best_ll = 0; best_xPx = best_yPx = best_yPy = 0;
        nt_root = 0; nt_iter = 0;
        lambda = beta = se = tau = F = p = NAN;
        br.done = 0;
        grid_ll0 = 0; grid_in_ll = -INFINITY; grid_in_lam = 0;
        grid_w0[0] = grid_w0[1] = grid_w0[2] = 0; grid_in_w[0] = grid_in_w[1] = grid_in_w[2] = 0;
        for (int k = 0; k < kNumFixed; ++k) d1_fixed[k] = 0;
        request_fixed(0);
    }

    PG_HD void set_best(double ll, double lam, const EvalOut& e)
    {
        best_ll = ll; best_lambda = lam; best_xPx = e.xPx; best_yPx = e.yPx; best_yPy = e.yPy;
    }

    PG_HD void finish()
    {
        phase = kDone;
        lambda = best_lambda;
        const int df = n - c0 - 1;
        // pyx:1529-1535, lmm.py:471,:482
        beta = best_yPx / best_xPx;
        se = sqrt(best_yPy) / (sqrt(cy_max(best_xPx, kMinVal)) * sqrt((double)df));
        tau = (double)df / best_yPy;
        const double z = beta / se;
        F = z * z;
        p = defer_p ? NAN : f_sf_1(F, (double)df);
        if (status & 1) { lambda = beta = se = tau = F = p = NAN; }   // bit 0: fatal; bits 1, 2: notes (MlSolver / skipped bracket)
    }

    PG_HD void start_newton(double lam0)
    {
        nt_root = lam0; nt_iter = 0;
        phase = kNewton;
        request(nt_root, 1, 0);
    }

    // scan brackets idx.. for a sign change of d1 (pyx:154-174); issue the first Brent request or finish
    PG_HD void next_bracket()
    {
        for (; idx < kNumFixed - 1; ++idx) {
            const double f0 = d1_fixed[idx], f1 = d1_fixed[idx + 1];
            if (copysign(1.0, f0) * copysign(1.0, f1) < 0) {  // pyx:174
                // a NaN derivative at a bracket end: the bracket is skipped and noted (status bit 2, not fatal) -- the other
                // brackets and the two boundary candidates still compete; only rows without any finite candidate end as NaN
                if (isnan(f0) || isnan(f1)) { status |= 4; continue; }
                br.start(fixed_lambda(idx), f0, fixed_lambda(idx + 1), f1, 2e-12, 0.1, 100);
                if (br.done) { start_newton(br.root); return; }
                phase = kBrent;
                request(br.query(), 0, 0);
                return;
            }
        }
        finish();
    }

    // number of brackets with a sign change (valid once the fixed phase is complete)
    PG_HD int count_brackets() const
    {
        int c = 0;
        for (int k = 0; k < kNumFixed - 1; ++k)
            if (copysign(1.0, d1_fixed[k]) * copysign(1.0, d1_fixed[k + 1]) < 0) c++;
        return c;
    }

    PG_HD void feed(const EvalOut& e)
    {
        if (rq_full) n_eval3++; else n_eval2++;
        if (phase == kFixed) {
            const int t = t_req;
            d1_fixed[t] = reml_d1(rq_lambda, n, cf, e.yPy, e.yPPy, e.trP);
            if (rq_ll) {
                const double ll = reml_loglik(n, cf, e.yPy, e.logdetH, e.logdetWHW);
                if (t == 0) {
                    set_best(ll, rq_lambda, e);  // pyx:144 / :109
                } else if (t == kNumFixed - 1) {
                    // the upper boundary beats the lower one on '<' (pyx:148 / :113); in grid mode the interior
                    // points were compared with '>' against the running best, which started at the lower
                    // boundary, so the upper boundary enters with the reference's boundary rule below
                    if (grid) {
                        // grid: best so far = argmax over {lo} U {10^-4..10^4} with strict '>' in lambda order;
                        // reference order is {lo, hi} first, then 10^-5..10^4.  Reproduce it exactly:
                        // hi wins over lo iff ll_lo < ll_hi; an interior point wins iff strictly greater than that.
                        const double ll_lo = grid_ll0;
                        double b_ll = ll_lo, b_lam = fixed_lambda(0);
                        double b_w[3] = {grid_w0[0], grid_w0[1], grid_w0[2]};
                        if (b_ll < ll) { b_ll = ll; b_lam = rq_lambda; b_w[0] = e.xPx; b_w[1] = e.yPx; b_w[2] = e.yPy; }
                        if (grid_in_ll > b_ll) { b_ll = grid_in_ll; b_lam = grid_in_lam; b_w[0] = grid_in_w[0]; b_w[1] = grid_in_w[1]; b_w[2] = grid_in_w[2]; }
                        best_ll = b_ll; best_lambda = b_lam; best_xPx = b_w[0]; best_yPx = b_w[1]; best_yPy = b_w[2];
                    } else if (best_ll < ll) {
                        set_best(ll, rq_lambda, e);
                    }
                } else if (grid) {
                    // running argmax over the interior grid points in increasing lambda, strict '>' (pyx:128)
                    if (ll > grid_in_ll) {
                        grid_in_ll = ll; grid_in_lam = rq_lambda;
                        grid_in_w[0] = e.xPx; grid_in_w[1] = e.yPx; grid_in_w[2] = e.yPy;
                    }
                }
                if (t == 0 && grid) { grid_ll0 = ll; grid_w0[0] = e.xPx; grid_w0[1] = e.yPx; grid_w0[2] = e.yPy; }
            }
            if (t + 1 < kNumFixed) { request_fixed(t + 1); return; }
            if (grid) { finish(); return; }
            idx = 0;
            next_bracket();
            return;
        }
        if (phase == kBrent) {
            const double f = reml_d1(rq_lambda, n, cf, e.yPy, e.yPPy, e.trP);
            if (isnan(f)) { status |= 4; idx++; next_bracket(); return; }   // as above: give up on this bracket only
            br.feed(f);
            if (br.done) { start_newton(br.root); return; }
            request(br.query(), 0, 0);
            return;
        }
        if (phase == kNewton) {
            const double d1 = reml_d1(nt_root, n, cf, e.yPy, e.yPPy, e.trP);
            const double d2 = reml_d2(nt_root, n, cf, e.yPy, e.yPPy, e.yPPPy, e.trP, e.trPP);
            const double ratio = d1 / d2;
            bool stop = false;
            if (np_sign(ratio) * np_sign(d1) * np_sign(d2) <= 0.0) stop = true;  // pyx:1392
            if (!stop) {
                const double lambda_new = nt_root - ratio;
                const double r_eps = fabs(lambda_new - nt_root) / fabs(nt_root);
                if (lambda_new < fixed_lambda(idx)) stop = true;           // pyx:1398-1400 (clamp discarded)
                else if (lambda_new > fixed_lambda(idx + 1)) stop = true;  // pyx:1402-1404
                else if (isnan(lambda_new) || isinf(lambda_new)) stop = true;  // pyx:1406
                else {
                    nt_root = lambda_new;
                    if (r_eps < 1e-5 || nt_iter > 100) stop = true;  // pyx:1411
                    else nt_iter++;
                }
            }
            if (!stop) { request(nt_root, 1, 0); return; }
            phase = kRootLL;
            request(nt_root, 0, 1);  // pyx:186
            return;
        }
        if (phase == kRootLL) {
            const double ll = reml_loglik(n, cf, e.yPy, e.logdetH, e.logdetWHW);
            if (ll > best_ll) set_best(ll, rq_lambda, e);  // pyx:190
            idx++;
            next_bracket();
            return;
        }
    }

    int defer_p;
    // grid-mode bookkeeping (lower boundary and running interior argmax)
    double grid_ll0, grid_in_ll, grid_in_lam;
    double grid_w0[3], grid_in_w[3];
};

// ------------------------------------------------------------------------------------------------
// The ML lambda of one model as lmm.calc_lambda finds it (reference lmm/lmm.py:22-84), resumable like SnpSolver:
//   * d1 = likelihood_derivative1_lambda at lambda = 10^-5 .. 10^5, reused between neighbouring brackets (:47-61);
//   * a bracket counts when np.sign(d1_lo) * np.sign(d1_hi) < 0 (:63) -- NaN never does;
//   * scipy.optimize.brentq(rtol = 0.1, maxiter = 5000) (:64-69), then scipy.optimize.newton with fprime = d2,
//     rtol = 1e-5, tol = 1.48e-8 (SciPy default), maxiter = 10, disp = False (:71-76): at most ten steps
//     p <- p - d1/d2, stopping when d1 == 0, d2 == 0 or np.isclose(p_new, p_old, rtol, atol = tol), no bracket clamp;
//   * roots = [1e-5, 1e5, newton results...]; the answer is roots[np.argmax(likelihood_lambda(roots))] (:81-84): the
//     first maximum wins, and a NaN likelihood counts as a maximum (np.argmax returns the first NaN).
// One deviation: the reference evaluates its closed forms at whatever lambda Newton produces; the tables here span
// [1e-5, 1e5], so an iterate outside that range (or not finite) ends Newton at the last iterate inside it and sets
// status bit 1.  It does not happen after a 10 %-accurate Brent start on the test problems.
// The same machine serves the null model (its evaluator has no x) and every alternative model [W0, x].
// ------------------------------------------------------------------------------------------------
struct MlSolver {
    enum Phase { kFixed = 0, kBrent = 1, kNewton = 2, kRootLL = 3, kDone = 4 };
    int n, phase, t_req, idx;
    double d1_fixed[kNumFixed];
    double best_lambda, best_ll, best_yPy;
    Brent br;
    double nt_p0;
    int nt_iter;
    double rq_lambda;
    int rq_fixed, rq_full;
    int status, n_eval2, n_eval3;

    PG_HD int pending() const { return phase != kDone; }
    PG_HD double req_lambda() const { return rq_lambda; }
    PG_HD int req_fixed() const { return rq_fixed; }
    PG_HD int req_full() const { return rq_full; }

    PG_HD void request_fixed(int t) { phase = kFixed; t_req = t; rq_fixed = t; rq_lambda = fixed_lambda(t); rq_full = 0; }
    PG_HD void request(double lam, int full) { rq_fixed = -1; rq_lambda = lam; rq_full = full; }

    PG_HD void init(int n_)
    {
        n = n_; idx = 0; status = 0; n_eval2 = 0; n_eval3 = 0;
        best_lambda = 0; best_ll = 0; best_yPy = 0; nt_p0 = 0; nt_iter = 0; br.done = 0;
        for (int k = 0; k < kNumFixed; ++k) d1_fixed[k] = 0;
        request_fixed(0);
    }

    // np.argmax over the growing list: first maximum, NaN beats everything and the first NaN stays
    PG_HD void candidate(double ll, double lam, double yPy)
    {
        if (isnan(best_ll)) return;
        if (isnan(ll) || ll > best_ll) { best_ll = ll; best_lambda = lam; best_yPy = yPy; }
    }

    PG_HD void next_bracket()
    {
        for (; idx < kNumFixed - 1; ++idx) {
            const double f0 = d1_fixed[idx], f1 = d1_fixed[idx + 1];
            if (np_sign(f0) * np_sign(f1) < 0) {  // lmm.py:63
                br.start(fixed_lambda(idx), f0, fixed_lambda(idx + 1), f1, 2e-12, 0.1, 5000);
                if (br.done) { start_newton(br.root); return; }
                phase = kBrent;
                request(br.query(), 0);
                return;
            }
        }
        phase = kDone;
    }

    PG_HD void start_newton(double lam0) { nt_p0 = lam0; nt_iter = 0; phase = kNewton; request(nt_p0, 1); }

    PG_HD static bool in_tables(double lam) { return lam >= fixed_lambda(0) && lam <= fixed_lambda(kNumFixed - 1); }

    PG_HD void feed(const EvalOut& e)
    {
        if (rq_full) n_eval3++; else n_eval2++;
        if (phase == kFixed) {
            const int t = t_req;
            d1_fixed[t] = ml_d1(rq_lambda, n, e.yPy, e.yPPy, e.trH);
            if (t == 0) {
                best_ll = ml_loglik(n, e.yPy, e.logdetH); best_lambda = rq_lambda; best_yPy = e.yPy;   // roots[0]
            } else if (t == kNumFixed - 1) {
                candidate(ml_loglik(n, e.yPy, e.logdetH), rq_lambda, e.yPy);                          // roots[1]
            }
            if (t + 1 < kNumFixed) { request_fixed(t + 1); return; }
            idx = 0;
            next_bracket();
            return;
        }
        if (phase == kBrent) {
            br.feed(ml_d1(rq_lambda, n, e.yPy, e.yPPy, e.trH));
            if (br.done) { start_newton(br.root); return; }
            request(br.query(), 0);
            return;
        }
        if (phase == kNewton) {
            // scipy.optimize.newton, Newton-Raphson branch
            const double fval = ml_d1(nt_p0, n, e.yPy, e.yPPy, e.trH);
            const double fder = ml_d2(nt_p0, n, e.yPy, e.yPPy, e.yPPPy, e.trH, e.trHH);
            double root = nt_p0;
            bool stop = (fval == 0) || (fder == 0);
            if (!stop) {
                const double p = nt_p0 - fval / fder;
                const bool fin = isfinite(p) && isfinite(nt_p0);
                const bool close = fin ? (fabs(p - nt_p0) <= 1.48e-8 + 1e-5 * fabs(nt_p0)) : (p == nt_p0);
                if (!(isfinite(p) && in_tables(p))) { status |= 2; stop = true; }   // deviation, see above
                else {
                    root = p;
                    if (close || nt_iter + 1 >= 10) stop = true;
                    else { nt_p0 = p; nt_iter++; }
                }
            }
            if (!stop) { request(nt_p0, 1); return; }
            phase = kRootLL;
            request(root, 0);
            return;
        }
        if (phase == kRootLL) {
            candidate(ml_loglik(n, e.yPy, e.logdetH), rq_lambda, e.yPy);
            idx++;
            next_bracket();
            return;
        }
    }
};

}  // namespace pg
