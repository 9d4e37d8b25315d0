/*
 * oracle/reml_oracle.c -- CPU restatement of pyGEMMA's per-SNP REML scan.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pygemma_b200/) may
 * link, import or call this file; it exists so that tests/, the smoke check
 * and bench.py's cpu_baseline leg have an independent fp64 statement of the
 * reference algorithm to compare the CUDA path against.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle.py) against
 * golden vectors produced by the reference's own Cython code compiled in the
 * build container (oracle/build_ref.py -> oracle/_ref, "ref64" = the reference
 * sources with float32 promoted to float64; tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it follows
 * (paths relative to the reference repository root).
 *
 * Conventions: W* = [W0 (c0 columns), x, y] has k = c0 + 2 columns; matrices
 * are k x k row-major and only the lower triangle (row >= col) is meaningful,
 * like the reference's lower=1 BLAS calls.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#ifndef PGO_ACC_T
#define PGO_ACC_T long double /* accumulate Gram sums wider than the thing under test */
#endif
typedef PGO_ACC_T acc_t;

#define PGO_MIN_VAL 1e-35 /* pygemma_model/pygemma_model.pyx:39 */

/* Cython lowers max(a, b) on C doubles to (b > a) ? b : a : NaN in `a` survives. */
static inline double cy_max(double a, double b) { return (b > a) ? b : a; }

typedef struct {
    int n, c0, k;
    const double *d;  /* eigenvalues, n */
    const double *w0; /* rotated covariates, column-major n x c0 */
    const double *y;  /* rotated phenotype, n */
    const double *x;  /* rotated genotype of the SNP under test, n */
    long n_eval2, n_eval3;
} pgo_ctx;

typedef struct {
    /* final-level and Wald-level scalars of one precompute_mat call */
    double yPy, yPPy, yPPPy; /* level c_f = c0+1 (W0 and x projected out) */
    double trP, trPP;        /* level c_f */
    double logdet_H, logdet_WHW;
    double xPx, yPx; /* level c0: wjt_Pi_wk[c0,c0,c0], wjt_Pi_wk[c0+1,c0,c0] */
} pgo_eval;

static inline const double *col(const pgo_ctx *c, int j)
{
    if (j < c->c0) return c->w0 + (size_t)j * c->n;
    return (j == c->c0) ? c->x : c->y;
}

/*
 * precompute_mat (pygemma_model/pygemma_model.pyx:880-1053).
 * full=0 follows :934-975, full=1 follows :977-1053.
 * If levels != NULL it receives, for each level i = 0..c_f:
 *   levels[i*5 + {0,1,2,3,4}] = yt_Pi_y[i], yt_Pi_Pi_y[i], yt_Pi_Pi_Pi_y[i], tr_Pi[i], tr_Pi_Pi[i]
 * and if Aout != NULL the level-c0 lower triangle of wjt_Pi_wk is not kept; only
 * the scalars the scan consumes are returned.
 */
static void precompute(pgo_ctx *c, double lam, int full, pgo_eval *out, double *levels)
{
    const int n = c->n, k = c->k, cf = c->c0 + 1, c0 = c->c0;
    double *A = (double *)malloc(sizeof(double) * 3 * k * k);
    double *B = A + k * k, *C = B + k * k;
    acc_t *acc = (acc_t *)calloc((size_t)3 * k * k + 4, sizeof(acc_t));
    acc_t *accA = acc, *accB = acc + k * k, *accC = accB + k * k;
    acc_t s1 = 0, s2 = 0, sl = 0;

    /* :903  Hi = 1/(lam*d + 1);  :938,:943,:1002 level-0 Gram matrices; :946,:1006 traces */
    for (int l = 0; l < n; ++l) {
        const double t = lam * c->d[l] + 1.0;
        const double h = 1.0 / t;
        const double h2 = h * h, h3 = h2 * h;
        s1 += h;
        s2 += (acc_t)h * h;
        sl += log(t); /* :972 logdet_H */
        for (int r = 0; r < k; ++r) {
            const double wr = col(c, r)[l];
            for (int s = 0; s <= r; ++s) {
                const acc_t p = (acc_t)wr * col(c, s)[l];
                accA[r * k + s] += p * h;
                accB[r * k + s] += p * h2;
                if (full) accC[r * k + s] += p * h3;
            }
        }
    }
    for (int i = 0; i < k * k; ++i) {
        A[i] = (double)accA[i];
        B[i] = (double)accB[i];
        C[i] = (double)accC[i];
    }
    free(acc);
    if (full) c->n_eval3++; else c->n_eval2++;

    double trP = (double)s1, trPP = (double)s2, logdet_WHW = 0.0;
    A[0] = cy_max(A[0], PGO_MIN_VAL); /* :939 / :993 (only temp_Pi is clamped at level 0) */
    if (levels) {
        const int y = k - 1;
        levels[0] = A[y * k + y]; levels[1] = B[y * k + y]; levels[2] = full ? C[y * k + y] : NAN;
        levels[3] = trP; levels[4] = full ? trPP : NAN;
    }

    for (int i = 1; i <= cf; ++i) {
        const int p = i - 1;
        const double app = A[p * k + p], bpp = B[p * k + p], cpp = C[p * k + p];
        if (i == cf) { /* level c0 is complete before the x pivot is applied: Wald inputs (:1529-1533) */
            out->xPx = A[c0 * k + c0];
            out->yPx = A[(c0 + 1) * k + c0];
        }
        if (full) {
            /* :1008-1009 */
            trPP = trPP + pow(bpp / app, 2.0) - 2 * (cpp / app);
            /* :1011-1014 */
            const double al1 = (cpp / pow(app, 2.0)) - (pow(bpp, 2.0) / pow(app, 3.0));
            const double al2 = -1.0 / app;
            const double al4 = bpp / pow(app, 2.0);
            for (int r = i; r < k; ++r)
                for (int s = i; s <= r; ++s) {
                    const double ar = A[r * k + p], as = A[s * k + p];
                    const double br = B[r * k + p], bs = B[s * k + p];
                    const double cr = C[r * k + p], cs = C[s * k + p];
                    C[r * k + s] = (C[r * k + s] + al1 * ar * as) + al2 * (ar * cs + cr * as)
                                   + al2 * (br * bs) + al4 * (ar * bs + br * as);
                }
            C[i * k + i] = cy_max(C[i * k + i], PGO_MIN_VAL); /* :1016 */
        }
        trP = trP - bpp / app; /* :948 / :1020 */
        {
            /* :950-951 / :1022-1023 */
            const double al1 = bpp / pow(app, 2.0), al2 = -1.0 / app;
            for (int r = i; r < k; ++r)
                for (int s = i; s <= r; ++s) {
                    const double ar = A[r * k + p], as = A[s * k + p];
                    const double br = B[r * k + p], bs = B[s * k + p];
                    B[r * k + s] = (B[r * k + s] + al1 * ar * as) + al2 * (ar * bs + br * as);
                }
            B[i * k + i] = cy_max(B[i * k + i], PGO_MIN_VAL); /* :953 / :1025 */
        }
        logdet_WHW += log(app); /* :957 / :1029 */
        {
            /* :959 / :1031 */
            const double al = -1.0 / app;
            for (int r = i; r < k; ++r)
                for (int s = i; s <= r; ++s)
                    A[r * k + s] = A[r * k + s] + al * A[r * k + p] * A[s * k + p];
            A[i * k + i] = cy_max(A[i * k + i], PGO_MIN_VAL); /* :961 / :1034 */
        }
        if (levels) {
            const int y = k - 1;
            levels[i * 5 + 0] = A[y * k + y]; levels[i * 5 + 1] = B[y * k + y];
            levels[i * 5 + 2] = full ? C[y * k + y] : NAN;
            levels[i * 5 + 3] = trP; levels[i * 5 + 4] = full ? trPP : NAN;
        }
    }
    const int y = k - 1;
    out->yPy = A[y * k + y];  /* yt_Pi_y[c_f]      :967 / :1045 */
    out->yPPy = B[y * k + y]; /* yt_Pi_Pi_y[c_f]   :968 / :1046 */
    out->yPPPy = full ? C[y * k + y] : NAN; /* :1047 */
    out->trP = trP; out->trPP = full ? trPP : NAN;
    out->logdet_H = (double)sl; out->logdet_WHW = logdet_WHW;
    free(A);
}

/* likelihood_derivative1_restricted_lambda_overload (pygemma_model.pyx:1656-1669); c = c_f */
double pgo_d1(double lam, int n, int c, double yPy_in, double yPPy, double trP)
{
    const double yPy = cy_max(yPy_in, PGO_MIN_VAL);
    double r = -0.5 * ((n - c - trP) / lam);
    r = r + 0.5 * (n - c) * ((yPy - cy_max(yPPy, 0)) / lam) / yPy;
    return r;
}

/* likelihood_derivative2_restricted_lambda_overload (pygemma_model.pyx:1675-1698) */
double pgo_d2(double lam, int n, int c, double yPy_in, double yPPy_in, double yPPPy_in, double trP, double trPP)
{
    const double yPy = cy_max(yPy_in, PGO_MIN_VAL);
    const double yPPy = cy_max(yPPy_in, PGO_MIN_VAL);
    const double yPPPy = cy_max(yPPPy_in, PGO_MIN_VAL);
    const double g2 = (yPy + yPPPy - 2 * yPPy) / pow(lam, 2.0);
    const double g1 = (yPy - yPPy) / lam;
    double r = 0.5 * (n - c + trPP - 2 * trP) / pow(lam, 2.0);
    r = r - (n - c) * ((g2 * yPy) - 0.5 * g1 * g1) / pow(yPy, 2.0);
    return r;
}

/* likelihood_restricted_lambda_overload (pygemma_model.pyx:1813-1830); logdet_Wt_W is 0.0 (:970) */
double pgo_loglik(int n, int c, double yPy, double logdet_H, double logdet_WHW)
{
    double r = 0.5 * (n - c) * log(0.5 * (n - c) / M_PI);
    r = r - 0.5 * (n - c);
    r = r + 0.5 * 0.0;
    r = r - 0.5 * logdet_H;
    r = r - 0.5 * logdet_WHW;
    r = r - 0.5 * (n - c) * log(yPy);
    return r;
}

/* wrapper_likelihood_derivative1_restricted_lambda (pygemma_model.pyx:1631-1649) */
static double wrapper_d1(pgo_ctx *c, double lam)
{
    pgo_eval e;
    precompute(c, lam, 0, &e, NULL);
    return pgo_d1(lam, c->n, c->c0 + 1, e.yPy, e.yPPy, e.trP);
}

/*
 * scipy.optimize.brentq (third-party: SciPy is unpinned in the reference's
 * requirements.txt:11; the build container has SciPy 1.18.1).  Restated from
 * SciPy's published C routine scipy/optimize/Zeros/brentq.c; call site
 * pygemma_model.pyx:176-182 (xtol = SciPy default 2e-12, rtol = 0.1,
 * maxiter = 100, disp=False).  tests/test_oracle.py checks call sequences and
 * roots against the installed scipy.optimize.brentq.
 */
typedef double (*pgo_fn)(void *, double);
double pgo_brentq(pgo_fn f, void *arg, double xa, double xb, double xtol, double rtol, int maxiter,
                  int *funcalls)
{
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    fpre = f(arg, xpre);
    fcur = f(arg, xcur);
    *funcalls = 2;
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return NAN; /* SIGNERR: unreachable from the scan */
    for (int i = 0; i < maxiter; ++i) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre); /* secant */
            } else {
                dpre = (fpre - fcur) / (xpre - xcur); /* inverse quadratic */
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            const double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(arg, xcur);
        (*funcalls)++;
    }
    return xcur; /* CONVERR with disp=False returns the last iterate */
}

static double wrapper_d1_cb(void *arg, double lam) { return wrapper_d1((pgo_ctx *)arg, lam); }

static double sgn(double v) { return (v > 0) - (v < 0) + (isnan(v) ? NAN : 0.0); } /* np.sign */

/* newton (pygemma_model.pyx:1349-1416), precompute=True branch */
static double newton(pgo_ctx *c, double lam, double lambda_min, double lambda_max)
{
    double lambda_root = lam, lambda_new, ratio, d1, d2, r_eps;
    int iteration = 0;
    const int n = c->n, cf = c->c0 + 1;
    for (;;) {
        pgo_eval e;
        precompute(c, lambda_root, 1, &e, NULL);
        d1 = pgo_d1(lambda_root, n, cf, e.yPy, e.yPPy, e.trP);
        d2 = pgo_d2(lambda_root, n, cf, e.yPy, e.yPPy, e.yPPPy, e.trP, e.trPP);
        ratio = d1 / d2;
        if (sgn(ratio) * sgn(d1) * sgn(d2) <= 0.0) break; /* :1392 (NaN compares false) */
        lambda_new = lambda_root - ratio;
        r_eps = fabs(lambda_new - lambda_root) / fabs(lambda_root);
        if (lambda_new < lambda_min) break; /* :1398-1400 clamped value is discarded */
        if (lambda_new > lambda_max) break; /* :1402-1404 */
        if (isnan(lambda_new) || isinf(lambda_new)) break; /* :1406 */
        lambda_root = lambda_new;
        if (r_eps < 1e-5 || iteration > 100) break; /* :1411 */
        iteration++;
    }
    return lambda_root;
}

/* calc_lambda_restricted (pygemma_model.pyx:64-194): grid branch :99-132, default branch :135-194 */
double pgo_calc_lambda(pgo_ctx *c, int grid)
{
    const int n = c->n, cf = c->c0 + 1;
    const double lo = pow(10.0, -5.0), hi = pow(10.0, 5.0);
    pgo_eval e;
    double best_lambda, best_likelihood, ll;

    precompute(c, lo, 0, &e, NULL);
    best_likelihood = pgo_loglik(n, cf, e.yPy, e.logdet_H, e.logdet_WHW);
    precompute(c, hi, 0, &e, NULL);
    ll = pgo_loglik(n, cf, e.yPy, e.logdet_H, e.logdet_WHW);
    if (best_likelihood < ll) { best_likelihood = ll; best_lambda = hi; }
    else best_lambda = lo;

    if (grid) {
        for (int idx = -5; idx < 5; ++idx) { /* np.arange(-5, 5, 1) :92 */
            const double lam = pow(10.0, (double)idx);
            precompute(c, lam, 0, &e, NULL);
            ll = pgo_loglik(n, cf, e.yPy, e.logdet_H, e.logdet_WHW);
            if (ll > best_likelihood) { best_likelihood = ll; best_lambda = lam; }
        }
        return best_lambda;
    }

    double f0 = 0.0, f1 = 0.0;
    for (int idx = -5; idx < 5; ++idx) {
        const double lambda0 = pow(10.0, (double)idx), lambda1 = pow(10.0, (double)idx + 1.0);
        if (idx == -5) f0 = wrapper_d1(c, lambda0); else f0 = f1; /* :161-167 */
        f1 = wrapper_d1(c, lambda1);
        if (copysign(1.0, f0) * copysign(1.0, f1) < 0) { /* :174 */
            int calls;
            double lam = pgo_brentq(wrapper_d1_cb, c, lambda0, lambda1, 2e-12, 0.1, 100, &calls);
            lam = newton(c, lam, lambda0, lambda1); /* :184 */
            precompute(c, lam, 0, &e, NULL);        /* :186 */
            ll = pgo_loglik(n, cf, e.yPy, e.logdet_H, e.logdet_WHW);
            if (ll > best_likelihood) { best_likelihood = ll; best_lambda = lam; }
        }
    }
    return best_lambda;
}

/* calc_beta_vg_ve_restricted_overload (pygemma_model.pyx:1514-1537) + F_wald (lmm/lmm.py:471) */
void pgo_wald(pgo_ctx *c, double lam, double *beta, double *se, double *tau, double *F)
{
    pgo_eval e;
    const int df = c->n - c->c0 - 1;
    precompute(c, lam, 0, &e, NULL);
    *beta = e.yPx / e.xPx;
    *se = sqrt(e.yPy) / (sqrt(cy_max(e.xPx, PGO_MIN_VAL)) * sqrt((double)df));
    *tau = df / e.yPy;
    *F = pow(*beta / *se, 2.0);
}

/*
 * calculate (lmm/lmm.py:461-495) over a block of SNPs.  xr is SNP-major: SNP g
 * occupies xr[g*n .. g*n+n).  p-values are left to the caller
 * (scipy.stats.f.sf, lmm/lmm.py:482).
 */
typedef struct {
    int n, c0, grid;
    long m;
    const double *d, *w0, *y, *xr;
    double *beta, *se, *tau, *lam, *F;
    int32_t *n_eval2, *n_eval3;
    atomic_long next;
} scan_job;

static void *scan_worker(void *arg)
{
    scan_job *j = (scan_job *)arg;
    for (;;) {
        const long g0 = atomic_fetch_add(&j->next, 4);
        if (g0 >= j->m) break;
        const long g1 = g0 + 4 < j->m ? g0 + 4 : j->m;
        for (long g = g0; g < g1; ++g) {
            pgo_ctx c = {j->n, j->c0, j->c0 + 2, j->d, j->w0, j->y, j->xr + (size_t)g * j->n, 0, 0};
            j->lam[g] = pgo_calc_lambda(&c, j->grid);
            pgo_wald(&c, j->lam[g], &j->beta[g], &j->se[g], &j->tau[g], &j->F[g]);
            if (j->n_eval2) j->n_eval2[g] = (int32_t)c.n_eval2;
            if (j->n_eval3) j->n_eval3[g] = (int32_t)c.n_eval3;
        }
    }
    return NULL;
}

int pgo_num_threads(void)
{
    const char *e = getenv("PGO_NUM_THREADS");
    long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1;
    if (t > 256) t = 256;
    return (int)t;
}

void pgo_scan(int n, int c0, long m, const double *d, const double *w0, const double *y, const double *xr,
              int grid, double *beta, double *se, double *tau, double *lam, double *F, int32_t *n_eval2,
              int32_t *n_eval3)
{
    scan_job j = {n, c0, grid, m, d, w0, y, xr, beta, se, tau, lam, F, n_eval2, n_eval3, 0};
    int nt = pgo_num_threads();
    if (nt > m) nt = m > 0 ? (int)m : 1;
    pthread_t th[256];
    for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, scan_worker, &j);
    scan_worker(&j);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
}

/* unit-level probe: one precompute_mat call; levels has (c0+2)*5 doubles, scal has 9 */
void pgo_precompute_probe(int n, int c0, const double *d, const double *w0, const double *y, const double *x,
                          double lam, int full, double *levels, double *scal)
{
    pgo_ctx c = {n, c0, c0 + 2, d, w0, y, x, 0, 0};
    pgo_eval e;
    precompute(&c, lam, full, &e, levels);
    scal[0] = e.yPy; scal[1] = e.yPPy; scal[2] = e.yPPPy; scal[3] = e.trP; scal[4] = e.trPP;
    scal[5] = e.logdet_H; scal[6] = e.logdet_WHW; scal[7] = e.xPx; scal[8] = e.yPx;
}

/* brentq probe on f(x) = p0 + p1*x + p2*x^2 + p3*x^3 + p4*exp(-x): returns root, call count and the x sequence */
typedef struct { double p[5]; double *xs; int nx, cap; } poly_arg;
static double poly_cb(void *a, double x)
{
    poly_arg *q = (poly_arg *)a;
    if (q->nx < q->cap) q->xs[q->nx] = x;
    q->nx++;
    return q->p[0] + x * (q->p[1] + x * (q->p[2] + x * q->p[3])) + q->p[4] * exp(-x);
}
double pgo_poly_eval(const double *p5, double x)
{
    return p5[0] + x * (p5[1] + x * (p5[2] + x * p5[3])) + p5[4] * exp(-x);
}
double pgo_brentq_probe(const double *p5, double a, double b, double xtol, double rtol, int maxiter, double *xs,
                        int cap, int *ncalls)
{
    poly_arg q;
    memcpy(q.p, p5, sizeof q.p);
    q.xs = xs; q.nx = 0; q.cap = cap;
    int calls;
    double r = pgo_brentq(poly_cb, &q, a, b, xtol, rtol, maxiter, &calls);
    *ncalls = calls;
    return r;
}


/* ------------------------------------------------------------------------------------------------
 * Likelihood-ratio scaffolding of the reference (commented out in its driver, lmm/lmm.py:137-141,:176-190,:278-300):
 * the LIVE functions those lines call, restated in their own dense form -- explicit (W^T H^-1 W)^-1 instead of the Pab
 * recursion -- so that this half of the oracle shares no arithmetic with the product's table / x-row machinery.
 *   compute_at_Pi_b        pygemma_model.pyx:2045-2095     a^T P b
 *   compute_at_Pi_Pi_b     pygemma_model.pyx:2120-2173     a^T P P b
 *   compute_at_Pi_Pi_Pi_b  pygemma_model.pyx:2208-2279     a^T P P P b
 *   likelihood_lambda / _derivative1_lambda / _derivative2_lambda   pyx:1542-1603
 *   lmm.calc_lambda        lmm/lmm.py:22-84  (bracket scan, scipy brentq rtol=0.1 maxiter=5000, scipy newton rtol=1e-5
 *                                             maxiter=10 with fprime, argmax of likelihood_lambda over the candidates)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int n, c;
    const double *d;
    const double **cols; /* c column pointers (fixed effects of this model) */
    const double *y;
    long evals;
} ml_ctx;

typedef struct {
    double yPy, yPPy, yPPPy, trH, trHH, logdetH;
} ml_eval_t;

/* in-place inverse of a c x c matrix (row-major), Gauss-Jordan with partial pivoting (np.linalg.inv is LU based) */
static int invert(long double *a, int c)
{
    long double *inv = (long double *)calloc((size_t)c * c, sizeof(long double));
    for (int i = 0; i < c; ++i) inv[i * c + i] = 1.0L;
    for (int col = 0; col < c; ++col) {
        int piv = col;
        for (int r = col + 1; r < c; ++r)
            if (fabsl(a[r * c + col]) > fabsl(a[piv * c + col])) piv = r;
        if (a[piv * c + col] == 0.0L) { free(inv); return -1; }
        if (piv != col)
            for (int j = 0; j < c; ++j) {
                long double t = a[col * c + j]; a[col * c + j] = a[piv * c + j]; a[piv * c + j] = t;
                t = inv[col * c + j]; inv[col * c + j] = inv[piv * c + j]; inv[piv * c + j] = t;
            }
        const long double p = a[col * c + col];
        for (int j = 0; j < c; ++j) { a[col * c + j] /= p; inv[col * c + j] /= p; }
        for (int r = 0; r < c; ++r) {
            if (r == col) continue;
            const long double f = a[r * c + col];
            if (f == 0.0L) continue;
            for (int j = 0; j < c; ++j) { a[r * c + j] -= f * a[col * c + j]; inv[r * c + j] -= f * inv[col * c + j]; }
        }
    }
    memcpy(a, inv, sizeof(long double) * (size_t)c * c);
    free(inv);
    return 0;
}

static void ml_evaluate(ml_ctx *m, double lam, ml_eval_t *o)
{
    const int n = m->n, c = m->c;
    m->evals++;
    long double *A = (long double *)calloc((size_t)c * c + 3 * (size_t)c + 1, sizeof(long double));
    long double *g = A + (size_t)c * c, *coef = g + c, *q = coef + c;
    long double yHy = 0, s1 = 0, s2 = 0, sl = 0;
    for (int l = 0; l < n; ++l) {
        const double t = lam * m->d[l] + 1.0, h = 1.0 / t;
        s1 += h; s2 += (long double)h * h; sl += logl((long double)t);
        yHy += (long double)m->y[l] * m->y[l] * h;
        for (int r = 0; r < c; ++r) {
            const long double wr = (long double)m->cols[r][l] * h;
            g[r] += wr * m->y[l];
            for (int s = 0; s <= r; ++s) A[r * c + s] += wr * m->cols[s][l];
        }
    }
    for (int r = 0; r < c; ++r)
        for (int s = r + 1; s < c; ++s) A[r * c + s] = A[s * c + r];
    if (c > 0 && invert(A, c) != 0) { o->yPy = o->yPPy = o->yPPPy = NAN; o->trH = (double)s1; o->trHH = (double)s2; o->logdetH = (double)sl; free(A); return; }
    long double gAg = 0;
    for (int r = 0; r < c; ++r) {
        long double v = 0;
        for (int s = 0; s < c; ++s) v += A[r * c + s] * g[s];
        coef[r] = v;          /* (W^T H^-1 W)^-1 W^T H^-1 y */
        gAg += g[r] * v;
    }
    /* P y = H^-1 (y - W coef)  (pyx:2162-2170) */
    long double pp = 0, ppp_a = 0;
    for (int l = 0; l < n; ++l) {
        const double h = 1.0 / (lam * m->d[l] + 1.0);
        long double res = m->y[l];
        for (int r = 0; r < c; ++r) res -= coef[r] * m->cols[r][l];
        const long double py = res * h;
        pp += py * py;                 /* a^T P P b   :2171 */
        ppp_a += py * py * h;          /* :2259 */
        for (int r = 0; r < c; ++r) q[r] += (long double)m->cols[r][l] * py * h; /* W^T H^-1 P y  :2262-2268 */
    }
    long double qAq = 0;
    for (int r = 0; r < c; ++r) {
        long double v = 0;
        for (int s = 0; s < c; ++s) v += A[r * c + s] * q[s];
        qAq += q[r] * v;
    }
    o->yPy = (double)(yHy - gAg);      /* :2093 */
    o->yPPy = (double)pp;
    o->yPPPy = (double)(ppp_a - qAq);  /* :2279 */
    o->trH = (double)s1; o->trHH = (double)s2; o->logdetH = (double)sl;
    free(A);
}

/* pyx:1566-1581 */
static double ml_d1(int n, double lam, const ml_eval_t *e)
{
    double r = -0.5 * ((n - e->trH) / lam);
    const double num = cy_max(e->yPPy, PGO_MIN_VAL), denom = cy_max(e->yPy, PGO_MIN_VAL);
    return r + (n / 2.0) * (1.0 - num / denom) / lam;
}
/* pyx:1586-1603 */
static double ml_d2(int n, double lam, const ml_eval_t *e)
{
    const double yPy = cy_max(e->yPy, PGO_MIN_VAL), yPPy = cy_max(e->yPPy, PGO_MIN_VAL);
    const double g2 = (yPy + cy_max(e->yPPPy, PGO_MIN_VAL) - 2 * yPPy) / (lam * lam);
    const double g1 = (yPy - yPPy) / lam;
    double r = 0.5 * (n + e->trHH - 2 * e->trH) / pow(lam, 2);
    return r - 0.5 * n * (2 * g2 - g1 * g1 / yPy) / yPy;
}
/* pyx:1542-1560 */
static double ml_ll(int n, const ml_eval_t *e)
{
    double r = (n / 2.0) * log(n / (2 * M_PI));
    r = r - n / 2.0;
    r = r - 0.5 * e->logdetH;
    return r - (n / 2.0) * log(cy_max(e->yPy, PGO_MIN_VAL));
}

static double ml_d1_cb(void *arg, double lam)
{
    ml_ctx *m = (ml_ctx *)arg;
    ml_eval_t e;
    ml_evaluate(m, lam, &e);
    return ml_d1(m->n, lam, &e);
}

/* scipy.optimize.newton, Newton-Raphson branch (fprime given, fprime2 None), tol = 1.48e-8 (default), rtol, maxiter,
 * disp=False: call site lmm/lmm.py:71-76 */
static double scipy_newton(ml_ctx *m, double x0, double rtol, int maxiter)
{
    const double tol = 1.48e-8;
    double p0 = x0, p = x0;
    for (int itr = 0; itr < maxiter; ++itr) {
        ml_eval_t e;
        ml_evaluate(m, p0, &e);
        const double fval = ml_d1(m->n, p0, &e);
        if (fval == 0) return p0;
        const double fder = ml_d2(m->n, p0, &e);
        if (fder == 0) return p0;
        p = p0 - fval / fder;
        /* np.isclose(p, p0, rtol=rtol, atol=tol) */
        if (isfinite(p) && isfinite(p0) ? fabs(p - p0) <= tol + rtol * fabs(p0) : p == p0) return p;
        p0 = p;
    }
    return p;
}

/* lmm.calc_lambda (lmm/lmm.py:22-84); returns lambda, *ll_out = likelihood_lambda there, *yPy_out = y^T P y there */
static double ml_calc_lambda(ml_ctx *m, double *ll_out, double *yPy_out)
{
    double roots[16], f0 = 0, f1 = 0;
    int nroots = 0;
    roots[nroots++] = pow(10.0, -5.0);
    roots[nroots++] = pow(10.0, 5.0);
    for (int idx = -5; idx < 5; ++idx) {
        const double lambda0 = pow(10.0, (double)idx), lambda1 = pow(10.0, (double)idx + 1.0);
        if (idx == -5) f0 = ml_d1_cb(m, lambda0); else f0 = f1; /* :53-58 */
        f1 = ml_d1_cb(m, lambda1);
        if (sgn(f0) * sgn(f1) < 0) { /* :63 */
            int calls;
            double lam = pgo_brentq(ml_d1_cb, m, lambda0, lambda1, 2e-12, 0.1, 5000, &calls);
            lam = scipy_newton(m, lam, 1e-5, 10);
            if (nroots < 16) roots[nroots++] = lam;
        }
    }
    /* roots[np.argmax(likelihood_list)]: first maximum; a NaN is a maximum */
    int best = 0;
    double best_ll = 0, best_yPy = 0;
    for (int i = 0; i < nroots; ++i) {
        ml_eval_t e;
        ml_evaluate(m, roots[i], &e);
        const double ll = ml_ll(m->n, &e);
        if (i == 0) { best_ll = ll; best_yPy = e.yPy; continue; }
        if (isnan(best_ll)) continue;
        if (isnan(ll) || ll > best_ll) { best = i; best_ll = ll; best_yPy = e.yPy; }
    }
    *ll_out = best_ll;
    if (yPy_out) *yPy_out = best_yPy;
    return roots[best];
}

/* null model [W0] -> null3 = {lambda_null, tau_null = n / yPy, l_null}; alternative [W0, x_g] for every SNP row of xr */
void pgo_ml_scan(int n, int c0, long m, const double *d, const double *w0, const double *y, const double *xr,
                 double *null3, double *lambda_ml, double *loglik_ml)
{
    const double **cols = (const double **)malloc(sizeof(double *) * (size_t)(c0 + 1));
    for (int j = 0; j < c0; ++j) cols[j] = w0 + (size_t)j * n;
    ml_ctx mc = {n, c0, d, cols, y, 0};
    if (null3) {
        double ll, yPy;
        null3[0] = ml_calc_lambda(&mc, &ll, &yPy);
        null3[1] = (double)n / yPy;
        null3[2] = ll;
    }
    mc.c = c0 + 1;
    for (long g = 0; g < m; ++g) {
        cols[c0] = xr + (size_t)g * n;
        double ll;
        lambda_ml[g] = ml_calc_lambda(&mc, &ll, NULL);
        loglik_ml[g] = ll;
    }
    free(cols);
}
