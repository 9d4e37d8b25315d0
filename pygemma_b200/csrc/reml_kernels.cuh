// reml_kernels.cuh -- device kernels of the per-SNP REML scan (sm_100a).
//
//   build_tables_kernel   SNP-independent [W0,y] Gram tables at the fixed lambdas and Chebyshev nodes
//   reml_scan_kernel      one warp owns one SNP: every optimiser iteration runs on-chip; the rotated
//                         genotype vector is the only per-SNP HBM read (n doubles, coalesced)
//   probe kernels         unit-level entry points for the parity tests
//
// Replaces the Cython code of pygemma_model.pyx:64-194 (lambda search), :880-1053 (Pab recursion),
// :1349-1416 (Newton), :1514-1537 (Wald), :1656-1698/:1813-1830 (derivatives / likelihood) and the
// Python loop lmm/lmm.py:461-495.
#pragma once

#include <cuda_runtime.h>

#include "pg_eval.cuh"

namespace pg {

constexpr int kChunkCols = 12;  // [W0,y] columns accumulated per pass over the genotype vector
constexpr int kTablePairs = 8;  // pairs per CTA in build_tables_kernel

struct ScanArgs {
    int n, c0, grid;
    long long m;          // SNPs in this block
    long long row0;       // global row of the block's first SNP
    const double* d;      // eigenvalues (clipped), n
    const double* wy;     // rotated [W0, y], column-major, column j at wy + j*ldw (zero-padded)
    long long ldw;
    const double* xr;     // rotated genotypes, SNP-major: SNP g at xr + g*ldx
    long long ldx;
    Tables tab;
    double* out[6];       // beta, se_beta, tau, lambda, F_wald, p_wald (global rows)
    int* status;
    int* n_eval2;
    int* n_eval3;
    unsigned long long* counter;  // dynamic work queue
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One pass over the rotated genotype vector: x^T H^-p [w_jb .. w_jb+NC-1] for p = 1,2(,3) and, when
// WITH_XX, x^T H^-p x.  Results land in the x row of the packed triangles (index c0 = x, c0+1 = y).
template <int NC, bool FULL, bool WITH_XX>
__device__ __forceinline__ void xrow_pass(const ScanArgs& a, const double* __restrict__ x, double lam, int jb,
                                          double* A, double* B, double* C)
{
    const int n = a.n, lane = threadIdx.x & 31;
    const double* __restrict__ d = a.d;
    const double* __restrict__ w = a.wy + (size_t)jb * a.ldw;
    const size_t ldw = (size_t)a.ldw;
    double a1[NC], a2[NC], a3[NC];
    double xx1 = 0.0, xx2 = 0.0, xx3 = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j) { a1[j] = 0.0; a2[j] = 0.0; a3[j] = 0.0; }
#pragma unroll 2
    for (int l = lane; l < n; l += 32) {
        const double xv = x[l];
        const double h = 1.0 / fma(lam, d[l], 1.0);
        const double xh = xv * h, xh2 = xh * h;
        const double xh3 = FULL ? xh2 * h : 0.0;
        if (WITH_XX) {
            xx1 = fma(xh, xv, xx1);
            xx2 = fma(xh2, xv, xx2);
            if (FULL) xx3 = fma(xh3, xv, xx3);
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const double wv = w[(size_t)j * ldw + l];
            a1[j] = fma(xh, wv, a1[j]);
            a2[j] = fma(xh2, wv, a2[j]);
            if (FULL) a3[j] = fma(xh3, wv, a3[j]);
        }
    }
    const int c0 = a.c0;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const double s1 = warp_sum(a1[j]), s2 = warp_sum(a2[j]);
        const double s3 = FULL ? warp_sum(a3[j]) : 0.0;
        if (lane == 0) {
            const int col = jb + j;  // reduced column: < c0 -> W0 column, == c0 -> y
            const int dst = (col < c0) ? tri(c0, col) : tri(c0 + 1, c0);
            A[dst] = s1; B[dst] = s2;
            if (FULL) C[dst] = s3;
        }
    }
    if (WITH_XX) {
        const double s1 = warp_sum(xx1), s2 = warp_sum(xx2);
        const double s3 = FULL ? warp_sum(xx3) : 0.0;
        if (lane == 0) {
            const int dst = tri(c0, c0);
            A[dst] = s1; B[dst] = s2;
            if (FULL) C[dst] = s3;
        }
    }
}

template <bool FULL>
__device__ __noinline__ void xrow_all(const ScanArgs& a, const double* __restrict__ x, double lam, double* A,
                                      double* B, double* C)
{
    const int k0 = a.c0 + 1;
    int jb = 0;
    bool first = true;
    while (k0 - jb > kChunkCols) {
        if (first) xrow_pass<kChunkCols, FULL, true>(a, x, lam, jb, A, B, C);
        else xrow_pass<kChunkCols, FULL, false>(a, x, lam, jb, A, B, C);
        first = false;
        jb += kChunkCols;
    }
    const int rem = k0 - jb;
#define PG_CASE(NCV)                                                             \
    case NCV:                                                                    \
        if (first) xrow_pass<NCV, FULL, true>(a, x, lam, jb, A, B, C);           \
        else xrow_pass<NCV, FULL, false>(a, x, lam, jb, A, B, C);                \
        break;
    switch (rem) {
        PG_CASE(1) PG_CASE(2) PG_CASE(3) PG_CASE(4) PG_CASE(5) PG_CASE(6)
        PG_CASE(7) PG_CASE(8) PG_CASE(9) PG_CASE(10) PG_CASE(11) PG_CASE(12)
    default: break;
    }
#undef PG_CASE
}

// one precompute_mat-equivalent evaluation for the SNP whose rotated genotype vector is x (warp-collective)
__device__ __forceinline__ void eval_snp(const ScanArgs& a, const double* __restrict__ x, double lam, int fixed_t,
                                         int full, int need_ll, double* A, double* B, double* C, EvalOut* e)
{
    Level0 l0;
    if (full) {
        assemble_w0y<true>(a.tab, lam, fixed_t, A, B, C, &l0);
        xrow_all<true>(a, x, lam, A, B, C);
        __syncwarp();
        pab_recursion<true>(a.tab, A, B, C, l0, need_ll != 0, e);
    } else {
        assemble_w0y<false>(a.tab, lam, fixed_t, A, B, C, &l0);
        xrow_all<false>(a, x, lam, A, B, C);
        __syncwarp();
        pab_recursion<false>(a.tab, A, B, C, l0, need_ll != 0, e);
    }
}

__global__ void __launch_bounds__(256) reml_scan_kernel(ScanArgs a)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.c0 + 2, TT = k * (k + 1) / 2;
    double* A = smem + (size_t)warp * 3 * TT;
    double* B = A + TT;
    double* C = B + TT;
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(a.counter, 1ULL);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= (unsigned long long)a.m) break;
        const double* __restrict__ x = a.xr + (size_t)g * a.ldx;
        SnpSolver s;
        s.init(a.n, a.c0, a.grid);
        while (s.pending()) {
            EvalOut e;
            eval_snp(a, x, s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll(), A, B, C, &e);
            s.feed(e);
        }
        if (lane == 0) {
            const long long row = a.row0 + (long long)g;
            a.out[0][row] = s.beta; a.out[1][row] = s.se; a.out[2][row] = s.tau;
            a.out[3][row] = s.lambda; a.out[4][row] = s.F; a.out[5][row] = s.p;
            if (a.status) a.status[row] = s.status;
            if (a.n_eval2) a.n_eval2[row] = s.n_eval2;
            if (a.n_eval3) a.n_eval3[row] = s.n_eval3;
        }
    }
}

// single evaluation probe (one warp)
__global__ void probe_precompute_kernel(ScanArgs a, double lam, int fixed_t, int full, double* out9)
{
    extern __shared__ double smem[];
    const int k = a.c0 + 2, TT = k * (k + 1) / 2;
    double* A = smem;
    double* B = A + TT;
    double* C = B + TT;
    EvalOut e;
    eval_snp(a, a.xr, lam, fixed_t, full, 1, A, B, C, &e);
    if ((threadIdx.x & 31) == 0) {
        out9[0] = e.yPy; out9[1] = e.yPPy; out9[2] = e.yPPPy; out9[3] = e.trP; out9[4] = e.trPP;
        out9[5] = e.logdetH; out9[6] = e.logdetWHW; out9[7] = e.xPx; out9[8] = e.yPx;
    }
}

__global__ void probe_f_sf_kernel(const double* F, double nu, long long k, double* p)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) p[i] = f_sf_1(F[i], nu);
}

// ------------------------------------------------------------------------------------------------
// Table builder: grid = (pair chunks + 1, table rows).  Each CTA reduces kTablePairs [W0,y] pairs
// (three powers each) or, for the last chunk, the three scalar functions, over all n eigenvalues.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_tables_kernel(int n, int c0, const double* __restrict__ d,
                                                            const double* __restrict__ wy, long long ldw,
                                                            const double* __restrict__ lambdas,
                                                            double* fixtab, double* itab, const TriAB* tri_ab)
{
    const int k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2, NF = 3 * T0 + 3;
    const int row = blockIdx.y;
    const double lam = lambdas[row];
    double* dst = (row < kNumFixed) ? fixtab + (size_t)row * NF : itab + (size_t)(row - kNumFixed) * NF;
    const int nchunks = (T0 + kTablePairs - 1) / kTablePairs;
    const int chunk = blockIdx.x;
    __shared__ double red[8][3 * kTablePairs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[3 * kTablePairs];
#pragma unroll
    for (int i = 0; i < 3 * kTablePairs; ++i) acc[i] = 0.0;
    if (chunk < nchunks) {
        const int q0 = chunk * kTablePairs;
        int ra[kTablePairs], sa[kTablePairs];
#pragma unroll
        for (int i = 0; i < kTablePairs; ++i) {
            const int q = min(q0 + i, T0 - 1);
            ra[i] = tri_ab[q].a; sa[i] = tri_ab[q].b;
        }
        for (int l = threadIdx.x; l < n; l += blockDim.x) {
            const double h = 1.0 / fma(lam, d[l], 1.0);
            const double h2 = h * h, h3 = h2 * h;
#pragma unroll
            for (int i = 0; i < kTablePairs; ++i) {
                const double p = wy[(size_t)ra[i] * ldw + l] * wy[(size_t)sa[i] * ldw + l];
                acc[3 * i] = fma(p, h, acc[3 * i]);
                acc[3 * i + 1] = fma(p, h2, acc[3 * i + 1]);
                acc[3 * i + 2] = fma(p, h3, acc[3 * i + 2]);
            }
        }
    } else {
        for (int l = threadIdx.x; l < n; l += blockDim.x) {
            const double t = fma(lam, d[l], 1.0);
            const double h = 1.0 / t;
            acc[0] += h;
            acc[1] = fma(h, h, acc[1]);
            acc[2] += log(t);
        }
    }
#pragma unroll
    for (int i = 0; i < 3 * kTablePairs; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3 * kTablePairs) {
        double v = 0.0;
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) v += red[w][threadIdx.x];
        const int i = threadIdx.x / 3, p = threadIdx.x % 3;
        if (chunk < nchunks) {
            const int q = chunk * kTablePairs + i;
            if (q < T0) dst[p * T0 + q] = v;
        } else if (threadIdx.x < 3) {
            dst[3 * T0 + threadIdx.x] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same tables as ONE dense contraction per design (round 2): T (rows x powers, pairs) = Hpow^T (rows x powers, n) . P (n, pairs)
// with Hpow[l][3 r + p] = 1 / (lambda_r d_l + 1)^(p+1), which depends on the eigen-system and the table lambdas only and is
// kept per handle, and P[l][q] = wy_a[l] wy_b[l] per design.  3.8 GFLOP at n = 10 000, c0 = 10 on the FP64 tensor pipe
// (cuBLAS DGEMM, setup path) instead of 1.2 ms of per-row reductions; the three scalar functions per row (sum h, sum h^2,
// sum log(lambda d + 1)) do not depend on the design at all and are reduced once per eigen-system.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_hpow_kernel(int n, long long ldh, const double* __restrict__ d,
                                                          const double* __restrict__ lambdas, double* __restrict__ Hpow)
{
    const int row = blockIdx.y;
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= ldh) return;
    double h = 0.0, h2 = 0.0, h3 = 0.0;
    if (l < n) {
        h = 1.0 / fma(lambdas[row], d[l], 1.0);   // the same arithmetic as build_tables_kernel
        h2 = h * h; h3 = h2 * h;
    }
    double* col = Hpow + (size_t)(3 * row) * ldh + l;
    col[0] = h; col[ldh] = h2; col[2 * ldh] = h3;
}

// hscal[row][0..2] = sum_l h, sum_l h^2, sum_l log(lambda d_l + 1): one CTA per table row
__global__ void __launch_bounds__(256) build_hscal_kernel(int n, const double* __restrict__ d, const double* __restrict__ lambdas,
                                                           double* __restrict__ hscal)
{
    const int row = blockIdx.x;
    const double lam = lambdas[row];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int l = threadIdx.x; l < n; l += blockDim.x) {
        const double t = fma(lam, d[l], 1.0);
        const double h = 1.0 / t;
        a0 += h;
        a1 = fma(h, h, a1);
        a2 += log(t);
    }
    __shared__ double red[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        hscal[(size_t)row * 3 + threadIdx.x] = v;
    }
}

// P[l + q * ldh] = wy_a[l] wy_b[l] for the T0 pairs (a >= b) of the [W0, y] columns
__global__ void __launch_bounds__(256) pair_products_kernel(int n, long long ldh, int T0, const double* __restrict__ wy,
                                                             long long ldw, const TriAB* __restrict__ tri_ab, double* __restrict__ P)
{
    const int q = blockIdx.y;
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= ldh || q >= T0) return;
    const TriAB ab = tri_ab[q];
    P[(size_t)q * ldh + l] = l < n ? wy[(size_t)ab.a * ldw + l] * wy[(size_t)ab.b * ldw + l] : 0.0;
}

// C[(3 row + p) + q * R3] -> table rows [row][p * T0 + q], plus the three scalars per row
__global__ void scatter_tables_kernel(int T0, int NF, int R3, const double* __restrict__ C, const double* __restrict__ hscal,
                                      double* __restrict__ fixtab, double* __restrict__ itab)
{
    const int row = blockIdx.x;
    double* dst = (row < kNumFixed) ? fixtab + (size_t)row * NF : itab + (size_t)(row - kNumFixed) * NF;
    for (int f = threadIdx.x; f < NF; f += blockDim.x) {
        if (f < 3 * T0) {
            const int p = f / T0, q = f - p * T0;
            dst[f] = C[(size_t)(3 * row + p) + (size_t)q * R3];
        } else {
            dst[f] = hscal[(size_t)row * 3 + (f - 3 * T0)];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Genotype staging: any dtype / layout -> fp64 SNP-major block (SNP g at dst + g*n), the layout both the
// rotation GEMM (as its column-major B operand) and the REML kernel read.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void to_snp_major_kernel(const T* __restrict__ src, long long ld, int layout, int n, long long mb,
                                    double* __restrict__ dst, long long ldd, const int* __restrict__ perm)
{
    // perm (nullable): destination sample index j reads source sample perm[j] (pre-rotated inputs arriving in the
    // caller's eigen order are brought into the handle's ascending-eigenvalue order)
    __shared__ double tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    const long long g0 = (long long)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    if (layout == 0) {
        // sample-major: rows j (samples), contiguous over g
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int j = j0 + r;
            const long long g = g0 + tx;
            tile[r][tx] = (j < n && g < mb) ? (double)src[(size_t)(perm ? perm[j] : j) * ld + g] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const long long g = g0 + r;
            const int j = j0 + tx;
            if (g < mb && j < n) dst[(size_t)g * ldd + j] = tile[tx][r];
        }
    } else {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const long long g = g0 + r;
            const int j = j0 + tx;
            if (g < mb && j < n) dst[(size_t)g * ldd + j] = (double)src[(size_t)g * ld + (perm ? perm[j] : j)];
        }
    }
}

__global__ void clip_nonneg_kernel(double* d, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = fmax(0.0, d[i]);  // lmm/lmm.py:157 (np.maximum(0.0, eigenVals))
}

}  // namespace pg
