// reml_solve.cuh -- stage 2b of the scan: the per-SNP REML optimiser on the compressed moments (sm_100a).
//
// One warp owns one SNP.  Its moments Z[j][k] (compress.cuh) are all that is left of the genotype vector;
// every optimiser iteration -- bracket scan, Brent, Newton, final likelihood, grid mode -- is
//     x row   : sum_k Z[j][k] / (lambda d_k + 1)^p ,  p = 1,2(,3)     lanes over nodes, halving butterfly
//     tables  : [W0,y] block from the exact / Chebyshev lambda tables   (pg_eval.cuh)
//     Pab     : projection recursion over covariates, warp-parallel     (pg_eval.cuh)
//     solver  : SnpSolver state machine                                 (pg_math.cuh)
// and never touches HBM-resident genotype data again: Z (#nodes x (c0+2) doubles, ~10 KB at n = 10 000,
// c0 = 10) stays in L1/L2 across the ~17 evaluations.
//
// Replaces: reference lmm/lmm.py:461-495 (calculate) and pygemma_model.pyx:64-194, :880-1053, :1349-1416,
// :1514-1537, :1631-1698, :1813-1830.
#pragma once

#include <cuda_runtime.h>

#include "reml_kernels.cuh"

namespace pg {

struct SolveArgs {
    int n, c0, grid;
    long long m, row0;
    const double* nodes;  // [Kcp] node eigenvalues (padding: 0)
    int Kcp;
    const double* Z;      // [m][c0+2][Kcp]
    Tables tab;
    double* out[6];
    int* status;
    int* n_eval2;
    int* n_eval3;
    unsigned long long* counter;
};

// Sum V per-lane values over the warp with V + O(V/8) shuffles instead of 5 V: offsets 16, 8, 4 exchange
// halves of the array, offsets 2 and 1 finish.  Afterwards lane group g = lane >> 2 holds in v[i], i < V/8,
// the warp total of the original entry  ((g>>2)&1) V/2 + ((g>>1)&1) V/4 + (g&1) V/8 + i.
template <int V>
__device__ __forceinline__ void warp_reduce_halving(double (&v)[V], int lane)
{
    static_assert(V % 8 == 0, "V must be a multiple of 8");
    {
        constexpr int H = V / 2;
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        constexpr int H = V / 4;
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        constexpr int H = V / 8;
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
#pragma unroll
    for (int i = 0; i < V / 8; ++i) {
        v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
        v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
    }
}

// x row entries for moment rows [jb, jb+NC): row j < c0 -> x.w_j, j == c0 -> x.y, j == c0+1 -> x.x
template <int NC, bool FULL>
__device__ __forceinline__ void solve_xrow_pass(const SolveArgs& a, const double* __restrict__ Zs, double lam, int jb,
                                                double* A, double* B, double* C)
{
    constexpr int NP = FULL ? 3 : 2;
    constexpr int V = ((NP * NC + 7) / 8) * 8;
    const int lane = threadIdx.x & 31, Kcp = a.Kcp;
    double v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.0;
    const double* __restrict__ z = Zs + (size_t)jb * Kcp;
    for (int k = lane; k < Kcp; k += 32) {
        const double h = 1.0 / fma(lam, __ldg(a.nodes + k), 1.0);
        const double h2 = h * h, h3 = h2 * h;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const double zz = __ldg(z + (size_t)j * Kcp + k);
            v[j] = fma(h, zz, v[j]);
            v[NC + j] = fma(h2, zz, v[NC + j]);
            if (FULL) v[2 * NC + j] = fma(h3, zz, v[2 * NC + j]);
        }
    }
    warp_reduce_halving<V>(v, lane);
    if ((lane & 3) == 0) {
        const int g = lane >> 2, c0 = a.c0;
        const int base = ((g >> 2) & 1) * (V / 2) + ((g >> 1) & 1) * (V / 4) + (g & 1) * (V / 8);
#pragma unroll
        for (int i = 0; i < V / 8; ++i) {
            const int o = base + i;
            if (o < NP * NC) {
                const int p = o / NC, j = jb + (o - p * NC);
                const int dst = (j < c0) ? tri(c0, j) : ((j == c0) ? tri(c0 + 1, c0) : tri(c0, c0));
                double* M = (p == 0) ? A : ((p == 1) ? B : C);
                M[dst] = v[i];
            }
        }
    }
}

template <bool FULL>
__device__ __noinline__ void solve_xrow_all(const SolveArgs& a, const double* __restrict__ Zs, double lam, double* A,
                                            double* B, double* C)
{
    const int k1 = a.c0 + 2;
    int jb = 0;
    while (k1 - jb > kChunkCols) {
        solve_xrow_pass<kChunkCols, FULL>(a, Zs, lam, jb, A, B, C);
        jb += kChunkCols;
    }
#define PG_CASE(NCV)                                             \
    case NCV:                                                    \
        solve_xrow_pass<NCV, FULL>(a, Zs, lam, jb, A, B, C);     \
        break;
    switch (k1 - jb) {
        PG_CASE(1) PG_CASE(2) PG_CASE(3) PG_CASE(4) PG_CASE(5) PG_CASE(6)
        PG_CASE(7) PG_CASE(8) PG_CASE(9) PG_CASE(10) PG_CASE(11) PG_CASE(12)
    default: break;
    }
#undef PG_CASE
}

__device__ __forceinline__ void eval_snp_compressed(const SolveArgs& a, const double* __restrict__ Zs, double lam,
                                                    int fixed_t, int full, int need_ll, double* A, double* B, double* C,
                                                    EvalOut* e)
{
    Level0 l0;
    if (full) {
        assemble_w0y<true>(a.tab, lam, fixed_t, A, B, C, &l0);
        solve_xrow_all<true>(a, Zs, lam, A, B, C);
        __syncwarp();
        pab_recursion<true>(a.tab, A, B, C, l0, need_ll != 0, e);
    } else {
        assemble_w0y<false>(a.tab, lam, fixed_t, A, B, C, &l0);
        solve_xrow_all<false>(a, Zs, lam, A, B, C);
        __syncwarp();
        pab_recursion<false>(a.tab, A, B, C, l0, need_ll != 0, e);
    }
}

__global__ void __launch_bounds__(256, 2) reml_solve_kernel(SolveArgs a)
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
        const double* __restrict__ Zs = a.Z + (size_t)g * k * a.Kcp;
        SnpSolver s;
        s.init(a.n, a.c0, a.grid);
        while (s.pending()) {
            EvalOut e;
            eval_snp_compressed(a, Zs, s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll(), A, B, C, &e);
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

// single evaluation from compressed moments (one warp): unit probe
__global__ void probe_precompute_compressed_kernel(SolveArgs a, double lam, int fixed_t, int full, double* out9)
{
    extern __shared__ double smem[];
    const int k = a.c0 + 2, TT = k * (k + 1) / 2;
    double* A = smem;
    double* B = A + TT;
    double* C = B + TT;
    EvalOut e;
    eval_snp_compressed(a, a.Z, lam, fixed_t, full, 1, A, B, C, &e);
    if ((threadIdx.x & 31) == 0) {
        out9[0] = e.yPy; out9[1] = e.yPPy; out9[2] = e.yPPPy; out9[3] = e.trP; out9[4] = e.trPP;
        out9[5] = e.logdetH; out9[6] = e.logdetWHW; out9[7] = e.xPx; out9[8] = e.yPx;
    }
}

}  // namespace pg
