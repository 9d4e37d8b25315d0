// reml_solve.cuh -- stage 2b of the scan: the per-SNP REML optimiser on the compressed moments (sm_100a).
//
// One warp owns one SNP.  Its moments Z[j][k] (compress.cuh) are all that is left of the genotype vector;
// every optimiser iteration -- bracket scan, Brent, Newton, final likelihood, grid mode -- is
//     x row   : sum_k Z[j][k] / (lambda d_k + 1)^p ,  p = 1,2(,3)     lanes over nodes, halving butterfly
//     tables  : covariate levels of the Pab recursion, eliminated once per table lambda ("table-2" rows,
//               exact at the 11 fixed lambdas, Chebyshev-interpolated elsewhere)    (pg_eval.cuh)
//     Pab     : the x row carried through the c0 covariate levels in registers (lane j owns column j, pivots
//               by shuffle, no division), then the SNP's own pivot                  (pg_eval.cuh)
//     solver  : SnpSolver state machine                                             (pg_math.cuh)
// and never touches HBM-resident genotype data again: Z (#nodes x (c0+2) doubles, ~10 KB at n = 10 000,
// c0 = 10) stays in L1/L2 across the ~17 evaluations.
//
// Replaces: reference lmm/lmm.py:461-495 (calculate) and pygemma_model.pyx:64-194, :880-1053, :1349-1416,
// :1514-1537, :1631-1698, :1813-1830.
#pragma once

#include <cuda_runtime.h>

#include "pg_debug.cuh"
#include "reml_kernels.cuh"

namespace pg {

struct SolveArgs {
    int n, c0, grid;
    long long m, row0;
    const double* nodes;  // [Kcp] node eigenvalues (padding: 0)
    int Kcp;
    int Kc;               // nodes really in use (<= Kcp): the padding nodes carry zero moments and are not read by the solver
    const double* Z;      // [m][zrows][Kcp], slab rows as laid out in compress.cuh
    int k1p;              // c0+2 rounded up to a multiple of 4 (padding rows are zero)
    int zrows;            // rows per SNP slab: k1p - 1 + number of phenotypes
    int yrow;             // slab row of this launch's phenotype (k1p - 1 + phenotype index)
    const double* FX;     // [kNumFixed][kFxOut][ldF] results of the fixed-lambda evaluations (fixed_phase_kernel), or nullptr
    long long ldF;        // SNP stride of FX
    int zsm;              // 1: every warp stages its SNP's slab (k1p rows x Kcp) in shared memory once and the ~7 SNP-specific
                          //    evaluations read it from there (ncu r02: 28 % of the solver's stall samples sat on these loads)
    int swap;             // 1: "de" mode, the genotype column is the phenotype and y the tested regressor (pg_eval.cuh)
    // likelihood-ratio outputs (nullable, all four or none): ML lambda and log-likelihood of the model [W0, x],
    // D_lrt = 2 (l_alt - l_null), p_lrt = chi2(1) upper tail (reference lmm/lmm.py:278-282,:300, commented there)
    double* lrt[4];
    double l_null;        // ML log-likelihood of the null model [W0] of this launch's phenotype (null_model_kernel)
    Tables2 t2;
    double* out[6];
    int* status;
    int* n_eval2;
    int* n_eval3;
    unsigned long long* counter;
};

// 1/t for t = lambda d + 1 in [1, 1e10+]: hardware seed (2^-20) + two Newton steps, within 1 ulp; t is always a normal
// number >= 1 here, so no special cases are needed (a full IEEE division costs ~6x the instructions)
__device__ __forceinline__ double rcp_ge1(double t)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    double e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---- x rows of the fixed-lambda evaluations ------------------------------------------------------------------
// The first kNumFixed evaluations of every SNP (the bracket scan, all of grid mode) sit at lambda_t = 10^(t-5) whatever
// the optimiser decides later, so their level-0 x rows  sum_k Z[row][k] / (lambda_t d_k + 1)^p,  p = 1, 2,  are one
// dense contraction per block:  F (slab rows x 22)  =  Z (slab rows x nodes)  .  H (nodes x 22)  on the FP64 tensor
// pipe (DMMA m8n8k4), instead of kNumFixed x-row passes + butterfly reductions per SNP inside the solver.
constexpr int kFxCols = 24;   // kNumFixed lambdas x 2 powers = 22 columns, padded to three 8-column DMMA tiles

// H[k][2 t + p] = (lambda_t d_k + 1)^-(p+1)   (the same rcp_ge1 arithmetic as solve_xrow_pass)
__global__ void build_h_kernel(const double* __restrict__ nodes, int Kcp, double* __restrict__ H)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kcp * kFxCols) return;
    const int k = i / kFxCols, c = i - k * kFxCols;
    double v = 0.0;
    if (c < 2 * kNumFixed) {
        const double h = rcp_ge1(fma(fixed_lambda(c >> 1), nodes[k], 1.0));
        v = (c & 1) ? h * h : h;
    }
    H[i] = v;
}

// Output layout: F2[(t * zrows + j) * ldF + snp] = (power 1, power 2) of slab row j of the SNP at lambda_t -- SNP-fastest,
// so that fixed_phase_kernel (one THREAD per SNP) reads it coalesced.
// One warp contracts 32 slab rows (4 m-tiles) against all 24 columns; Z is read once.  H passes through shared memory in
// slabs of `kslab` nodes (a multiple of 4; the whole table when it fits): a wide spectrum can need more nodes than one
// CTA's shared memory holds (Kcp * 192 bytes), the accumulators simply stay in registers across slabs.
constexpr int kFxSlabMax = 960;   // nodes per shared-memory slab: 960 * 24 * 8 = 180 KB
__global__ void __launch_bounds__(256)
fixed_xrow_kernel(const double* __restrict__ Z, long long rows, int Kcp, const double* __restrict__ H, double2* __restrict__ F2,
                  int kslab, int zrows, long long ldF)
{
    extern __shared__ double Hs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r0 = ((long long)blockIdx.x * 8 + warp) * 32;
    const bool active = r0 < rows;   // inactive warps still take part in the slab barriers
    const int ar = lane >> 2, ac = lane & 3;
    double acc[4][3][2];
    const double* zp[4];
    bool ok[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const long long row = r0 + mt * 8 + ar;
        ok[mt] = active && row < rows;
        zp[mt] = Z + (size_t)(ok[mt] ? row : 0) * Kcp + ac;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    }
    for (int k0 = 0; k0 < Kcp; k0 += kslab) {
        const int kn = min(kslab, Kcp - k0);
        if (k0) __syncthreads();   // everybody is done with the previous slab
        for (int i = threadIdx.x; i < kn * kFxCols; i += 256) Hs[i] = H[(size_t)k0 * kFxCols + i];
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int ks = 0; ks < (kn >> 2); ++ks) {
            double b[3];
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) b[nt] = Hs[(ks * 4 + ac) * kFxCols + nt * 8 + ar];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                PG_BOUNDS(k0 + ks * 4 + ac < Kcp && (ks * 4 + ac) < kslab, "fixed x rows: node index");
                const double av = ok[mt] ? __ldg(zp[mt] + k0 + ks * 4) : 0.0;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                                 : "d"(av), "d"(b[nt]));
                }
            }
        }
    }
    if (!active) return;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        if (!ok[mt]) continue;
        const long long row = r0 + mt * 8 + ar;
        const long long snp = row / zrows;
        const int j = (int)(row - snp * zrows);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            const int t = nt * 4 + ac;   // accumulator columns 2t, 2t+1 = the two powers at lambda_t
            if (t < kNumFixed) F2[((size_t)t * zrows + j) * ldF + snp] = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
    }
}

// ---- the fixed-lambda evaluations, one THREAD per SNP -----------------------------------------------------------
// The first kNumFixed evaluations of every SNP (bracket scan; all of grid mode; the same for the ML optimiser of the LRT
// outputs) share their lambda across SNPs, so the covariate levels of the Pab recursion read the SAME table-2 row for
// every SNP.  Here a CTA stages that row in shared memory once per lambda and each thread carries its own SNP's x row
// (shared memory, SNP-fastest: conflict-free) through the c0 levels: no shuffles, no dependent global loads, all 32
// lanes busy -- against 12 of 32 lanes, two shuffles and two L2 loads per level in the warp-per-SNP form
// (xrow_recursion_warp), which remains for the SNP-specific lambdas of Brent / Newton.  Arithmetic per entry is the
// expression of pab_recursion (pg_eval.cuh).  Output FX[(t * kFxOut + f) * ldF + snp], f = yPy, yPPy, trP, logdetWHW,
// xPx, yPx (level c_f / c0 as in EvalOut).
constexpr int kFxOut = 6;

struct FixedArgs {
    int c0, zrows, yrow, swap;
    long long m, ldF;
    const double2* F2;
    Tables2 t2;
    double* FX;
};

__global__ void __launch_bounds__(128) fixed_phase_kernel(FixedArgs a)
{
    extern __shared__ double fsm[];
    const int c0 = a.c0, k1 = c0 + 2, dg = c0 + 1, NF2 = a.t2.NF2, Tp = a.t2.Tp;
    const int NF2p = (NF2 + 1) & ~1;
    double* row2 = fsm;                 // the table-2 row of lambda_t
    double* xa = fsm + NF2p;            // [k1][128]
    double* xb = xa + (size_t)k1 * 128;
    const int tid = threadIdx.x;
    const long long snp = (long long)blockIdx.x * 128 + tid;
    const bool ok = snp < a.m;
    const bool swap = a.swap != 0;
    for (int t = 0; t < kNumFixed; ++t) {
        __syncthreads();   // everybody is done with the previous lambda's row
        for (int i = tid; i < NF2; i += 128) row2[i] = __ldg(a.t2.fix2 + (size_t)t * NF2 + i);
        for (int j = 0; j < k1; ++j) {
            const int r = j < c0 ? j : (j == c0 ? a.yrow : c0);   // Pab column order [W0, y, x] -> slab rows
            const double2 v = ok ? __ldg(a.F2 + ((size_t)t * a.zrows + r) * a.ldF + snp) : make_double2(0.0, 0.0);
            xa[j * 128 + tid] = v.x;
            xb[j * 128 + tid] = v.y;
        }
        __syncthreads();
        if (c0 == 0 && !swap) xa[dg * 128 + tid] = cy_max(xa[dg * 128 + tid], kMinVal);   // pyx:939 / :993
        for (int p = 0; p < c0; ++p) {
            const double al2 = row2[3 * p], al4 = row2[3 * p + 1];
            const double ar = xa[p * 128 + tid], br = xb[p * 128 + tid];
            const int cb = t2_col(c0, p, p + 1) - (p + 1);
            for (int j = p + 1; j <= dg; ++j) {
                double as = ar, bs = br;
                if (j <= c0) { as = row2[cb + j]; bs = row2[Tp + cb + j]; }
                const bool clamp = (p == c0 - 1) && (j == dg) && !swap;
                double v = (xb[j * 128 + tid] + al4 * ar * as) + al2 * (ar * bs + br * as);
                if (clamp) v = cy_max(v, kMinVal);
                xb[j * 128 + tid] = v;
                v = xa[j * 128 + tid] + al2 * ar * as;
                if (clamp) v = cy_max(v, kMinVal);
                xa[j * 128 + tid] = v;
            }
        }
        EvalOut e;
        const double* fin = row2 + t2_fin(c0);
        if (swap)
            xrow_final_level_swapped<false>(c0, fin, xa[dg * 128 + tid], xb[dg * 128 + tid], 0.0, xa[c0 * 128 + tid],
                                            xb[c0 * 128 + tid], 0.0, true, &e);
        else
            xrow_final_level<false>(fin, xa[dg * 128 + tid], xb[dg * 128 + tid], 0.0, xa[c0 * 128 + tid], xb[c0 * 128 + tid],
                                    0.0, true, &e);
        if (ok) {
            double* fx = a.FX + (size_t)t * kFxOut * a.ldF + snp;
            fx[0] = e.yPy; fx[a.ldF] = e.yPPy; fx[2 * a.ldF] = e.trP; fx[3 * a.ldF] = e.logdetWHW;
            fx[4 * a.ldF] = e.xPx; fx[5 * a.ldF] = e.yPx;
        }
    }
}

// Sum V per-lane values over the tile with V + O(V/8) shuffles instead of 5 V: offsets 16, 8, 4 exchange
// halves of the array, offsets 2 and 1 finish.  Afterwards lane group g = lane >> 2 holds in v[i], i < V/8,
// the warp total of the original entry  ((g>>2)&1) V/2 + ((g>>1)&1) V/4 + (g&1) V/8 + i   (T = 16: g = (lane & 15) >> 2,
// i < V/4, entry ((g>>1)&1) V/2 + (g&1) V/4 + i).
// T lanes own one SNP: T = 32 (a warp) or 16 (two SNPs per warp, for x rows of at most 16 entries: the covariate recursion
// and the optimiser's scalar code keep at most 16 lanes busy, so a half warp per SNP doubles their throughput).  Every
// collective below names only the lanes of its own tile, the two tiles of a warp are free to diverge.
template <int T>
__device__ __forceinline__ unsigned tile_mask()
{
    return T == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
}

template <int V, int T>
__device__ __forceinline__ void warp_reduce_halving(double (&v)[V], int lane)
{
    static_assert(V % 8 == 0, "V must be a multiple of 8");
    const unsigned tm = tile_mask<T>();
    constexpr int S0 = (T == 32) ? 2 : 1;   // the array half exchanged with offset 16 exists only for a whole warp
    if (T == 32) {
        constexpr int H = V / 2;
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(tm, send, 16);
        }
    }
    {
        constexpr int H = V / (2 * S0);
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(tm, send, 8);
        }
    }
    {
        constexpr int H = V / (4 * S0);
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(tm, send, 4);
        }
    }
#pragma unroll
    for (int i = 0; i < V / (4 * S0); ++i) {
        v[i] += __shfl_xor_sync(tm, v[i], 2);
        v[i] += __shfl_xor_sync(tm, v[i], 1);
    }
}

// level-0 x row for slab rows [jb, jb+NC) -> xs[p * k1p + j] (p = power index; row j < c0: x.w_j, c0: x.x, k1p-1: x.y,
// the last one read from the launch's own phenotype row a.yrow)
// ZSM: Zs is the warp's shared-memory copy of the slab (rows 0..k1p-2, then the phenotype row at k1p-1), else the slab
// in global memory (phenotype row a.yrow, read-only cache path).
template <int NC, bool FULL, bool ZSM, int T>
__device__ __forceinline__ void solve_xrow_pass(const SolveArgs& a, const double* __restrict__ Zs, double lam, int jb,
                                                double* xs)
{
    constexpr int NP = FULL ? 3 : 2;
    constexpr int V = ((NP * NC + 7) / 8) * 8;
    const int lane = threadIdx.x & (T - 1), Kcp = a.Kcp;
    double v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.0;
    const double* __restrict__ z = Zs + (size_t)jb * Kcp;
    const double* __restrict__ zl = Zs + (size_t)(jb + NC == a.k1p ? (ZSM ? a.k1p - 1 : a.yrow) : jb + NC - 1) * Kcp;
    // Global-memory slabs: the padding nodes (>= Kc) carry zero moments and are not fetched (adding 0 is exact); see
    // solve_xrow_all for the padding rows.
    const int kend = ZSM ? Kcp : a.Kc;
#pragma unroll 2
    for (int k = lane; k < kend; k += T) {
        const double h = rcp_ge1(fma(lam, __ldg(a.nodes + k), 1.0));
        const double h2 = h * h, h3 = h2 * h;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const double zz = ZSM ? ((j == NC - 1) ? zl[k] : z[(size_t)j * Kcp + k])
                                  : ((j == NC - 1) ? __ldg(zl + k) : __ldg(z + (size_t)j * Kcp + k));
            v[j] = fma(h, zz, v[j]);
            v[NC + j] = fma(h2, zz, v[NC + j]);
            if (FULL) v[2 * NC + j] = fma(h3, zz, v[2 * NC + j]);
        }
    }
    warp_reduce_halving<V, T>(v, lane);
    if ((lane & 3) == 0) {
        const int g = lane >> 2;
        const int base = T == 32 ? ((g >> 2) & 1) * (V / 2) + ((g >> 1) & 1) * (V / 4) + (g & 1) * (V / 8)
                                 : ((g >> 1) & 1) * (V / 2) + (g & 1) * (V / 4);
#pragma unroll
        for (int i = 0; i < (T == 32 ? V / 8 : V / 4); ++i) {
            const int o = base + i;
            if (o < NP * NC) {
                const int p = o / NC, j = jb + (o - p * NC);
                PG_BOUNDS(p < 3 && j < a.k1p, "solver: x-row scratch");
                xs[p * a.k1p + j] = v[i];
            }
        }
    }
}

template <bool FULL, bool ZSM, int T>
__device__ __noinline__ void solve_xrow_all(const SolveArgs& a, const double* __restrict__ Zs, double lam, double* xs)
{
    int jb = 0;
    // rows per pass: 12 with two powers (24 accumulators per lane), 8 with three (24 again): the 36 accumulators of a
    // 12-row three-power pass spill at 128 registers (measured: solve stage -7..11 % with the narrower pass)
    constexpr int kBig = FULL ? 8 : 12;
    while (a.k1p - jb >= kBig) {
        solve_xrow_pass<kBig, FULL, ZSM, T>(a, Zs, lam, jb, xs);
        jb += kBig;
    }
    if (a.k1p - jb == 8) solve_xrow_pass<8, FULL, ZSM, T>(a, Zs, lam, jb, xs);
    else if (a.k1p - jb == 4) {
        // rows c0+1 .. k1p-2 are padding (all zero): when only the phenotype row is left (c0 = 11: rows 12-14 padding, 15 = y)
        // a one-row pass reads 13 instead of 16 rows per evaluation -- with the node bound above the slabs of the SNPs in
        // flight (re-read from the L2 by every evaluation) shrink from 95 MB to 71 MB at the mouse shape (Kc = 129)
        if (!ZSM && a.c0 < jb) solve_xrow_pass<1, FULL, ZSM, T>(a, Zs, lam, a.k1p - 1, xs);
        else solve_xrow_pass<4, FULL, ZSM, T>(a, Zs, lam, jb, xs);
    }
}

// The covariate levels applied to the x row held in registers: lane (j & 31), slot (j >> 5) owns entry j.
// Pivot values travel by shuffle, pivot columns and level scalars come from the table-2 row (pg_eval.cuh).
template <bool FULL, int NS, int T>
__device__ __forceinline__ void xrow_recursion_warp(int c0, const double* __restrict__ row2, double (&xa)[NS],
                                                    double (&xb)[NS], double (&xc)[NS], bool need_logdet, EvalOut* out,
                                                    bool swap)
{
    static_assert(T == 32 || NS == 1, "a half warp holds x rows of at most 16 entries");
    const int lane = threadIdx.x & (T - 1), Tp = t2_pairs(c0), dg = c0 + 1;
    const int tb = (threadIdx.x & 31) & ~(T - 1);   // first lane of this tile inside the warp
    const unsigned tm = tile_mask<T>();
    if (c0 == 0 && lane == 1 && !swap) xa[0] = cy_max(xa[0], kMinVal);  // pyx:939 / :993 hits (x,x) without covariates
    auto pick = [&](const double (&x)[NS], int slot) -> double {
        if (NS == 1) return x[0];
        return slot ? x[NS - 1] : x[0];
    };
    for (int p = 0; p < c0; ++p) {
        const int src = tb | (p & (T - 1)), sl = p >> 5;
        const double ar = __shfl_sync(tm, pick(xa, sl), src);
        const double br = __shfl_sync(tm, pick(xb, sl), src);
        const double cr = FULL ? __shfl_sync(tm, pick(xc, sl), src) : 0.0;
        const double al2 = row2[3 * p], al4 = row2[3 * p + 1], alc = FULL ? row2[3 * p + 2] : 0.0;
        const int cb = t2_col(c0, p, p + 1) - (p + 1);
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const int j = lane + 32 * q;
            if (j > p && j <= dg) {
                double as = ar, bs = br, cs = cr;
                if (j <= c0) {
                    as = row2[cb + j]; bs = row2[Tp + cb + j];
                    if (FULL) cs = row2[2 * Tp + cb + j];
                }
                const bool clamp = (p == c0 - 1) && (j == dg) && !swap;   // de mode: x is no pivot (pg_eval.cuh)
                if (FULL) {
                    double v = (xc[q] + alc * ar * as) + al2 * (ar * cs + cr * as) + al2 * (br * bs) + al4 * (ar * bs + br * as);
                    if (clamp) v = cy_max(v, kMinVal);
                    xc[q] = v;
                }
                double v = (xb[q] + al4 * ar * as) + al2 * (ar * bs + br * as);
                if (clamp) v = cy_max(v, kMinVal);
                xb[q] = v;
                v = xa[q] + al2 * ar * as;
                if (clamp) v = cy_max(v, kMinVal);
                xa[q] = v;
            }
        }
    }
    const double app = __shfl_sync(tm, pick(xa, dg >> 5), tb | (dg & (T - 1)));
    const double bpp = __shfl_sync(tm, pick(xb, dg >> 5), tb | (dg & (T - 1)));
    const double cpp = FULL ? __shfl_sync(tm, pick(xc, dg >> 5), tb | (dg & (T - 1))) : 0.0;
    const double ar = __shfl_sync(tm, pick(xa, c0 >> 5), tb | (c0 & (T - 1)));
    const double br = __shfl_sync(tm, pick(xb, c0 >> 5), tb | (c0 & (T - 1)));
    const double cr = FULL ? __shfl_sync(tm, pick(xc, c0 >> 5), tb | (c0 & (T - 1))) : 0.0;
    if (swap) xrow_final_level_swapped<FULL>(c0, row2 + t2_fin(c0), app, bpp, cpp, ar, br, cr, need_logdet, out);
    else xrow_final_level<FULL>(row2 + t2_fin(c0), app, bpp, cpp, ar, br, cr, need_logdet, out);
}

// One precompute_mat-equivalent evaluation from the compressed moments (warp-collective).
// scratch (shared memory, per warp): 3 * k1p doubles for the level-0 x row, then NF2 for an interpolated table-2 row.
template <int NS, int T = 32>
__device__ __forceinline__ void eval_snp_compressed(const SolveArgs& a, const double* __restrict__ Zs, double lam,
                                                    int fixed_t, int full, int need_ll, double* scratch, EvalOut* e)
{
    const int lane = threadIdx.x & (T - 1), c0 = a.c0, k1 = c0 + 2, k1p = a.k1p, NF2 = a.t2.NF2;
    const int tb = (threadIdx.x & 31) & ~(T - 1);
    const unsigned tm = tile_mask<T>();
    double* xs = scratch;
    double* rowbuf = scratch + 3 * k1p;
    if (a.zsm) {
        if (full) solve_xrow_all<true, true, T>(a, Zs, lam, xs);
        else solve_xrow_all<false, true, T>(a, Zs, lam, xs);
    } else {
        if (full) solve_xrow_all<true, false, T>(a, Zs, lam, xs);
        else solve_xrow_all<false, false, T>(a, Zs, lam, xs);
    }
    const double* row2;
    if (fixed_t >= 0) {
        row2 = a.t2.fix2 + (size_t)fixed_t * NF2;
    } else {
        // Chebyshev weights: lane k < kNodes computes L_k, everybody collects the twelve by shuffle
        int iv;
        double xloc;
        interval_of(lam, &iv, &xloc);
        double Lk = 0.0;
        {
            const double* __restrict__ bk = a.t2.basis + (size_t)(lane < kNodes ? lane : 0) * kNodes;
            double t0 = 1.0, t1 = xloc;
            Lk = __ldg(bk) + __ldg(bk + 1) * xloc;
#pragma unroll
            for (int j = 2; j < kNodes; ++j) {
                const double t2 = 2.0 * xloc * t1 - t0;
                Lk = fma(__ldg(bk + j), t2, Lk);
                t0 = t1; t1 = t2;
            }
        }
        double L[kNodes];
#pragma unroll
        for (int k = 0; k < kNodes; ++k) L[k] = __shfl_sync(tm, Lk, tb | k);
        const double* __restrict__ base = a.t2.itab2 + (size_t)iv * kNodes * NF2;
        // two-power evaluations never read the third-power pivot columns: skip that third of the row
        const int skip0 = full ? NF2 : 3 * c0 + 2 * a.t2.Tp, skip1 = full ? NF2 : t2_fin(c0);
        for (int f0 = lane; f0 < NF2 - (skip1 - skip0); f0 += T) {
            const int f = f0 < skip0 ? f0 : f0 + (skip1 - skip0);
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < kNodes; ++k) v += L[k] * __ldg(base + (size_t)k * NF2 + f);
            rowbuf[f] = v;
        }
        row2 = rowbuf;
    }
    __syncwarp(tm);
    double xa[NS], xb[NS], xc[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        const int j = lane + 32 * q;
        const bool in = j < k1;
        const int r = j < c0 ? j : (j == c0 ? k1p - 1 : c0);   // Pab column order [W0, y, x] -> x-row scratch (y at k1p - 1)
        xa[q] = in ? xs[r] : 0.0;
        xb[q] = in ? xs[k1p + r] : 0.0;
        xc[q] = (in && full) ? xs[2 * k1p + r] : 0.0;
    }
    if (full) xrow_recursion_warp<true, NS, T>(c0, row2, xa, xb, xc, need_ll != 0, e, a.swap != 0);
    else xrow_recursion_warp<false, NS, T>(c0, row2, xa, xb, xc, need_ll != 0, e, a.swap != 0);
    __syncwarp(tm);
}

// one fixed-lambda evaluation of SNP g as fixed_phase_kernel left it (every lane reads the same words)
__device__ __forceinline__ void load_fixed(const SolveArgs& a, long long g, int t, EvalOut* e)
{
    const double* __restrict__ fx = a.FX + (size_t)t * kFxOut * a.ldF + g;
    e->yPy = __ldg(fx); e->yPPy = __ldg(fx + a.ldF); e->trP = __ldg(fx + 2 * a.ldF); e->logdetWHW = __ldg(fx + 3 * a.ldF);
    e->xPx = __ldg(fx + 4 * a.ldF); e->yPx = __ldg(fx + 5 * a.ldF);
    const double* __restrict__ fin = a.t2.fix2 + (size_t)t * a.t2.NF2 + t2_fin(a.c0);
    e->logdetH = __ldg(fin + 5); e->trH = __ldg(fin + 7); e->trHH = __ldg(fin + 8);
    e->yPPPy = NAN; e->trPP = NAN;   // fixed-lambda evaluations are never full (SnpSolver::request_fixed)
}

template <int NS, int MINB, int T = 32>
__global__ void __launch_bounds__(256, MINB) reml_solve_kernel(SolveArgs a)
{
    extern __shared__ double smem[];
    __shared__ __align__(16) unsigned char solver_mem[256 / T][(sizeof(SnpSolver) + 15) / 16 * 16];
    const int lane = threadIdx.x & (T - 1), warp = threadIdx.x / T;   // `warp` = this tile's index inside the CTA
    const int tb = (threadIdx.x & 31) & ~(T - 1);
    const unsigned tm = tile_mask<T>();
    const int k1p = a.k1p;
    const size_t per_warp = (size_t)((3 * k1p + a.t2.NF2 + 1) & ~1) + (a.zsm ? (size_t)k1p * a.Kcp : 0);
    double* scratch = smem + (size_t)warp * per_warp;
    double* zcopy = scratch + ((3 * k1p + a.t2.NF2 + 1) & ~1);
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(a.counter, 1ULL);
        g = __shfl_sync(tm, g, tb);
        if (g >= (unsigned long long)a.m) break;
        PG_BOUNDS(a.yrow < a.zrows && a.k1p - 1 <= a.zrows, "solver: slab rows");
        const double* __restrict__ Zs = a.Z + (size_t)g * a.zrows * a.Kcp;
        if (a.zsm) {
            // rows 0..k1p-2 are contiguous in the slab; the phenotype row follows them in the copy
            __syncwarp(tm);
            const int half_row = a.Kcp >> 1, total = k1p * half_row;
            const double2* __restrict__ src = reinterpret_cast<const double2*>(Zs);
            const double2* __restrict__ srcy = reinterpret_cast<const double2*>(Zs + (size_t)a.yrow * a.Kcp);
            double2* dst = reinterpret_cast<double2*>(zcopy);
            const int body = (k1p - 1) * half_row;
            for (int i = lane; i < total; i += T) dst[i] = i < body ? __ldg(src + i) : __ldg(srcy + (i - body));
            __syncwarp(tm);
            Zs = zcopy;
        }
        // The optimiser state (448 bytes, identical in every lane) lives in shared memory, one copy per warp: all lanes
        // run the state machine in lock step and store the same values, and the ~50 registers it would pin per thread
        // go to the evaluation instead (solve stage -10 %; with c0 <= 6 the kernel then fits 4 CTAs per SM: -27 %).
        SnpSolver& s = *reinterpret_cast<SnpSolver*>(solver_mem[warp]);
        s.init(a.n, a.c0, a.grid, /*defer_p=*/1);
        while (s.pending()) {
            EvalOut e;
            if (s.req_fixed() >= 0 && a.FX) load_fixed(a, (long long)g, s.req_fixed(), &e);   // fixed_phase_kernel did it
            else eval_snp_compressed<NS, T>(a, Zs, s.req_lambda(), s.req_fixed(), s.req_full(), s.req_ll(), scratch, &e);
            s.feed(e);
        }
        int st_bits = s.status;
        if (lane == 0) {
            const long long row = a.row0 + (long long)g;
            a.out[0][row] = s.beta; a.out[1][row] = s.se; a.out[2][row] = s.tau;
            a.out[3][row] = s.lambda; a.out[4][row] = s.F; a.out[5][row] = s.p;
            if (a.n_eval2) a.n_eval2[row] = s.n_eval2;
            if (a.n_eval3) a.n_eval3[row] = s.n_eval3;
        }
        if (a.lrt[0]) {
            // the ML optimisation of the alternative model on the same moments; its state takes over the REML solver's
            // shared-memory slot (the Wald row has been written)
            __syncwarp(tm);
            static_assert(sizeof(MlSolver) <= sizeof(SnpSolver), "MlSolver must fit the solver slot");
            MlSolver& ms = *reinterpret_cast<MlSolver*>(solver_mem[warp]);
            ms.init(a.n);
            while (ms.pending()) {
                EvalOut e;
                if (ms.req_fixed() >= 0 && a.FX) load_fixed(a, (long long)g, ms.req_fixed(), &e);
                else eval_snp_compressed<NS, T>(a, Zs, ms.req_lambda(), ms.req_fixed(), ms.req_full(), 0, scratch, &e);
                ms.feed(e);
            }
            if (lane == 0) {
                const long long row = a.row0 + (long long)g;
                const double D = 2.0 * (ms.best_ll - a.l_null);
                a.lrt[0][row] = ms.best_lambda; a.lrt[1][row] = ms.best_ll; a.lrt[2][row] = D;
                a.lrt[3][row] = chi2_sf_1(D);
            }
            st_bits |= ms.status;
        }
        if (lane == 0 && a.status) a.status[a.row0 + (long long)g] = st_bits;
    }
}

// The null model [W0] of one phenotype (reference lmm/lmm.py:176-190, commented there): ML lambda as lmm.calc_lambda
// finds it, tau = n / y^T P y, and the log-likelihood the LRT subtracts.  No genotype is involved: every evaluation
// is a table-2 row (level c0 of the [W0, y] block).  One warp; out4 = {lambda_null, tau_null, l_null, status}.
__global__ void null_model_kernel(int n, Tables2 t2, double* out4)
{
    __shared__ double rowbuf[16];
    __shared__ MlSolver ms;
    const int lane = threadIdx.x & 31, NF2 = t2.NF2, c0 = t2.c0;
    ms.init(n);
    __syncwarp();
    while (ms.pending()) {
        EvalOut e;
        const double lam = ms.req_lambda();
        const int fixed_t = ms.req_fixed();
        const double* fin;
        if (fixed_t >= 0) {
            fin = t2.fix2 + (size_t)fixed_t * NF2 + t2_fin(c0);
        } else {
            int iv;
            double L[kNodes];
            table_weights(t2.basis, lam, &iv, L);
            const double* base = t2.itab2 + (size_t)iv * kNodes * NF2 + t2_fin(c0);
            if (lane < 9) {
                double v = 0.0;
                for (int k = 0; k < kNodes; ++k) v += L[k] * base[(size_t)k * NF2 + lane];
                rowbuf[lane] = v;
            }
            __syncwarp();
            fin = rowbuf;
        }
        null_eval_from_row(fin, &e);
        __syncwarp();
        ms.feed(e);
        __syncwarp();
    }
    if (lane == 0) {
        out4[0] = ms.best_lambda; out4[1] = (double)n / ms.best_yPy; out4[2] = ms.best_ll; out4[3] = (double)ms.status;
    }
}

// p_wald = F(1, n - c0 - 1) survival function of F_wald (scipy.stats.f.sf, reference lmm/lmm.py:482), one thread per SNP:
// reml_solve_kernel leaves p = NaN (SnpSolver defer_p) so that the continued fraction is not run 32 lanes wide
// lnbeta = lgamma((nu+1)/2) - lgamma(nu/2) - lgamma(1/2), formed once per scan on the host in long double
__global__ void pvalue_kernel(const double* __restrict__ F, double* __restrict__ p, long long row0, long long m, double nu,
                              double lnbeta)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) p[row0 + i] = f_sf_1_pre(F[row0 + i], nu, lnbeta);
}

// table-2 rows: one thread eliminates the covariate levels of one table lambda (pg_eval.cuh: eliminate_w0y_row).  The serial
// form; the default is eliminate_tables_cta_kernel below (PG_ELIM_SERIAL=1 selects this one, the cross-check).
__global__ void eliminate_tables_kernel(int c0, int NF, int NF2, const double* __restrict__ fixtab,
                                        const double* __restrict__ itab, double* __restrict__ fix2,
                                        double* __restrict__ itab2, double* __restrict__ work)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= kNumTableRows) return;
    const int k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2;
    const double* src = (row < kNumFixed) ? fixtab + (size_t)row * NF : itab + (size_t)(row - kNumFixed) * NF;
    double* dst = (row < kNumFixed) ? fix2 + (size_t)row * NF2 : itab2 + (size_t)(row - kNumFixed) * NF2;
    eliminate_w0y_row(c0, src, work + (size_t)row * 3 * T0, dst);
}

// The same elimination with one CTA per table lambda: inside a level every trailing entry (r, s) is updated from the pivot
// column alone, so the entries are spread over the threads and the levels are separated by a barrier.  Every entry goes
// through the expression of eliminate_w0y_row (same operations, same order): the rows are bit-identical to the serial
// kernel's.  The serial kernel keeps 971 threads busy for ~c0^3 / 6 dependent entry updates each: 19 ms per design at
// c0 = 40 (n = 10 000), most of an lmm.pygemma call's set-up there; this one takes a fraction of a millisecond.
__global__ void __launch_bounds__(128) eliminate_tables_cta_kernel(int c0, int NF, int NF2, const double* __restrict__ fixtab,
                                                                    const double* __restrict__ itab, double* __restrict__ fix2,
                                                                    double* __restrict__ itab2, double* __restrict__ work_all)
{
    const int row = blockIdx.x;
    const int k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2, Tp = t2_pairs(c0);
    const double* row0 = (row < kNumFixed) ? fixtab + (size_t)row * NF : itab + (size_t)(row - kNumFixed) * NF;
    double* out = (row < kNumFixed) ? fix2 + (size_t)row * NF2 : itab2 + (size_t)(row - kNumFixed) * NF2;
    double* A = work_all + (size_t)row * 3 * T0;
    double* B = A + T0;
    double* C = A + 2 * T0;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int q = tid; q < 3 * T0; q += nth) A[q] = row0[q];
    double trP = row0[3 * T0], trPP = row0[3 * T0 + 1], logdet = 0.0;   // carried by thread 0
    __syncthreads();
    if (tid == 0 && c0 >= 1) A[0] = cy_max(A[0], kMinVal);  // pyx:939 / :993
    for (int i = 1; i <= c0; ++i) {
        __syncthreads();   // level i - 1 is complete (and the clamp above)
        const int p = i - 1, pp = tri(p, p);
        const double app = A[pp], bpp = B[pp], cpp = C[pp];
        const double inv = 1.0 / app;
        const double al2 = -inv, al4 = bpp / (app * app);
        const double alc = (cpp / (app * app)) - ((bpp * bpp) / (app * app * app));
        if (tid == 0) {
            trPP = trPP + (bpp / app) * (bpp / app) - 2 * (cpp / app);
            trP = trP - bpp / app;
            logdet += log(app);
            out[3 * p] = al2; out[3 * p + 1] = al4; out[3 * p + 2] = alc;
        }
        for (int s = i + tid; s <= c0; s += nth) {
            const int q = t2_col(c0, p, s), sp = tri(s, p);
            out[q] = A[sp]; out[Tp + q] = B[sp]; out[2 * Tp + q] = C[sp];
        }
        // trailing entries (r, s2), i <= s2 <= r <= c0: a warp per r, lanes over s2
        for (int r = i + (tid >> 5); r <= c0; r += (nth >> 5))
            for (int s2 = i + (tid & 31); s2 <= r; s2 += 32) {
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
    __syncthreads();
    if (tid == 0) {
        const int f = t2_fin(c0), yy = tri(c0, c0);
        out[f] = A[yy]; out[f + 1] = B[yy]; out[f + 2] = C[yy];
        out[f + 3] = trP; out[f + 4] = trPP; out[f + 5] = row0[3 * T0 + 2]; out[f + 6] = logdet;
        out[f + 7] = row0[3 * T0]; out[f + 8] = row0[3 * T0 + 1];
    }
}

template <int NS>
__global__ void probe_precompute_compressed_kernel(SolveArgs a, double lam, int fixed_t, int full, double* out9)
{
    extern __shared__ double smem[];
    EvalOut e;
    eval_snp_compressed<NS>(a, a.Z, lam, fixed_t, full, 1, smem, &e);   // a.zsm == 0: the probe reads the slab in place
    if ((threadIdx.x & 31) == 0) {
        out9[0] = e.yPy; out9[1] = e.yPPy; out9[2] = e.yPPPy; out9[3] = e.trP; out9[4] = e.trPP;
        out9[5] = e.logdetH; out9[6] = e.logdetWHW; out9[7] = e.xPx; out9[8] = e.yPx;
    }
}

}  // namespace pg
