// compress.cuh -- stage 2a of the scan: collapse each rotated genotype vector onto the compression nodes.
//
// For SNP s with rotated genotypes x (n doubles) the REML kernel needs, per node k of the plan
// (compress_plan.h), the moments
//     Z[s][row(j)][k] = sum_{l in segment(k)} L_k(log d_l) x_l w_jl     j = 0..klin-1   ([W0, y..] columns)
//     Z[s][c0][k]     = sum_l L_k(log d_l) x_l^2
// Row layout of one SNP's slab (z_row / z_xx_row below; k1p = (c0+2) rounded up to a multiple of 4):
//     rows 0..c0-1: x.w_j    row c0: x.x    rows c0+1..k1p-2: padding (stay zero)    row k1p-1+t: x.y_t, t = 0..q-1
// so that the k1p rows the solver reads for one phenotype are contiguous except for its own y row, and q phenotypes
// on the same covariates (pg_set_design_multi) share one compression: the linear columns are [W0, y_0 .. y_{q-1}].
// which replace the reference's per-SNP, per-lambda dsyrk over n samples (pygemma_model.pyx:938-943,:1002).
// For one segment this is a dense contraction  Z_seg (SNPs x (c0+1) kq)  =  X[:, l0:l1]  .  V[l0:l1, :]  with
// V[l, j kq + k] = L_k(log d_l) w_jl  -- the FP64 tensor pipe (DMMA m8n8k4) -- and the x^2 moments reuse the
// same A fragments squared in registers against B = L.  One CTA owns 128 SNPs x one segment x one group of
// up to 11 [W0, y] columns; the sample dimension streams through a 4-stage cp.async ring (X tile 128 x 16,
// V tile 16 x 128).  The rotated genotypes are read from HBM once per column group (once for c0 <= 10).
//
// COPY segments (isolated eigenvalues, one node each) are an elementwise product, compress_copy_kernel.
#pragma once

#include <cuda_runtime.h>

#include "compress_plan.h"
#include "pg_debug.cuh"

namespace pg {

constexpr int kCtSnps = 128;    // SNPs per CTA (8 warps x 16)
#ifndef PG_CT_L
#define PG_CT_L 16
#endif
constexpr int kCtL = PG_CT_L;   // eigen-indices per pipeline stage (16 or 32; rows of xr / V are padded to 32)
constexpr int kCtStages = (kCtL == 16) ? 4 : 3;
constexpr int kXsPitch = kCtL + 4;   // doubles; pitch mod 16 == 4 -> conflict-free A-fragment reads
constexpr int kVsPitch = 132;   // doubles; 132 mod 16 == 4 -> conflict-free B-fragment reads
constexpr int kLinTiles = 14;   // 8-column tiles of the linear moments (112 >= 11 * kCq)
constexpr int kSqTiles = 2;     // 8-column tiles of the x^2 moments (16 >= kCq)
constexpr int kGroupCols = 8 * (kLinTiles + kSqTiles);  // 128 columns of V per column group
constexpr int kJGroup = (8 * kLinTiles) / kCq;          // 11 [W0, y] columns per group
constexpr int kCtStageDoubles = kCtSnps * kXsPitch + kCtL * kVsPitch;
constexpr size_t kCtSmemBytes = sizeof(double) * (size_t)kCtStages * kCtStageDoubles;

struct CompItem {
    int l0, l1;   // eigen-index range of the segment
    int kb, kq;   // first node, node count (1 or kCq)
    int j0, nj;   // [W0, y] columns of this group
    int vcol0;    // first column of the group inside a V row
    int pad;
};

// slab row of linear column j of [W0, y_0, y_1, ...]
__host__ __device__ __forceinline__ int z_row(int j, int c0, int k1p) { return j < c0 ? j : k1p - 1 + (j - c0); }

struct DevPlan {
    int Kc = 0, Kcp = 0;        // nodes, padded to a multiple of 32 (padding nodes: d = 0, moments 0)
    double* nodes = nullptr;    // [Kcp]
    double* H = nullptr;        // [Kcp][kFxCols] powers of 1/(lambda_t d_k + 1) at the fixed lambdas (reml_solve.cuh)
    double* Lw = nullptr;       // [n][kCq]
    int* seg_kq = nullptr;      // [n] kq of the COMPRESS segment of l, 0 for COPY rows
    double* V = nullptr;        // [npad16][vpitch]
    int vpitch = 0, ngroups = 0, npad16 = 0;   // npad16: rows of V, n rounded up to 32
    int klin = 0;               // linear columns the items / V were built for: c0 + q
    CompItem* items = nullptr;
    int nitems = 0;
    int* copy_l = nullptr;      // COPY rows: eigen index and node
    int* copy_node = nullptr;
    int ncopy = 0;
};

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cpa16_zfill(void* smem_dst, const void* gsrc, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cpa_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// V[l][g*128 + j*kq + k] = Lw[l][k] * wy[j0_g + j][l]  (k < kq, j < nj_g);  V[l][g*128 + 112 + k] = Lw[l][k] (g == 0)
__global__ void build_v_kernel(int n, int klin, const double* __restrict__ Lw, const int* __restrict__ seg_kq,
                               const double* __restrict__ wy, long long ldw, double* __restrict__ V, int vpitch,
                               int ngroups)
{
    const int l = blockIdx.x;
    if (l >= n) return;
    const int kq = seg_kq[l];
    const int k0 = klin;
    for (int col = threadIdx.x; col < vpitch; col += blockDim.x) {
        const int g = col / kGroupCols, c = col - g * kGroupCols;
        double v = 0.0;
        if (kq > 0) {
            if (c < 8 * kLinTiles) {
                const int jl = c / kq, k = c - jl * kq;
                const int j = g * kJGroup + jl;
                if (jl < kJGroup && j < k0) v = Lw[(size_t)l * kCq + k] * wy[(size_t)j * ldw + l];
            } else if (g == 0) {
                const int k = c - 8 * kLinTiles;
                if (k < kq) v = Lw[(size_t)l * kCq + k];
            }
        }
        V[(size_t)l * vpitch + col] = v;
    }
    (void)ngroups;
}

// Fused moments (rotate_i8_tc2.cuh, FUSE): the compact right-hand operand of G = U V.  Column j * kCq + k holds
// L_k(log d_l) w_jl for the rows l of COMPRESS segments (zero on COPY rows); a segment with kq = kCq nodes contributes the
// n x (klin kCq) block  G_s = U[:, l0:l1] . Vc[l0:l1, :]  and a one-node segment the columns k = 0 (leading dimension
// n kCq): one DGEMM per segment.  Column order linear-column-major, node fastest: the eight columns an epilogue thread of
// the fused rotation holds are then (mostly) eight consecutive nodes of one slab row, i.e. whole 32-byte sectors.
__global__ void build_vc_kernel(int n, int klin, const double* __restrict__ Lw, const int* __restrict__ seg_kq,
                                const double* __restrict__ wy, long long ldw, double* __restrict__ Vc)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const int col = blockIdx.y;            // j * kCq + k
    const int j = col / kCq, k = col - j * kCq;
    Vc[(size_t)col * n + l] = seg_kq[l] > 0 ? Lw[(size_t)l * kCq + k] * wy[(size_t)j * ldw + l] : 0.0;
}

__global__ void __launch_bounds__(256, 1)
compress_dmma_kernel(const double* __restrict__ xr, long long ldx, long long mb, const CompItem* __restrict__ items,
                     const double* __restrict__ V, int vpitch, int c0, int k1p, int zrows, int Kcp, double* __restrict__ Z,
                     int ntiles)
{
    extern __shared__ __align__(16) double csm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CompItem it = items[blockIdx.x / ntiles];
    const long long snp0 = (long long)(blockIdx.x % ntiles) * kCtSnps;
    const int nt_lin = (it.nj * it.kq + 7) >> 3;
    const int nt_sq = (it.j0 == 0) ? ((it.kq + 7) >> 3) : 0;
    const int c_begin = it.l0 & ~(kCtL - 1);
    const int c_end = (it.l1 + kCtL - 1) & ~(kCtL - 1);
    const int nchunks = (c_end - c_begin) / kCtL;

    auto load = [&](int stage, int chunk) {
        double* Xs = csm + (size_t)stage * kCtStageDoubles;
        double* Vs = Xs + kCtSnps * kXsPitch;
        const int lbase = c_begin + chunk * kCtL;
#pragma unroll
        for (int c = tid; c < kCtSnps * (kCtL / 2); c += 256) {
            const int r = c / (kCtL / 2), o = (c % (kCtL / 2)) * 2;
            const long long snp = snp0 + r;
            const bool valid = snp < mb;
            PG_BOUNDS(lbase >= 0 && lbase + o + 2 <= ldx && r < kCtSnps && o + 2 <= kCtL, "compress: X tile load");
            cpa16_zfill(Xs + r * kXsPitch + o, xr + (size_t)(valid ? snp : 0) * ldx + lbase + o, valid ? 16 : 0);
        }
#pragma unroll
        for (int c = tid; c < kCtL * (kGroupCols / 2); c += 256) {
            const int r = c >> 6, o = (c & 63) * 2;
            PG_BOUNDS(r < kCtL && o + 2 <= kGroupCols && it.vcol0 + o + 2 <= vpitch, "compress: V tile load");
            cpa16(Vs + r * kVsPitch + o, V + (size_t)(lbase + r) * vpitch + it.vcol0 + o);
        }
    };

    double acc[2][kLinTiles + kSqTiles][2];
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int t = 0; t < kLinTiles + kSqTiles; ++t) { acc[b][t][0] = 0.0; acc[b][t][1] = 0.0; }

    for (int s = 0; s < kCtStages - 1; ++s) {
        if (s < nchunks) load(s, s);
        cpa_commit();
    }
    const int arow = warp * 16 + (lane >> 2), acol = lane & 3;
    for (int ch = 0; ch < nchunks; ++ch) {
        cpa_wait<kCtStages - 2>();
        __syncthreads();
        const int nx = ch + kCtStages - 1;
        if (nx < nchunks) load(nx % kCtStages, nx);
        cpa_commit();
        const double* Xs = csm + (size_t)(ch % kCtStages) * kCtStageDoubles;
        const double* Vs = Xs + kCtSnps * kXsPitch;
        const int lbase = c_begin + ch * kCtL;
#pragma unroll
        for (int kk = 0; kk < kCtL / 4; ++kk) {
            const int l = lbase + kk * 4 + acol;
            const bool inb = (l >= it.l0) && (l < it.l1);
            double a0 = Xs[arow * kXsPitch + kk * 4 + acol];
            double a1 = Xs[(arow + 8) * kXsPitch + kk * 4 + acol];
            a0 = inb ? a0 : 0.0;
            a1 = inb ? a1 : 0.0;
            const double* vrow = Vs + (kk * 4 + acol) * kVsPitch + (lane >> 2);
#pragma unroll
            for (int t = 0; t < kLinTiles; ++t) {
                if (t < nt_lin) {
                    const double b = vrow[8 * t];
                    dmma884(acc[0][t], a0, b);
                    dmma884(acc[1][t], a1, b);
                }
            }
            const double a0s = a0 * a0, a1s = a1 * a1;
#pragma unroll
            for (int t = 0; t < kSqTiles; ++t) {
                if (t < nt_sq) {
                    const double b = vrow[8 * (kLinTiles + t)];
                    dmma884(acc[0][kLinTiles + t], a0s, b);
                    dmma884(acc[1][kLinTiles + t], a1s, b);
                }
            }
        }
    }
    cpa_wait<0>();

    const int ncol_lin = it.nj * it.kq;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const long long snp = snp0 + warp * 16 + b * 8 + (lane >> 2);
        if (snp >= mb) continue;
        double* Zs = Z + (size_t)snp * zrows * Kcp;
#pragma unroll
        for (int t = 0; t < kLinTiles; ++t) {
            if (t < nt_lin) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * t + 2 * (lane & 3) + e;
                    if (col < ncol_lin) {
                        const int jl = col / it.kq, k = col - jl * it.kq;
                        PG_BOUNDS(z_row(it.j0 + jl, c0, k1p) < zrows && it.kb + k < Kcp, "compress: linear moment store");
                        Zs[(size_t)z_row(it.j0 + jl, c0, k1p) * Kcp + it.kb + k] = acc[b][t][e];
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < kSqTiles; ++t) {
            if (t < nt_sq) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * t + 2 * (lane & 3) + e;
                    PG_BOUNDS(col >= it.kq || (c0 < zrows && it.kb + col < Kcp), "compress: x^2 moment store");
                    if (col < it.kq) Zs[(size_t)c0 * Kcp + it.kb + col] = acc[b][kLinTiles + t][e];
                }
            }
        }
    }
}

// COPY rows: node(l) holds the single eigenvalue d_l:  Z[s][row(j)][node] = x_l w_jl,  Z[s][c0][node] = x_l^2
__global__ void __launch_bounds__(128)
compress_copy_kernel(const double* __restrict__ xr, long long ldx, long long mb, const int* __restrict__ copy_l,
                     const int* __restrict__ copy_node, int ncopy, const double* __restrict__ wy, long long ldw, int c0,
                     int klin, int k1p, int zrows, int Kcp, double* __restrict__ Z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncopy) return;
    const int l = copy_l[i], node = copy_node[i];
    const long long s0 = (long long)blockIdx.y * 8;
    for (int r = 0; r < 8; ++r) {
        const long long snp = s0 + r;
        if (snp >= mb) break;
        const double x = xr[(size_t)snp * ldx + l];
        double* Zs = Z + (size_t)snp * zrows * Kcp + node;
        for (int j = 0; j < klin; ++j) Zs[(size_t)z_row(j, c0, k1p) * Kcp] = x * wy[(size_t)j * ldw + l];
        Zs[(size_t)c0 * Kcp] = x * x;
    }
}

}  // namespace pg
