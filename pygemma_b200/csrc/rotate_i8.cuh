// rotate_i8.cuh -- exact integer-split rotation for int8 dosages.
//
// U^T x for an integer genotype vector x is a sum of (double) * (small integer).  Each eigenvector
// (row i of U^T) is rewritten in fixed point relative to its own largest entry,
//     U[j,i] * 2^-e_i  =  sum_{t=1..7} q_t[i][j] * 256^-t ,   q_t in int8 (balanced base-256 digits),
// with e_i chosen so that |U[j,i]| 2^-e_i < 1/4 (the leading digit then stays within [-65, 65]).  The
// truncation is 2^-57 of 2^e_i, i.e. at most 2^-54 of the row maximum -- finer than the 2^-53 relative
// rounding of the entries that dominate the dot product.  The seven digit planes times the int8 genotypes
// are seven int8 x int8 -> int32 tensor-core GEMMs whose results are exact integers; they are recombined
// in int64 and rounded to fp64 once.  The result is at least as accurate as an FP64 GEMM and runs on the
// int8 tensor pipe (B200: ~3.6 Pop/s measured vs ~36 TFLOP/s FP64), which is what lifts the rotation from
// ~180 k to >1.5 M SNPs/s at n = 10 000.
#pragma once

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace pg {

// Digit planes per eigenvector: U is held as kSlices balanced base-256 digits of |u| 2^-e < 1/4, i.e. 8 kSlices bits of
// fixed point below the largest entry of the eigenvector.  7 planes (56 bits) lose nothing an FP64 U carries for its
// large entries; 6 planes (48 bits, build with -DPG_SLICES=6) round each entry at 2^-50 of the column scale, which is
// the size of the rounding error an FP64 GEMM commits itself (measured in DESIGN.md) for 6/7 of the tensor work.
#ifndef PG_SLICES
#define PG_SLICES 7
#endif
constexpr int kSlices = PG_SLICES;
static_assert(kSlices == 6 || kSlices == 7, "6 or 7 digit planes");
constexpr double kLoScale = (kSlices == 7) ? 2.3283064365386963e-10 /* 2^-32 */ : 5.9604644775390625e-08 /* 2^-24 */;
constexpr int kLoShift = 8 * (kSlices - 3);   // value = (hi + lo 2^-kLoShift) 2^(e-24), hi: planes 0..2, lo: planes 3..

// one CTA per vector i: find e_i, write the digit planes
// U holds nvec vectors of n entries (the n eigenvectors; the columns of G = U V for the fused moments, rotate_i8_tc2.cuh);
// vector i is at U + i*n when u_cols_contig (column-major), else strided (U + i, stride nvec)
__global__ void __launch_bounds__(256) slice_u_kernel(const double* __restrict__ U, int u_cols_contig, int n, int npad,
                                                       int ldk, int8_t* __restrict__ planes /* [npad][kSlices][ldk]: the planes of a vector are adjacent rows */,
                                                       int* __restrict__ exps, int nvec)
{
    const int i = blockIdx.x;
    __shared__ double red[8];
    __shared__ int e_sh;
    const size_t stride = u_cols_contig ? 1 : (size_t)nvec;
    const double* u = u_cols_contig ? U + (size_t)i * n : U + i;
    double mx = 0.0;
    if (i < nvec)
        for (int j = threadIdx.x; j < n; j += blockDim.x) mx = fmax(mx, fabs(u[(size_t)j * stride]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        // |u| * 2^-e < 0.25 for every entry
        e_sh = (m > 0.0 && isfinite(m)) ? ilogb(m) + 3 : 0;
        if (i < nvec) exps[i] = e_sh;
    }
    __syncthreads();
    const int e = e_sh;
    (void)npad;
    for (int j = threadIdx.x; j < ldk; j += blockDim.x) {
        long long Q = 0;
        if (i < nvec && j < n) Q = llrint(ldexp(u[(size_t)j * stride], 8 * kSlices - e));
        int8_t dig[kSlices];
#pragma unroll
        for (int t = kSlices - 1; t >= 1; --t) {
            long long r = ((Q + 128) & 255) - 128;  // balanced digit in [-128, 127]
            dig[t] = (int8_t)r;
            Q = (Q - r) >> 8;
        }
        dig[0] = (int8_t)Q;  // |Q| <= 65 here
#pragma unroll
        for (int t = 0; t < kSlices; ++t) planes[((size_t)i * kSlices + t) * ldk + j] = dig[t];
    }
}

// int8 genotype block, any layout -> SNP-major int8 [g][ldk] with zero padding (K-contiguous GEMM operand)
__global__ void stage_i8_kernel(const int8_t* __restrict__ src, long long ld, int layout, int n, long long mb, int ldk,
                                int8_t* __restrict__ dst)
{
    __shared__ int8_t tile[64][65];
    const int tx = threadIdx.x, ty = threadIdx.y;  // 64 x 4
    const long long g0 = (long long)blockIdx.x * 64;
    const int j0 = blockIdx.y * 64;
    if (layout == 0) {
        for (int r = ty; r < 64; r += 4) {
            const int j = j0 + r;
            const long long g = g0 + tx;
            tile[r][tx] = (j < n && g < mb) ? src[(size_t)j * ld + g] : (int8_t)0;
        }
        __syncthreads();
        for (int r = ty; r < 64; r += 4) {
            const long long g = g0 + r;
            const int j = j0 + tx;
            if (g < mb && j < ldk) dst[(size_t)g * ldk + j] = tile[tx][r];
        }
    } else {
        for (int r = ty; r < 64; r += 4) {
            const long long g = g0 + r;
            const int j = j0 + tx;
            if (g < mb && j < ldk) dst[(size_t)g * ldk + j] = (j < n) ? src[(size_t)g * ld + j] : (int8_t)0;
        }
    }
}

// P: [g][npad*kSlices] int32 (column-major (npad*kSlices) x mb as cuBLAS writes it; row i*kSlices + t = plane t of
// eigenvector i).  xr[g*n + i] fp64.
__global__ void __launch_bounds__(256) combine_i8_kernel(const int32_t* __restrict__ P, const int* __restrict__ exps,
                                                          int n, int npad, long long mb, double* __restrict__ xr, long long ldx)
{
    const long long g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || g >= mb) return;
    const int32_t* p = P + (size_t)g * kSlices * npad + (size_t)i * kSlices;
    // sum_t P_t 256^(kSlices-1-t): planes 0..2 (<= 2^46) and planes 3.. (<= 2^54) separately, then one fp64 sum
    long long hi = 0, lo = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) hi = hi * 256 + (long long)p[t];
#pragma unroll
    for (int t = 3; t < kSlices; ++t) lo = lo * 256 + (long long)p[t];
    const double v = (double)hi + ldexp((double)lo, -kLoShift);
    xr[(size_t)g * ldx + i] = ldexp(v, exps[i] - 24);
}


// ------------------------------------------------------------------------------------------------
// Level coding of float genotypes.  Callers of the reference usually hand over dosages as floats, raw
// (0.0 / 1.0 / 2.0) or standardised per SNP ((g - mean) / std, experiments/wtccc/run_pygemma.py:432): every column
// then takes at most three distinct, equally spaced values  x = v0 + s * g,  g in {0, 1, 2}.  The rotation is linear,
//     U^T x = v0 * (U^T 1) + s * (U^T g),
// so such columns go through the exact int8-split tensor-core path on the codes g and an affine fix-up in the
// recombination, instead of the 8x slower FP64 GEMM.  "Equally spaced" is tested to double rounding
// (|v2 - 2 v1 + v0| <= 2^-50 max|v|): standardising in float64 produces exactly this much deviation.  Columns whose
// three levels are NOT equally spaced (standardised in float32: ~1e-7 off; or any three-valued coding) are
//     x = v0 + s g + eps [g = 2] ,   U^T x = v0 (U^T 1) + s (U^T g) + eps (U^T [g = 2]) :
// a second exact int8 rotation of the indicator [g = 2] is accumulated with weight eps (2x the tensor work, still
// ~4x faster than the FP64 GEMM).  The same two-component form covers FOUR-valued columns made of three equally
// spaced levels plus one outlier -- mean-imputed dosages {0, 1, 2, mean} as the reference's callers produce them
// (SimpleImputer(strategy='mean'), experiments/animal_gwas/run_gwas.py:93-94), also after standardisation: the
// outlier gets code 0 and the indicator carries eps = outlier - v0.  Any column with NaN/Inf, more than four levels,
// or four levels without such a triple sends the whole block down the FP64 path.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxLevels = 4;

struct LevelInfo {
    double v0, s;   // x = v0 + s * code + eps * indicator
    double eps;     // weight of the indicator component, 0 when the column is an exact affine image of its code
    double t[3];    // level index = (x > t[0]) + (x > t[1]) + (x > t[2]): midpoints of the sorted levels (+inf when absent)
    unsigned char code_of[kMaxLevels];  // code (0, 1, 2) of each sorted level
    unsigned char ind_of[kMaxLevels];   // indicator (0 / 1) of each sorted level
    int nlev;       // 1..4: codeable; 0: not codeable
    int pad;
};

struct LevelPartial {
    double v[kMaxLevels];
    int k;      // number of distinct values seen in this sample chunk; -1: more than four or a non-finite value
    int pad;
};

__device__ __forceinline__ void level_insert(double x, double (&v)[kMaxLevels], int& k)
{
    if (k < 0) return;
    if (!isfinite(x)) { k = -1; return; }
    if ((k > 0 && x == v[0]) || (k > 1 && x == v[1]) || (k > 2 && x == v[2]) || (k > 3 && x == v[3])) return;
    if (k == kMaxLevels) { k = -1; return; }
    v[k++] = x;
}

// pass 1: thread (g, chunk) collects the distinct values of SNP column g over one chunk of kLevelChunk samples.
// Element (j, g) at src[j*ld + g] (layout 0, sample-major: coalesced over g) or src[g*ld + j] (layout 1).
constexpr int kLevelChunk = 256;
template <typename T>
__global__ void __launch_bounds__(128) find_levels_kernel(const T* __restrict__ src, long long ld, int layout, int n,
                                                           long long mb, LevelPartial* __restrict__ part)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= mb) return;
    const int j0 = blockIdx.y * kLevelChunk, j1 = min(n, j0 + kLevelChunk);
    const size_t step = layout == 0 ? (size_t)ld : 1, base = layout == 0 ? (size_t)g : (size_t)g * ld;
    double v[kMaxLevels] = {0.0, 0.0, 0.0, 0.0};
    int k = 0;
    int j = j0;
    for (; j + 8 <= j1; j += 8) {
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (double)src[base + (size_t)(j + u) * step];  // eight loads in flight
#pragma unroll
        for (int u = 0; u < 8; ++u) level_insert(x[u], v, k);
    }
    for (; j < j1; ++j) level_insert((double)src[base + (size_t)j * step], v, k);
    LevelPartial lp;
    lp.v[0] = v[0]; lp.v[1] = v[1]; lp.v[2] = v[2]; lp.v[3] = v[3]; lp.k = k; lp.pad = 0;
    part[(size_t)blockIdx.y * mb + g] = lp;
}

// pass 2: merge the chunk results of one SNP, sort the levels, test the spacing
__global__ void merge_levels_kernel(const LevelPartial* __restrict__ part, int nchunks, long long mb, double tol,
                                    LevelInfo* __restrict__ info, int* __restrict__ n_bad /* [0]: not codeable, [1]: need eps */)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= mb) return;
    double v[kMaxLevels] = {0.0, 0.0, 0.0, 0.0};
    int k = 0;
    for (int c = 0; c < nchunks; ++c) {
        const LevelPartial lp = part[(size_t)c * mb + g];
        if (lp.k < 0) { k = -1; break; }
        for (int u = 0; u < lp.k; ++u) level_insert(lp.v[u], v, k);
        if (k < 0) break;
    }
    LevelInfo li;
    li.v0 = 0.0; li.s = 0.0; li.eps = 0.0; li.nlev = 0; li.pad = 0;
    for (int u = 0; u < 3; ++u) li.t[u] = INFINITY;
    for (int u = 0; u < kMaxLevels; ++u) { li.code_of[u] = 0; li.ind_of[u] = 0; }
    if (k > 0) {
        for (int a = 1; a < k; ++a)   // insertion sort, ascending
            for (int b = a; b > 0 && v[b] < v[b - 1]; --b) { const double t = v[b]; v[b] = v[b - 1]; v[b - 1] = t; }
        for (int u = 0; u + 1 < k; ++u) li.t[u] = 0.5 * v[u] + 0.5 * v[u + 1];
        const double mx = fmax(fabs(v[0]), fabs(v[k - 1]));
        li.nlev = k;
        li.v0 = v[0];
        if (k == 2) { li.s = v[1] - v[0]; li.code_of[1] = 1; }
        if (k == 3) {
            li.s = v[1] - v[0];
            li.code_of[1] = 1; li.code_of[2] = 2;
            if (fabs((v[2] - v[1]) - (v[1] - v[0])) > tol * mx) {   // unequal spacing: the top level carries the indicator
                li.ind_of[2] = 1;
                li.eps = (v[2] - v[0]) - 2.0 * li.s;
            }
        }
        if (k == 4) {
            // three equally spaced levels plus one outlier (mean-imputed dosages): outlier -> code 0 + indicator
            int out = -1;
            for (int o = 0; o < 4 && out < 0; ++o) {
                double w[3];
                int q = 0;
                for (int u = 0; u < 4; ++u)
                    if (u != o) w[q++] = v[u];
                if (fabs((w[2] - w[1]) - (w[1] - w[0])) <= tol * mx) out = o;
            }
            if (out < 0) {
                li.nlev = 0;
            } else {
                int q = 0;
                double w0 = 0.0, w1 = 0.0;
                for (int u = 0; u < 4; ++u) {
                    if (u == out) continue;
                    if (q == 0) w0 = v[u];
                    if (q == 1) w1 = v[u];
                    li.code_of[u] = (unsigned char)q++;
                }
                li.v0 = w0; li.s = w1 - w0;
                li.code_of[out] = 0; li.ind_of[out] = 1;
                li.eps = v[out] - w0;
            }
        }
        if (li.nlev > 0 && li.eps != 0.0) atomicAdd(n_bad + 1, 1);
    }
    if (li.nlev == 0) atomicAdd(n_bad, 1);
    info[g] = li;
}

// codes in the layout of the input: sample-major (n x mb, ld = mb) or SNP-major (mb x n, ld = n)
// indicator != 0: write the indicator of the level instead of its code (operand of the eps rotation)
template <typename T>
__global__ void encode_levels_kernel(const T* __restrict__ src, long long ld, int layout, int n, long long mb,
                                     const LevelInfo* __restrict__ info, int8_t* __restrict__ codes, int indicator)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * 64;
    if (g >= mb) return;
    const LevelInfo li = info[g];
    const int j1 = min(n, j0 + 64);
    int j = j0;
    for (; j + 8 <= j1; j += 8) {
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (double)src[layout == 0 ? (size_t)(j + u) * ld + g : (size_t)g * ld + j + u];
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
            const int idx = (x[u] > li.t[0]) + (x[u] > li.t[1]) + (x[u] > li.t[2]);
            codes[layout == 0 ? (size_t)(j + u) * mb + g : (size_t)g * n + j + u] =
                (int8_t)(indicator ? li.ind_of[idx] : li.code_of[idx]);
        }
    }
    for (; j < j1; ++j) {
        const size_t si = layout == 0 ? (size_t)j * ld + g : (size_t)g * ld + j;
        const size_t di = layout == 0 ? (size_t)j * mb + g : (size_t)g * n + j;
        const double xv = (double)src[si];
        const int idx = (xv > li.t[0]) + (xv > li.t[1]) + (xv > li.t[2]);
        codes[di] = (int8_t)(indicator ? li.ind_of[idx] : li.code_of[idx]);
    }
}

// recombination with the affine fix-up: xr[g][i] = v0_g * u1_i + s_g * (U^T code_g)_i
// accumulate != 0 (second pass, indicator rotation): xr[g][i] += eps_g * (U^T indicator_g)_i
__global__ void __launch_bounds__(256) combine_i8_affine_kernel(const int32_t* __restrict__ P, const int* __restrict__ exps,
                                                                 int n, int npad, long long mb, double* __restrict__ xr,
                                                                 long long ldx, const LevelInfo* __restrict__ info,
                                                                 const double* __restrict__ u1, int accumulate)
{
    const long long g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || g >= mb) return;
    const int32_t* p = P + (size_t)g * kSlices * npad + (size_t)i * kSlices;
    long long hi = 0, lo = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) hi = hi * 256 + (long long)p[t];
#pragma unroll
    for (int t = 3; t < kSlices; ++t) lo = lo * 256 + (long long)p[t];
    const double r = ldexp((double)hi + ldexp((double)lo, -kLoShift), exps[i] - 24);
    const LevelInfo li = info[g];
    double* dst = xr + (size_t)g * ldx + i;
    if (accumulate) {
        if (li.eps != 0.0) *dst = fma(li.eps, r, *dst);
    } else {
        *dst = fma(li.s, r, li.v0 * u1[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// PLINK .bed ingest.  A .bed body is SNP-major: ceil(n/4) bytes per SNP, four samples per byte, sample
// 4b+i in bits 2i..2i+1; 00 = homozygous A1, 01 = missing, 10 = heterozygous, 11 = homozygous A2.  The reference's
// callers read it with pysnptools (NaN for missing), impute the column mean (SimpleImputer(strategy='mean'),
// experiments/benchmarks/benchmarks.py:22, experiments/animal_gwas/run_gwas.py:93-94) and optionally standardise.
// Here the packed block is uploaded as it is (2 bits per genotype over PCIe) and one CTA per SNP
//   mode 0: counts the genotype classes, writes the dosage codes (missing -> 0) straight into the K-contiguous int8
//           operand of the rotation, and the affine description of the imputed (standardised) column
//               x = v0 + s * code + eps * [missing]      (v0 = 0, s = 1, eps = mean;  standardised: (x - mean) / sd)
//   mode 1: writes the missing indicator (operand of the eps rotation)
// so that imputed .bed genotypes run on the exact int8 tensor path without ever existing as floats.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bed_decode_kernel(const uint8_t* __restrict__ bed, long long ld_bytes, int n,
                                                          long long mb, int ldk, int count_a1, int standardize, int mode,
                                                          int8_t* __restrict__ x8, LevelInfo* __restrict__ info,
                                                          int* __restrict__ n_bad)
{
    const long long g = blockIdx.x;
    if (g >= mb) return;
    const uint8_t* row = bed + (size_t)g * ld_bytes;
    const int nbytes = (n + 3) >> 2;
    __shared__ int red[3][4];
    int c1 = 0, c2 = 0, cm = 0;
    uint32_t* out = reinterpret_cast<uint32_t*>(x8 + (size_t)g * ldk);   // ldk is a multiple of 128: 4-byte aligned rows
    for (int b = threadIdx.x; b < ldk / 4; b += blockDim.x) {
        uint32_t word = 0;
        if (b < nbytes) {
            const unsigned v = row[b];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (4 * b + i >= n) break;
                const unsigned two = (v >> (2 * i)) & 3u;
                int code = 0, miss = 0;
                if (two == 1u) miss = 1;
                else if (two == 2u) code = 1;
                else code = ((two == 0u) == (count_a1 != 0)) ? 2 : 0;
                c1 += (code == 1); c2 += (code == 2); cm += miss;
                word |= (uint32_t)(mode ? miss : code) << (8 * i);
            }
        }
        out[b] = word;
    }
    if (mode != 0) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        cm += __shfl_xor_sync(0xffffffffu, cm, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = c1; red[1][threadIdx.x >> 5] = c2; red[2][threadIdx.x >> 5] = cm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t1 = 0, t2 = 0, tm = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t1 += red[0][w]; t2 += red[1][w]; tm += red[2][w]; }
        const int nobs = n - tm, t0 = nobs - t1 - t2;
        const double mu = nobs > 0 ? (double)(t1 + 2 * t2) / (double)nobs : 0.0;   // np.nanmean; all-missing column -> 0
        LevelInfo li;
        li.pad = 0; li.nlev = 4;
        for (int u = 0; u < 3; ++u) li.t[u] = INFINITY;
        for (int u = 0; u < kMaxLevels; ++u) { li.code_of[u] = 0; li.ind_of[u] = 0; }
        if (!standardize) {
            li.v0 = 0.0; li.s = 1.0; li.eps = mu;
        } else {
            // population variance of the imputed column (imputed entries sit on the mean); sd == 0 -> 1 (StandardScaler)
            const double var = ((double)t0 * mu * mu + (double)t1 * (1.0 - mu) * (1.0 - mu) + (double)t2 * (2.0 - mu) * (2.0 - mu)) / (double)n;
            const double sd = var > 0.0 ? sqrt(var) : 1.0;
            li.s = 1.0 / sd; li.v0 = -mu / sd; li.eps = mu / sd;
        }
        if (tm == 0) li.eps = 0.0;
        else atomicAdd(n_bad + 1, 1);
        info[g] = li;
    }
}

// u1 = U^T 1: column sums of U (eigenvector i at U + i*n when cols_contig, else strided)
__global__ void __launch_bounds__(256) column_sums_kernel(const double* __restrict__ U, int cols_contig, int n,
                                                           double* __restrict__ u1)
{
    const int i = blockIdx.x;
    __shared__ double red[8];
    const size_t stride = cols_contig ? 1 : (size_t)n;
    const double* u = cols_contig ? U + (size_t)i * n : U + i;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) s += u[(size_t)j * stride];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        u1[i] = t;
    }
}

}  // namespace pg
