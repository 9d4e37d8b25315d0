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

constexpr int kSlices = 7;

// one CTA per eigenvector i: find e_i, write the digit planes
// U is n x n; eigenvector i is at U + i*n when u_cols_contig (column-major U), else strided (U + i, stride n)
__global__ void __launch_bounds__(256) slice_u_kernel(const double* __restrict__ U, int u_cols_contig, int n, int npad,
                                                       int ldk, int8_t* __restrict__ planes /* [kSlices][npad][ldk] */,
                                                       int* __restrict__ exps)
{
    const int i = blockIdx.x;
    __shared__ double red[8];
    __shared__ int e_sh;
    const size_t stride = u_cols_contig ? 1 : (size_t)n;
    const double* u = u_cols_contig ? U + (size_t)i * n : U + i;
    double mx = 0.0;
    if (i < n)
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
        if (i < n) exps[i] = e_sh;
    }
    __syncthreads();
    const int e = e_sh;
    const size_t plane = (size_t)npad * ldk;
    for (int j = threadIdx.x; j < ldk; j += blockDim.x) {
        long long Q = 0;
        if (i < n && j < n) Q = llrint(ldexp(u[(size_t)j * stride], 8 * kSlices - e));
        int8_t dig[kSlices];
#pragma unroll
        for (int t = kSlices - 1; t >= 1; --t) {
            long long r = ((Q + 128) & 255) - 128;  // balanced digit in [-128, 127]
            dig[t] = (int8_t)r;
            Q = (Q - r) >> 8;
        }
        dig[0] = (int8_t)Q;  // |Q| <= 65 here
#pragma unroll
        for (int t = 0; t < kSlices; ++t) planes[(size_t)t * plane + (size_t)i * ldk + j] = dig[t];
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

// P: [g][kSlices*npad] int32 (column-major (kSlices*npad) x mb as cuBLAS writes it).  xr[g*n + i] fp64.
__global__ void __launch_bounds__(256) combine_i8_kernel(const int32_t* __restrict__ P, const int* __restrict__ exps,
                                                          int n, int npad, long long mb, double* __restrict__ xr, long long ldx)
{
    const long long g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || g >= mb) return;
    const int32_t* p = P + (size_t)g * kSlices * npad + i;
    // sum_t P_t 256^(7-t): planes 0..2 (<= 2^46) and planes 3..6 (<= 2^54) separately, then one fp64 sum
    long long hi = 0, lo = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) hi = hi * 256 + (long long)p[(size_t)t * npad];
#pragma unroll
    for (int t = 3; t < kSlices; ++t) lo = lo * 256 + (long long)p[(size_t)t * npad];
    const double v = (double)hi + ldexp((double)lo, -32);
    xr[(size_t)g * ldx + i] = ldexp(v, exps[i] - 24);
}

}  // namespace pg
