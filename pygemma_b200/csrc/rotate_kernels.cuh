// rotate_kernels.cuh -- stage 1 of the scan: Xr = U^T X for one SNP block, emitted SNP-major in fp64.
//
// Replaces `X = U.T @ X` (reference lmm/lmm.py:244, an fp32 sgemm over the whole matrix) with a blocked
// device GEMM whose output layout is what the REML kernel streams (SNP g at xr + g*n).
//
// Engines
//   PG_ROT_FP64     FP64 GEMM (cuBLAS DGEMM on the staged fp64 block): any genotype dtype.
//   PG_ROT_I8SPLIT  exact integer path for int8 dosages (see rotate_i8.cuh when built in).
#pragma once

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <string>

#include "../../include/pygemma_b200.h"
#include "reml_kernels.cuh"

namespace pg {

struct RotWorkspace {
    std::string err;
    bool valid = false;
};

inline void rot_free(RotWorkspace* w) { w->valid = false; }
inline void rot_invalidate(RotWorkspace* w) { w->valid = false; }
inline const char* rot_error(RotWorkspace* w) { return w->err.c_str(); }

inline int stage_to_snp_major(cudaStream_t stream, int n, const void* src, int xdtype, long long ld, int layout,
                              long long mb, double* dst)
{
    dim3 block(32, 8), grid((unsigned)((mb + 31) / 32), (unsigned)((n + 31) / 32));
    switch (xdtype) {
    case PG_X_I8:
        to_snp_major_kernel<int8_t><<<grid, block, 0, stream>>>((const int8_t*)src, ld, layout, n, mb, dst);
        break;
    case PG_X_F32:
        to_snp_major_kernel<float><<<grid, block, 0, stream>>>((const float*)src, ld, layout, n, mb, dst);
        break;
    default:
        to_snp_major_kernel<double><<<grid, block, 0, stream>>>((const double*)src, ld, layout, n, mb, dst);
        break;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : PG_ERR_CUDA;
}

// Rotates one block.  xf: staging buffer (mb x n fp64), xr: output (mb x n fp64, SNP-major).
inline int rot_run(RotWorkspace* w, cublasHandle_t blas, cudaStream_t stream, int rotation, const double* U, int u_op_t,
                   int n, const void* src, int xdtype, long long ld, int layout, long long mb, double* xf, double* xr,
                   bool* staged, int* used_i8, int* n_launch, cudaEvent_t ev_conv_end, cudaEvent_t ev_rot_begin,
                   cudaEvent_t ev_rot_end, int sm_count)
{
    (void)rotation; (void)sm_count;
    *used_i8 = 0;
    int rc = stage_to_snp_major(stream, n, src, xdtype, ld, layout, mb, xf);
    if (rc) { w->err = "staging kernel launch failed"; return rc; }
    *staged = true;
    cudaEventRecord(ev_conv_end, stream);
    cudaEventRecord(ev_rot_begin, stream);
    const double one = 1.0, zero = 0.0;
    // column-major view: Xr (n x mb, ld n) = op(U) (n x n) * Xf (n x mb, ld n)
    cublasStatus_t s = cublasDgemm(blas, u_op_t ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, n, (int)mb, n, &one, U, n, xf,
                                   n, &zero, xr, n);
    if (s != CUBLAS_STATUS_SUCCESS) { w->err = "cublasDgemm failed, status " + std::to_string((int)s); return PG_ERR_CUBLAS; }
    (*n_launch)++;
    cudaEventRecord(ev_rot_end, stream);
    return 0;
}

}  // namespace pg
