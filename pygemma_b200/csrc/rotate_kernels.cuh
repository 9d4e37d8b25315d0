// rotate_kernels.cuh -- stage 1 of the scan: Xr = U^T X for one SNP block, emitted SNP-major in fp64.
//
// Replaces `X = U.T @ X` (reference lmm/lmm.py:244, an fp32 sgemm over the whole matrix) with a blocked
// device GEMM whose output layout is what the REML kernel streams (SNP g at xr + g*n).
//
// Engines
//   PG_ROT_FP64     FP64 GEMM on the staged fp64 block: any genotype dtype.
//   PG_ROT_I8SPLIT  exact integer-split path for int8 dosages (rotate_i8.cuh): eight int8 tensor-core
//                   GEMMs against the digit planes of U^T, recombined exactly.
//   PG_ROT_AUTO     int8 genotypes -> I8SPLIT, everything else -> FP64.
#pragma once

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/pygemma_b200.h"
#include "reml_kernels.cuh"
#include "rotate_i8.cuh"
#include "rotate_i8_tc.cuh"
#include "rotate_i8_tc2.cuh"
#include "rotate_i8_tc4.cuh"

namespace pg {

struct RotWorkspace {
    std::string err;
    // int8-split state
    bool planes_valid = false;
    int n = 0, npad = 0, ldk = 0;
    int8_t* planes = nullptr;  // [npad][kSlices][ldk]: the digit planes of an eigenvector are adjacent rows
    int* exps = nullptr;       // [n]
    int8_t* x8 = nullptr;      // [cap_snps][ldk]
    long long cap_snps = 0;
    int32_t* P[2] = {nullptr, nullptr};  // [(kSlices*npad) x sub] column-major, double-buffered
    cudaEvent_t ev_gemm[2] = {nullptr, nullptr}, ev_pfree[2] = {nullptr, nullptr};
    unsigned pcount = 0;
    double* scale = nullptr;   // [n] 2^(e_i - 24) for the fused tcgen05 engine
    // level coding of float genotypes (rotate_i8.cuh)
    double* u1 = nullptr;      // [n] U^T 1
    LevelInfo* info = nullptr; // [cap_snps]
    LevelPartial* part = nullptr;  // [chunks][cap_snps]
    int8_t* codes = nullptr;   // [cap_snps * n] codes in the layout of the input block
    int* n_bad = nullptr;      // device counter
    long long code_cap = 0;
    long long blocks_coded = 0, blocks_dense = 0;
    int bed_count_a1 = 0, bed_standardize = 0;   // PG_X_BED decoding options (pg_set_bed_options)
    long long sub = 0, want_cap = 0;
    float slice_ms = 0.f;
};

inline void rot_free(RotWorkspace* w)
{
    if (w->planes) cudaFree(w->planes);
    if (w->exps) cudaFree(w->exps);
    if (w->scale) cudaFree(w->scale);
    w->scale = nullptr;
    if (w->u1) cudaFree(w->u1);
    if (w->info) cudaFree(w->info);
    if (w->part) cudaFree(w->part);
    if (w->codes) cudaFree(w->codes);
    if (w->n_bad) cudaFree(w->n_bad);
    w->u1 = nullptr; w->info = nullptr; w->part = nullptr; w->codes = nullptr; w->n_bad = nullptr; w->code_cap = 0;
    if (w->x8) cudaFree(w->x8);
    for (int t = 0; t < 2; ++t) {
        if (w->P[t]) cudaFree(w->P[t]);
        w->P[t] = nullptr;
    }
    for (int t = 0; t < 2; ++t) {
        if (w->ev_gemm[t]) cudaEventDestroy(w->ev_gemm[t]);
        if (w->ev_pfree[t]) cudaEventDestroy(w->ev_pfree[t]);
        w->ev_gemm[t] = w->ev_pfree[t] = nullptr;
    }
    w->planes = nullptr; w->exps = nullptr; w->x8 = nullptr;
    w->planes_valid = false; w->cap_snps = 0; w->sub = 0;
}
inline void rot_invalidate(RotWorkspace* w) { w->planes_valid = false; }

// int32 partial products of the cuBLAS split path: kSlices*npad x sub, two buffers of about 1 GiB (allocated on first
// use: the fused tcgen05 engine never materialises them)
inline int rot_ensure_P(RotWorkspace* w)
{
    static const double p_gib = getenv("PG_P_GIB") ? atof(getenv("PG_P_GIB")) : 1.0;
    long long sub = (long long)((size_t)(p_gib * (double)(size_t(1) << 30)) / ((size_t)kSlices * w->npad * 4));
    sub = std::max<long long>(256, std::min<long long>((sub / 256) * 256, (w->want_cap + 255) / 256 * 256));
    if (sub > w->sub) {
        for (int t = 0; t < 2; ++t) {
            if (w->P[t]) cudaFree(w->P[t]);
            w->P[t] = nullptr;
            if (cudaMalloc(&w->P[t], (size_t)kSlices * w->npad * sub * sizeof(int32_t)) != cudaSuccess) {
                w->err = "cudaMalloc of the int32 partial-product buffer failed";
                return PG_ERR_CUDA;
            }
        }
        w->sub = sub;
    }
    return 0;
}
inline const char* rot_error(RotWorkspace* w) { return w->err.c_str(); }

inline int stage_to_snp_major(cudaStream_t stream, int n, const void* src, int xdtype, long long ld, int layout,
                              long long mb, double* dst, long long ldd, const int* perm = nullptr)
{
    dim3 block(32, 8), grid((unsigned)((mb + 31) / 32), (unsigned)((n + 31) / 32));
    switch (xdtype) {
    case PG_X_I8:
        to_snp_major_kernel<int8_t><<<grid, block, 0, stream>>>((const int8_t*)src, ld, layout, n, mb, dst, ldd, perm);
        break;
    case PG_X_F32:
        to_snp_major_kernel<float><<<grid, block, 0, stream>>>((const float*)src, ld, layout, n, mb, dst, ldd, perm);
        break;
    default:
        to_snp_major_kernel<double><<<grid, block, 0, stream>>>((const double*)src, ld, layout, n, mb, dst, ldd, perm);
        break;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : PG_ERR_CUDA;
}

#define PG_ROT_CK(call)                                                                  \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            w->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return PG_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

inline int rot_prepare_i8(RotWorkspace* w, cudaStream_t stream, const double* U, int u_op_t, int n, long long blk)
{
    // rows of the K-contiguous operands start on 128-byte lines (TMA boxes and GEMM tiles then touch whole sectors)
    const int npad = (n + 15) / 16 * 16, ldk = (n + 127) / 128 * 128;
    if (w->n != n || !w->planes) {
        rot_free(w);
        w->n = n; w->npad = npad; w->ldk = ldk;
        PG_ROT_CK(cudaMalloc(&w->planes, (size_t)kSlices * npad * ldk));
        PG_ROT_CK(cudaMalloc(&w->exps, sizeof(int) * n));
        PG_ROT_CK(cudaMalloc(&w->scale, sizeof(double) * n));
        PG_ROT_CK(cudaMalloc(&w->u1, sizeof(double) * n));
        PG_ROT_CK(cudaMalloc(&w->n_bad, 2 * sizeof(int)));
    }
    const long long cap = (blk + 63) / 64 * 64;
    if (cap > w->cap_snps) {
        if (w->x8) cudaFree(w->x8);
        w->x8 = nullptr;
        PG_ROT_CK(cudaMalloc(&w->x8, (size_t)cap * ldk));
        PG_ROT_CK(cudaMemsetAsync(w->x8, 0, (size_t)cap * ldk, stream));
        w->cap_snps = cap;
    }
    w->want_cap = cap;
    for (int t = 0; t < 2; ++t) {
        if (!w->ev_gemm[t]) PG_ROT_CK(cudaEventCreateWithFlags(&w->ev_gemm[t], cudaEventDisableTiming));
        if (!w->ev_pfree[t]) PG_ROT_CK(cudaEventCreateWithFlags(&w->ev_pfree[t], cudaEventDisableTiming));
    }
    if (!w->planes_valid) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
        slice_u_kernel<<<npad, 256, 0, stream>>>(U, u_op_t ? 1 : 0, n, npad, ldk, w->planes, w->exps, n);
        PG_ROT_CK(cudaGetLastError());
        tc::plane_scale_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w->exps, n, w->scale);
        PG_ROT_CK(cudaGetLastError());
        column_sums_kernel<<<n, 256, 0, stream>>>(U, u_op_t ? 1 : 0, n, w->u1);
        PG_ROT_CK(cudaGetLastError());
        cudaEventRecord(e1, stream);
        PG_ROT_CK(cudaStreamSynchronize(stream));
        cudaEventElapsedTime(&w->slice_ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        w->planes_valid = true;
    }
    return 0;
}

// Moments straight from the fused rotation (rotate_i8_tc2.cuh, FUSE): what the handle prepared for the current design.
// args.Z is the block's moment buffer; planes_g shares the row pitch ldk of the eigenvector planes.
struct FuseLaunch {
    tc2::FuseArgs args;
    const int8_t* planes_g = nullptr;     // [32 g_tiles][kSlices][ldk]
    const tc2::SegRed* segs = nullptr;    // x^2 reduction: one entry per COMPRESS segment
    int nsegs = 0;
};

// Rotates one block.  xf: staging buffer (mb x n fp64), xr: output (mb x n fp64, SNP-major).
// `stream` carries staging and the GEMMs, `cmb` the int8 recombination kernels; the block's rotated vectors are
// complete when `ev_rot_end` (recorded on the stream that wrote them last) has fired.
inline int rot_run(RotWorkspace* w, cublasHandle_t blas, cudaStream_t stream, cudaStream_t cmb, int rotation,
                   const double* U, int u_op_t, int n, const void* src, int xdtype, long long ld, int layout, long long mb,
                   long long blk, double* xf, double* xr, long long ldx, int* used_i8, int* n_launch,
                   cudaEvent_t ev_conv_end, cudaEvent_t ev_rot_begin, cudaEvent_t ev_rot_end,
                   const FuseLaunch* fuse = nullptr, int* moments_done = nullptr)
{
    // fuse: the fused tcgen05 engine may write the block's compressed moments (fuse->args.Z) instead of the rotated
    // genotypes; *moments_done says whether it did (it does not for blocks that need the second, eps-weighted pass or
    // that fall back to another engine: the caller then runs the compression on xr as usual)
    if (moments_done) *moments_done = 0;
    const bool i8 = (xdtype == PG_X_I8) && (rotation == PG_ROT_AUTO || rotation == PG_ROT_I8SPLIT || rotation == PG_ROT_I8TC);
    const bool forced_i8 = (rotation == PG_ROT_I8SPLIT || rotation == PG_ROT_I8TC);
    *used_i8 = i8 ? PG_ROT_I8SPLIT : 0;   // engine actually used (refined below)
    // float genotypes whose columns are (affine images of) dosage codes go through the int8 path on the codes
    static const bool level_coding = !(getenv("PG_LEVEL_CODING") && atoi(getenv("PG_LEVEL_CODING")) == 0);
    const LevelInfo* affine = nullptr;
    bool need_eps = false;
    const void* fsrc = nullptr;   // the float block (for the indicator pass)
    long long fld = 0;
    int fdtype = 0;
    // level codes (indicator = 0) or indicators (1) of the float block -> w->codes, in the layout of the input
    auto encode_block = [&](int indicator) -> int {
        if (fdtype == PG_X_BED) {   // the indicator of a .bed block goes straight into x8
            bed_decode_kernel<<<(unsigned)mb, 128, 0, stream>>>((const uint8_t*)fsrc, fld, n, mb, w->ldk, w->bed_count_a1,
                                                                w->bed_standardize, indicator, w->x8, w->info, w->n_bad);
            PG_ROT_CK(cudaGetLastError());
            (*n_launch)++;
            return 0;
        }
        dim3 ge((unsigned)((mb + 127) / 128), (unsigned)((n + 63) / 64));
        if (fdtype == PG_X_F32)
            encode_levels_kernel<float><<<ge, 128, 0, stream>>>((const float*)fsrc, fld, layout, n, mb, w->info, w->codes, indicator);
        else
            encode_levels_kernel<double><<<ge, 128, 0, stream>>>((const double*)fsrc, fld, layout, n, mb, w->info, w->codes, indicator);
        PG_ROT_CK(cudaGetLastError());
        (*n_launch)++;
        return 0;
    };
    // int8 block (src, ld, layout) -> SNP-major K-contiguous copy x8
    auto stage_block = [&]() -> int {
        dim3 block(64, 4), grid((unsigned)((mb + 63) / 64), (unsigned)((w->ldk + 63) / 64));
        stage_i8_kernel<<<grid, block, 0, stream>>>((const int8_t*)src, ld, layout, n, mb, w->ldk, w->x8);
        PG_ROT_CK(cudaGetLastError());
        (*n_launch)++;
        return 0;
    };
    bool pre_staged = false;   // x8 already holds the K-contiguous int8 operand
    const bool bed = (xdtype == PG_X_BED);
    if (bed) {
        // packed PLINK block -> dosage codes in x8 + the affine description of every (imputed, standardised) column
        if (rotation == PG_ROT_FP64) { w->err = "PLINK .bed input runs on the int8 engines only"; return PG_ERR_ARG; }
        int rc = rot_prepare_i8(w, stream, U, u_op_t, n, blk);
        if (rc) return rc;
        const long long cap = (blk + 63) / 64 * 64;
        if (cap > w->code_cap) {   // only `info` is needed; keep the three buffers in step
            if (w->info) cudaFree(w->info);
            if (w->part) cudaFree(w->part);
            if (w->codes) cudaFree(w->codes);
            w->info = nullptr; w->part = nullptr; w->codes = nullptr; w->code_cap = 0;
            PG_ROT_CK(cudaMalloc(&w->info, sizeof(LevelInfo) * cap));
            PG_ROT_CK(cudaMalloc(&w->part, sizeof(LevelPartial) * cap * ((n + kLevelChunk - 1) / kLevelChunk)));
            PG_ROT_CK(cudaMalloc(&w->codes, (size_t)cap * n));
            w->code_cap = cap;
        }
        PG_ROT_CK(cudaMemsetAsync(w->n_bad, 0, 2 * sizeof(int), stream));
        bed_decode_kernel<<<(unsigned)mb, 128, 0, stream>>>((const uint8_t*)src, ld, n, mb, w->ldk, w->bed_count_a1,
                                                            w->bed_standardize, 0, w->x8, w->info, w->n_bad);
        PG_ROT_CK(cudaGetLastError());
        (*n_launch)++;
        int bad[2] = {0, 0};
        PG_ROT_CK(cudaMemcpyAsync(bad, w->n_bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        PG_ROT_CK(cudaStreamSynchronize(stream));
        affine = w->info;
        need_eps = bad[1] > 0;
        fsrc = src; fld = ld; fdtype = PG_X_BED;
        pre_staged = true;
        *used_i8 = PG_ROT_I8SPLIT;
    }
    if (!bed && !i8 && rotation != PG_ROT_FP64 && level_coding && xdtype != PG_X_I8) {
        int rc = rot_prepare_i8(w, stream, U, u_op_t, n, blk);
        if (rc) return rc;
        const long long cap = (blk + 63) / 64 * 64;
        if (cap > w->code_cap) {
            if (w->info) cudaFree(w->info);
            if (w->part) cudaFree(w->part);
            if (w->codes) cudaFree(w->codes);
            w->info = nullptr; w->part = nullptr; w->codes = nullptr; w->code_cap = 0;
            PG_ROT_CK(cudaMalloc(&w->info, sizeof(LevelInfo) * cap));
            PG_ROT_CK(cudaMalloc(&w->part, sizeof(LevelPartial) * cap * ((n + kLevelChunk - 1) / kLevelChunk)));
            PG_ROT_CK(cudaMalloc(&w->codes, (size_t)cap * n));
            w->code_cap = cap;
        }
        PG_ROT_CK(cudaMemsetAsync(w->n_bad, 0, 2 * sizeof(int), stream));
        const unsigned gb = (unsigned)((mb + 127) / 128);
        const int nchunks = (n + kLevelChunk - 1) / kLevelChunk;
        // "equally spaced" means to double rounding; anything looser goes to the indicator component (treating
        // float32-standardised columns as affine moved beta by 1.4e-6 in the parity test)
        const double tol = ldexp(1.0, -50);
        if (xdtype == PG_X_F32)
            find_levels_kernel<float><<<dim3(gb, nchunks), 128, 0, stream>>>((const float*)src, ld, layout, n, mb, w->part);
        else
            find_levels_kernel<double><<<dim3(gb, nchunks), 128, 0, stream>>>((const double*)src, ld, layout, n, mb, w->part);
        PG_ROT_CK(cudaGetLastError());
        merge_levels_kernel<<<gb, 128, 0, stream>>>(w->part, nchunks, mb, tol, w->info, w->n_bad);
        PG_ROT_CK(cudaGetLastError());
        (*n_launch) += 2;
        int bad[2] = {1, 0};
        PG_ROT_CK(cudaMemcpyAsync(bad, w->n_bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        PG_ROT_CK(cudaStreamSynchronize(stream));  // one small sync per block: the path choice is made on the host
        if (bad[0] == 0) {
            affine = w->info;
            need_eps = bad[1] > 0;
            fsrc = src; fld = ld; fdtype = xdtype;
            int er = encode_block(0);
            if (er) return er;
            src = w->codes;
            ld = (layout == PG_X_SAMPLE_MAJOR) ? mb : n;
            xdtype = PG_X_I8;
            *used_i8 = PG_ROT_I8SPLIT;
            w->blocks_coded++;
        } else {
            w->blocks_dense++;
        }
    }
    if (!i8 && !affine && forced_i8) {
        w->err = "PG_ROT_I8SPLIT / PG_ROT_I8TC need int8 genotypes or float genotypes with at most three levels per SNP";
        return PG_ERR_ARG;
    }
    if (!i8 && !affine) {
        int rc = stage_to_snp_major(stream, n, src, xdtype, ld, layout, mb, xf, n);
        if (rc) { w->err = "staging kernel launch failed"; return rc; }
        (*n_launch)++;
        cudaEventRecord(ev_conv_end, stream);
        cudaEventRecord(ev_rot_begin, stream);
        const double one = 1.0, zero = 0.0;
        // column-major view: Xr (n x mb, ld n) = op(U) (n x n) * Xf (n x mb, ld n)
        cublasStatus_t s = cublasDgemm(blas, u_op_t ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, n, (int)mb, n, &one, U, n,
                                       xf, n, &zero, xr, (int)ldx);
        if (s != CUBLAS_STATUS_SUCCESS) { w->err = "cublasDgemm failed, status " + std::to_string((int)s); return PG_ERR_CUBLAS; }
        cudaEventRecord(ev_rot_end, stream);
        return 0;
    }
    int rc = rot_prepare_i8(w, stream, U, u_op_t, n, blk);
    if (rc) return rc;
    // A sample-major int8 block is handed to the GEMM as a transposed (MN-major) operand: no staging kernel, no x8 copy,
    // and the cuBLAS kernel picked for this layout is ~17 % faster here (10.7 vs 12.8 ms per 25 088 SNPs at n = 10 000).
    // PG_GEMM_TT=0 forces the staged K-major path (also used for SNP-major input, ragged tails and n % 16 != 0).
    static const bool gemm_tt = !(getenv("PG_GEMM_TT") && atoi(getenv("PG_GEMM_TT")) == 0);
    // PG_ROT_AUTO on int8 dosages: the hand-written fused tcgen05 kernel (PG_ROT_DEFAULT=cublas selects the library GEMM
    // + recombination pair, which level-coded float blocks always use: their affine fix-up lives in the recombination)
    static const bool default_tc = !(getenv("PG_ROT_DEFAULT") && !strcmp(getenv("PG_ROT_DEFAULT"), "cublas"));
    const bool fused_tc = (rotation == PG_ROT_I8TC || (rotation == PG_ROT_AUTO && default_tc));
    (void)forced_i8;
    const bool direct = !pre_staged && gemm_tt && layout == PG_X_SAMPLE_MAJOR && (ld % 16 == 0) && (((uintptr_t)src) % 16 == 0) &&
                        (n % 16 == 0) && (mb % 16 == 0);
    if (!direct && !pre_staged) {
        rc = stage_block();
        if (rc) return rc;
    }
    cudaEventRecord(ev_conv_end, stream);
    cudaEventRecord(ev_rot_begin, stream);
    if (fused_tc) {
        *used_i8 = PG_ROT_I8TC;
        // one kernel: TMA -> tcgen05.mma kind::i8 (7 planes in TMEM) -> exact recombination in the epilogue
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // CTA-pair kernel (cta_group::2), or two pairs per cluster sharing the genotype tiles by TMA multicast (PG_TC_CLUSTER=4)
        static const int tc_cluster = getenv("PG_TC_CLUSTER") ? atoi(getenv("PG_TC_CLUSTER")) : 2;
        const bool fuse_now = fuse && tc_cluster == 2 && !(affine && need_eps);
        auto tc_launch = [&](int accumulate) -> int {
            if (fuse_now)
                return tc2::launch(stream, sms, w->x8, w->cap_snps, w->planes, w->npad, w->ldk, n, mb, w->scale, xr, ldx,
                                   direct ? (const int8_t*)src : nullptr, ld, affine, w->u1, 0, &fuse->args, fuse->planes_g);
            if (tc_cluster == 4)
                return tc4::launch(stream, sms, w->x8, w->cap_snps, w->planes, w->npad, w->ldk, n, mb, w->scale, xr, ldx,
                                   direct ? (const int8_t*)src : nullptr, ld, affine, w->u1, accumulate);
            int launched = 1;
            const int rl = tc2::launch(stream, sms, w->x8, w->cap_snps, w->planes, w->npad, w->ldk, n, mb, w->scale, xr, ldx,
                                       direct ? (const int8_t*)src : nullptr, ld, affine, w->u1, accumulate, nullptr, nullptr,
                                       &launched);
            (*n_launch) += launched - 1;   // one launch per eigen-tile group when the planes are pinned in the L2
            return rl;
        };
        int r = tc_launch(0);
        if (r == 0 && affine && need_eps) {
            // second component (unequally spaced levels / an outlier level): the indicator, accumulated with weight eps
            rc = encode_block(1);
            if (rc == 0 && !direct && !pre_staged) rc = stage_block();
            if (rc) return rc;
            r = tc_launch(1);
            (*n_launch)++;
        }
        if (r) { w->err = "fused tcgen05 rotation launch failed, code " + std::to_string(r); return PG_ERR_CUDA; }
        (*n_launch)++;
        if (fuse_now) {
            if (fuse->nsegs) {
                tc2::moments_reduce_kernel<<<dim3((unsigned)((mb + 127) / 128), (unsigned)fuse->nsegs), 128, 0, stream>>>(
                    fuse->args.P2, fuse->args.ldp, mb, fuse->segs, fuse->args.Z, fuse->args.ldz, fuse->args.x2row);
                PG_ROT_CK(cudaGetLastError());
                (*n_launch)++;
            }
            if (moments_done) *moments_done = 1;
        }
        cudaEventRecord(ev_rot_end, stream);
        return 0;
    }
    rc = rot_ensure_P(w);
    if (rc) return rc;
    const int32_t ione = 1, izero = 0;
    const int M = kSlices * w->npad;
    const int npass = (affine && need_eps) ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        if (pass == 1) {
            // second component (unequally spaced levels / an outlier level): the indicator, weight eps; the GEMMs of
            // pass 0 that read the codes precede this on the same stream
            rc = encode_block(1);
            if (rc == 0 && !direct && !pre_staged) rc = stage_block();
            if (rc) return rc;
        }
    for (long long g0 = 0; g0 < mb; g0 += w->sub) {
        const long long cnt = std::min(w->sub, mb - g0);
        const long long cnt_pad = (cnt + 15) / 16 * 16;  // x8 rows beyond mb are zero / stale: ignored downstream
        const int t = (int)(w->pcount++ & 1);
        PG_ROT_CK(cudaStreamWaitEvent(stream, w->ev_pfree[t], 0));  // the recombination that last read P[t] is done
        cublasStatus_t s;
        if (direct)
            // B = X block as an (cnt x n) column-major matrix with leading dimension ld (sample-major storage), transposed
            s = cublasGemmEx(blas, CUBLAS_OP_T, CUBLAS_OP_T, M, (int)cnt, n, &ione, w->planes, CUDA_R_8I, w->ldk,
                             (const int8_t*)src + g0, CUDA_R_8I, (int)ld, &izero, w->P[t], CUDA_R_32I, M, CUBLAS_COMPUTE_32I,
                             CUBLAS_GEMM_DEFAULT);
        else
            s = cublasGemmEx(blas, CUBLAS_OP_T, CUBLAS_OP_N, M, (int)cnt_pad, w->ldk, &ione, w->planes,
                             CUDA_R_8I, w->ldk, w->x8 + (size_t)g0 * w->ldk, CUDA_R_8I, w->ldk, &izero, w->P[t],
                             CUDA_R_32I, M, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
        if (s != CUBLAS_STATUS_SUCCESS) { w->err = "cublasGemmEx(int8) failed, status " + std::to_string((int)s); return PG_ERR_CUBLAS; }
        PG_ROT_CK(cudaEventRecord(w->ev_gemm[t], stream));
        PG_ROT_CK(cudaStreamWaitEvent(cmb, w->ev_gemm[t], 0));
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)cnt);
        if (affine)
            combine_i8_affine_kernel<<<grid, 256, 0, cmb>>>(w->P[t], w->exps, n, w->npad, cnt, xr + (size_t)g0 * ldx, ldx,
                                                            affine + g0, w->u1, pass);
        else
            combine_i8_kernel<<<grid, 256, 0, cmb>>>(w->P[t], w->exps, n, w->npad, cnt, xr + (size_t)g0 * ldx, ldx);
        PG_ROT_CK(cudaGetLastError());
        PG_ROT_CK(cudaEventRecord(w->ev_pfree[t], cmb));
        (*n_launch)++;
    }
    }
    cudaEventRecord(ev_rot_end, cmb);
    return 0;
}

}  // namespace pg
