// rotate_i8_tc.cuh -- the exact int8-split rotation as ONE hand-written sm_100a kernel: TMA-staged operands,
// tcgen05.mma kind::i8 with the accumulators of all seven digit planes in tensor memory, and the exact
// recombination (int32 planes -> fp64, rounded once) fused into the epilogue.  Replaces the cuBLAS int8 GEMM +
// combine_i8_kernel pair of rotate_kernels.cuh: the int32 partial products (280 KB per SNP at n = 10 000) never
// touch HBM.  Reference: `X = U.T @ X`, lmm/lmm.py:244.
//
// Tile: 256 SNPs x 32 eigenvectors x 7 planes.  Per k-step (32 samples) two MMAs of shape M = 128 (SNPs) x
// N = 224 (plane-major: column p*32 + e) x K = 32 share the same B operand; accumulators live in TMEM columns
// [0, 224) and [256, 480).  Shared-memory stage = 128 samples: A0, A1 (128 SNP rows x 128 B each) and B (224 rows x
// 128 B), all K-major with the 128-byte swizzle TMA writes and the UMMA descriptor reads.  Arithmetic intensity
// 122 MAC per L2 byte, the same regime as cuBLAS's 2-SM 256x256 tile.
//
// Warp roles (192 threads, persistent CTAs, static tile schedule):
//   warp 0    TMA producer (one elected lane): waits `empty[s]`, issues 2 + 1 bulk tensor loads onto `full[s]`
//   warp 1    TMEM allocator + MMA issuer (one elected lane): waits `full[s]`, issues 8 tcgen05.mma, commits to `empty[s]`;
//             after the last k-step commits to `tmem_full`; waits `tmem_empty` before overwriting the accumulators
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 bit x 8 columns per plane), exact recombination in fp64 FMAs,
//             64-byte stores of 8 rotated values per SNP row; arrive on `tmem_empty`
// Every mbarrier wait carries a watchdog that traps instead of hanging the GPU.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "rotate_i8.cuh"

namespace pg {
namespace tc {

constexpr int kTileSnps = 256;                 // two M = 128 accumulators
constexpr int kTileEig = 32;                   // eigenvectors per tile
constexpr int kTileN = kSlices * kTileEig;     // 224 accumulator columns
#ifndef PG_TC_STAGE_K
#define PG_TC_STAGE_K 128
#endif
constexpr int kStageK = PG_TC_STAGE_K;         // samples (bytes) per shared-memory stage = one swizzle atom (64 or 128)
constexpr int kStages = (kStageK == 128) ? 3 : 7;   // 180 / 210 KB of operands in flight
constexpr int kABytes = 128 * kStageK;         // one 128-row A tile
constexpr int kBBytes = kTileN * kStageK;
constexpr int kStageBytes = 2 * kABytes + kBBytes;   // 61 440
constexpr int kTmemCols = 512;
constexpr int kAcc1Col = 256;
constexpr int kEigGroup = 24;                  // eigen tiles swept together (L2 residency of the B panels)
constexpr int kThreads = 192;
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;

// instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, S8 x S8, both K-major, N = 224, M = 128
constexpr uint32_t kInstrDesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (unsigned long long spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        if (spin > (1ull << 26)) __trap();  // watchdog: a pipeline bug must abort the kernel, not hang the GPU
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major, swizzle atom = one stage row (128 or 64 B): 8-row groups 8*kStageK B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                  // leading byte offset (ignored for swizzled K-major), 16 B units
    d |= (uint64_t)((8 * kStageK) >> 4) << 32;   // stride byte offset: 8 rows of one swizzle atom
    d |= (uint64_t)1 << 46;                  // descriptor version
    d |= (uint64_t)(kStageK == 128 ? 2 : 4) << 61;   // SWIZZLE_128B / SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

struct Args {
    long long mb;          // SNPs in the block
    int n;                 // eigenvectors / samples
    int snp_tiles, eig_tiles;
    const double* scale;   // [n] 2^(e_i - 24)
    double* xr;            // [mb][ldx]
    long long ldx;
};

__global__ void __launch_bounds__(kThreads, 1)
rotate_i8_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_p, Args a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)kStages * kStageBytes);
    uint64_t* full = bars;                   // [kStages]
    uint64_t* empty = bars + kStages;        // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total_tiles = (long long)a.snp_tiles * a.eig_tiles;
    const int ksteps = (a.n + kStageK - 1) / kStageK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    // tile t -> (snp tile, eigen tile): eigen tiles in groups of kEigGroup, SNP tiles swept inside a group
    auto decode = [&](long long t, int& st, int& et) {
        const long long per_group = (long long)kEigGroup * a.snp_tiles;
        const int g = (int)(t / per_group);
        const long long r = t - (long long)g * per_group;
        const int e0 = g * kEigGroup;
        const int ecount = min(kEigGroup, a.eig_tiles - e0);
        st = (int)(r / ecount);
        et = e0 + (int)(r % ecount);
    };
    // the last eigen group may hold fewer than kEigGroup tiles: recompute the tile count it really has
    const int full_groups = a.eig_tiles / kEigGroup, tail = a.eig_tiles - full_groups * kEigGroup;
    (void)tail;

    if (warp == 0) {
        // whole warp walks the schedule (keeps the warp convergent for the final barrier); lane 0 issues
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            int st, et;
            decode(t, st, et);
            for (int k = 0; k < ksteps; ++k) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sA = smem + (size_t)stage * kStageBytes;
                    mbar_expect_tx(&full[stage], kStageBytes);
                    tma_load_2d(sA, &map_x, &full[stage], k * kStageK, st * kTileSnps);
                    tma_load_2d(sA + kABytes, &map_x, &full[stage], k * kStageK, st * kTileSnps + 128);
                    tma_load_3d(sA + 2 * kABytes, &map_p, &full[stage], k * kStageK, et * kTileEig, 0);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            mbar_wait(tmem_empty, acc_phase ^ 1);  // epilogue has drained the accumulators of the previous tile
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int k = 0; k < ksteps; ++k) {
                mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t sA = smem_u32(smem + (size_t)stage * kStageBytes);
                    const uint64_t da0 = umma_desc_sw128(sA), da1 = umma_desc_sw128(sA + kABytes);
                    const uint64_t db = umma_desc_sw128(sA + 2 * kABytes);
#pragma unroll
                    for (int kk = 0; kk < kStageK / 32; ++kk) {
                        const uint32_t accum = (k | kk) ? 1u : 0u;
                        const uint64_t adv = (uint64_t)((kk * 32) >> 4);  // 32 bytes along K inside the swizzle atom
                        umma_i8(tmem_base, da0 + adv, db + adv, accum);
                        umma_i8(tmem_base + kAcc1Col, da1 + adv, db + adv, accum);
                    }
                    umma_commit(&empty[stage]);  // frees the stage once the MMAs above have read it
                    if (k == ksteps - 1) umma_commit(tmem_full);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            acc_phase ^= 1;
        }
    } else {
        // epilogue warps 2..5 own TMEM lanes 32*(warp%4) .. +31
        const int quarter = warp & 3;
        uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            int st, et;
            decode(t, st, et);
            mbar_wait(tmem_full, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int eig0 = et * kTileEig;
#pragma unroll 1
            for (int acc = 0; acc < 2; ++acc) {
                const long long snp = (long long)st * kTileSnps + acc * 128 + quarter * 32 + lane;
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAcc1Col);
#pragma unroll 1
                for (int c = 0; c < kTileEig / 8; ++c) {
                    uint32_t r[kSlices][8];
#pragma unroll
                    for (int p = 0; p < kSlices; ++p) tmem_ld8(tbase + (uint32_t)(p * kTileEig + c * 8), r[p]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (snp < a.mb) {
                        double* dst = a.xr + (size_t)snp * a.ldx + eig0 + c * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int e = eig0 + c * 8 + j;
                            if (e < a.n) {
                                // sum_p P_p 256^(kSlices-1-p): planes 0..2 and 3.. are exact in fp64, one rounding in the final sum
                                const double hi = fma((double)(int)r[0][j], 65536.0, fma((double)(int)r[1][j], 256.0, (double)(int)r[2][j]));
                                double lo = (double)(int)r[3][j];
#pragma unroll
                                for (int p = 4; p < kSlices; ++p) lo = fma(lo, 256.0, (double)(int)r[p][j]);
                                const double v = fma(lo, kLoScale, hi);
                                dst[j] = v * __ldg(a.scale + e);
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tmem_empty);
            acc_phase ^= 1;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// 2^(e_i - 24): exact power-of-two scale of eigenvector i's fixed-point representation
__global__ void plane_scale_kernel(const int* __restrict__ exps, int n, double* __restrict__ scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scale[i] = ldexp(1.0, exps[i] - 24);
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// x8: [rows][ldk] int8 (SNP-major, K contiguous); planes: [kSlices][npad][ldk] int8
inline int launch(cudaStream_t stream, int sm_count, const int8_t* x8, long long x8_rows, const int8_t* planes, int npad,
                  int ldk, int n, long long mb, const double* scale, double* xr, long long ldx)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return -1;
    CUtensorMap mx, mp;
    {
        cuuint64_t dims[2] = {(cuuint64_t)ldk, (cuuint64_t)x8_rows};
        cuuint64_t strides[1] = {(cuuint64_t)ldk};
        cuuint32_t box[2] = {(cuuint32_t)kStageK, 128};
        cuuint32_t es[2] = {1, 1};
        if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)x8, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                (kStageK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)ldk, (cuuint64_t)npad, (cuuint64_t)kSlices};
        cuuint64_t strides[2] = {(cuuint64_t)ldk, (cuuint64_t)ldk * (cuuint64_t)npad};
        cuuint32_t box[3] = {(cuuint32_t)kStageK, (cuuint32_t)kTileEig, (cuuint32_t)kSlices};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(&mp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                (kStageK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -3;
    }
    Args a;
    a.mb = mb; a.n = n;
    a.snp_tiles = (int)((mb + kTileSnps - 1) / kTileSnps);
    a.eig_tiles = (n + kTileEig - 1) / kTileEig;
    a.scale = scale; a.xr = xr; a.ldx = ldx;
    // per call: the attribute is per device, and a process may hold handles on several devices
    if (cudaFuncSetAttribute(rotate_i8_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)
        return -4;
    const long long tiles = (long long)a.snp_tiles * a.eig_tiles;
    const int grid = (int)std::min<long long>(tiles, sm_count);
    rotate_i8_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(mx, mp, a);
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace tc
}  // namespace pg
