// rotate_i8_tc.cuh -- shared pieces of the fused int8-split rotation (rotate_i8_tc2.cuh): tile constants, mbarrier /
// TMA / tcgen05 inline-PTX helpers with watchdog waits, the UMMA shared-memory descriptor, the per-eigenvector scale
// kernel and the driver entry point for tensor-map encoding.  Reference: `X = U.T @ X`, lmm/lmm.py:244.
//
// Tile: per k-step (32 samples) MMAs of shape M = 128 (SNPs) x N = 224 (32 eigenvectors x 7 digit planes) x K = 32;
// accumulators live in TMEM columns [0, 224) and [256, 480).  Operands are K-major (or MN-major for the caller's
// sample-major block) with the 128-byte swizzle TMA writes and the UMMA descriptor reads.
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

// (the single-CTA kernel that used to live here -- 14.4 ms per 25 k SNPs against 10.0 for the CTA-pair kernel -- was
// removed in round 2; rotate_i8_tc2.cuh is the only fused engine and this header keeps the shared helpers)


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


}  // namespace tc
}  // namespace pg
