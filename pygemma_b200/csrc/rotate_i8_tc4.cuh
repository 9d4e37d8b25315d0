// rotate_i8_tc4.cuh -- the fused int8-split rotation on clusters of TWO CTA pairs with the genotype tiles multicast.
//
// Same mathematics and the same bits as rotate_i8_tc2.cuh.  What changes is how often a genotype byte leaves the L2.
// The CTA-pair kernel is bound by the L2 -> SM path, not by the tensor pipe: ncu (profiles/ncu_r02_tc2_16384snps.json)
// shows 74.5 GB of TMA loads per 16 384 SNPs = 11.06 TB/s, about 6 700 B per L2 clock -- the chip-wide L2 output ceiling --
// with the tensor pipe 78 % busy.  Per cluster tile (512 SNPs x 32 eigenvectors x 7 planes) a pair pulls 5.1 MB of
// genotypes (A) and 2.24 MB of digit planes (B).  Here a cluster of four CTAs = two pairs works on the SAME 512 SNPs and
// two neighbouring eigen tiles: every A sub-tile is fetched once and multicast into both pairs
// (cp.async.bulk.tensor ... .multicast::cluster), i.e. 5.1 + 2 x 2.24 = 9.6 MB per two tiles instead of 14.7 (-35 %).
//
//   CTA rank r: pair p = r >> 1 (eigen tile 2 u + p), SNP half h = r & 1 (SNP rows st * 512 + h * 256 ...).
//   warp 0    TMA producer: own B half (unicast, as in the pair kernel) + ONE of the two 128-row A sub-tiles of its SNP half
//             (sub-tile p), multicast to CTAs {h, h + 2}; all bytes of a pair land on that pair leader's `full[s]`
//             (the mbarrier operand has the peer bit cleared: every destination CTA signals the even CTA of its pair)
//   warp 1    per pair as before; `tcgen05.commit` of a stage goes to `empty[s]` of ALL FOUR CTAs (count 2: a stage may be
//             refilled once BOTH pairs have consumed it, because either pair's producers write into both pairs)
//   warps 2.. epilogue, unchanged (per pair accumulators, per pair `tmem_full` / `tmem_empty`)
// The two pairs therefore run in lock step at stage granularity.  An odd number of eigen tiles leaves pair 1 of the last
// unit with a tile past the end: its B box is out of bounds (TMA zero fill) and its stores are masked by eig < n.
#pragma once

#include "rotate_i8_tc2.cuh"

namespace pg {
namespace tc4 {

// tile constants, helpers and Args of the pair kernel
using tc2::Args;
using tc2::cluster_ctarank;
using tc2::cluster_sync_all;
using tc2::kABytes;
using tc2::kAcc1Col;
using tc2::kClusterSnps;
using tc2::kCtaSnps;
using tc2::kEigGroup;
using tc2::kEpiWarps;
using tc2::kInstrDesc2;
using tc2::kPeerBitMask;
using tc2::kSmemBytes;
using tc2::kStageBytes;
using tc2::kStageK;
using tc2::kStages;
using tc2::kThreads;
using tc2::kTileEig;
using tc2::kTileN;
using tc2::kTmemCols;
using tc2::l2_policy_evict_first;
using tc2::l2_policy_evict_last;
using tc2::mbar_arrive_on_cta;
using tc2::mbar_expect_tx_local;
using tc2::mbar_init;
using tc2::mbar_wait;
using tc2::smem_u32;
using tc2::tma2_load_3d;
using tc2::tmem_ld8;
using tc2::umma2_i8;
using tc2::umma_desc_mn_sw128;
using tc2::umma_desc_sw128;

__device__ __forceinline__ void tma4_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask,
                                                uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
        " [%0], [%1, {%4, %5}], [%2], %3, %6;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "h"(mask), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void umma4_commit(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <bool A_MN>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kThreads, 1)
rotate_i8_tc4_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_p, Args a)
{
    constexpr uint32_t idesc = kInstrDesc2 | (A_MN ? (1u << 15) : 0u);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)kStages * kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = rank >> 1, half = rank & 1;
    const bool leader = half == 0;                      // MMA issuer of this pair
    const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
    const int cluster_id = blockIdx.x >> 2, num_clusters = gridDim.x >> 2;
    const int eig_units = (a.eig_tiles + 1) >> 1;       // units of two neighbouring eigen tiles
    const long long total_units = (long long)a.snp_tiles * eig_units;
    const int ksteps = (a.n + kStageK - 1) / kStageK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }   // empty: both pairs' commits
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * 32 * kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    // unit t -> (SNP tile, eigen unit): groups of eig_group/2 units are swept together, SNP tiles inside a group
    auto decode = [&](long long t, int& st, int& eu) {
        const int ug = max(1, a.eig_group >> 1);
        const long long per_group = (long long)ug * a.snp_tiles;
        const int g = (int)(t / per_group);
        const long long r = t - (long long)g * per_group;
        const int u0 = g * ug;
        const int ucount = min(ug, eig_units - u0);
        st = (int)(r / ucount);
        eu = u0 + (int)(r % ucount);
    };

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        uint64_t pol_a, pol_b;
        {
            uint64_t pol_n;
            asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_n));
            pol_a = (a.hints & 1) ? l2_policy_evict_first() : pol_n;
            pol_b = (a.hints & 2) ? l2_policy_evict_last() : pol_n;
        }
        const uint16_t a_mask = (uint16_t)((1u << half) | (1u << (half + 2)));   // the same SNP half in both pairs
        for (long long t = cluster_id; t < total_units; t += num_clusters) {
            int st, eu;
            decode(t, st, eu);
            const int et = 2 * eu + (int)pair;
            const int snp0 = st * kClusterSnps + (int)half * kCtaSnps + (int)pair * 128;   // the sub-tile THIS CTA multicasts
            for (int k = 0; k < ksteps; ++k) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sA = smem + (size_t)stage * kStageBytes;
                    if (leader) mbar_expect_tx_local(&full[stage], 2 * kStageBytes);   // all bytes landing in this pair
                    uint8_t* dstA = sA + (size_t)pair * kABytes;                        // sub-tile p sits at offset p * kABytes
                    if (A_MN) tma4_load_2d_mc(dstA, &map_x, &full[stage], snp0, k * kStageK, a_mask, pol_a);
                    else tma4_load_2d_mc(dstA, &map_x, &full[stage], k * kStageK, snp0, a_mask, pol_a);
                    tma2_load_3d(sA + 2 * kABytes, &map_p, &full[stage], k * kStageK, 0, et * kTileEig + (int)half * (kTileEig / 2), pol_b);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long t = cluster_id; t < total_units; t += num_clusters) {
                mbar_wait(tmem_empty, acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int k = 0; k < ksteps; ++k) {
                    mbar_wait(&full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint32_t sA = smem_u32(smem + (size_t)stage * kStageBytes);
                        const uint64_t da0 = A_MN ? umma_desc_mn_sw128(sA) : umma_desc_sw128(sA);
                        const uint64_t da1 = A_MN ? umma_desc_mn_sw128(sA + kABytes) : umma_desc_sw128(sA + kABytes);
                        const uint64_t db = umma_desc_sw128(sA + 2 * kABytes);
#pragma unroll
                        for (int kk = 0; kk < kStageK / 32; ++kk) {
                            const uint32_t accum = (k | kk) ? 1u : 0u;
                            const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                            const uint64_t adv_a = A_MN ? (uint64_t)((kk * 32 * 128) >> 4) : adv;
                            umma2_i8(tmem_base, da0 + adv_a, db + adv, accum, idesc);
                            umma2_i8(tmem_base + kAcc1Col, da1 + adv_a, db + adv, accum, idesc);
                        }
                        umma4_commit(&empty[stage], (uint16_t)0xF);        // the stage is free once BOTH pairs have read it
                        if (k == ksteps - 1) umma4_commit(tmem_full, pair_mask);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        const int part = (warp - 2) >> 2;
        const int acc = part & 1;
        constexpr int kCPer = (kTileEig / 8) / (kEpiWarps / 8);
        const int c_begin = (part >> 1) * kCPer;
        uint32_t acc_phase = 0;
        for (long long t = cluster_id; t < total_units; t += num_clusters) {
            int st, eu;
            decode(t, st, eu);
            const int et = 2 * eu + (int)pair;
            mbar_wait(tmem_full, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int eig0 = et * kTileEig;
            {
                const long long snp = (long long)st * kClusterSnps + (long long)half * kCtaSnps + acc * 128 + quarter * 32 + lane;
                double lv0 = 0.0, ls = 1.0, leps = 0.0;
                if (a.info && snp < a.mb) { const LevelInfo li = a.info[snp]; lv0 = li.v0; ls = li.s; leps = li.eps; }
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAcc1Col);
#pragma unroll 1
                for (int c = c_begin; c < c_begin + kCPer; ++c) {
                    uint32_t r[kSlices][8];
                    PG_BOUNDS(c * 8 * kSlices + kSlices * 8 <= kTileN, "TMEM column range of an accumulator");
#pragma unroll
                    for (int q = 0; q < kSlices; ++q) tmem_ld8(tbase + (uint32_t)(c * 8 * kSlices + q * 8), r[q]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (snp < a.mb && eig0 + c * 8 < a.n) {
                        double* dst = a.xr + (size_t)snp * a.ldx + eig0 + c * 8;
                        PG_BOUNDS(eig0 + c * 8 < a.ldx, "rotated-genotype store outside the block");
                        double out[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int e = min(eig0 + c * 8 + j, a.n - 1);
#define PG_PL(P) ((long long)(int)r[(j * kSlices + (P)) >> 3][(j * kSlices + (P)) & 7])
                            const long long hi_i = (PG_PL(0) << 16) + (PG_PL(1) << 8) + PG_PL(2);
                            long long lo_i = PG_PL(3);
#pragma unroll
                            for (int p = 4; p < kSlices; ++p) lo_i = (lo_i << 8) + PG_PL(p);
#undef PG_PL
                            const double hi = (double)hi_i, lo = (double)lo_i;
                            const double v = fma(lo, kLoScale, hi);
                            out[j] = v * __ldg(a.scale + e);
                            if (a.info) {
                                if (a.accumulate) out[j] = (leps != 0.0) ? fma(leps, out[j], dst[j]) : dst[j];
                                else out[j] = fma(ls, out[j], lv0 * __ldg(a.u1 + e));
                            }
                        }
                        if (eig0 + c * 8 + 8 <= a.n) {
#pragma unroll
                            for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(out[j], out[j + 1]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (eig0 + c * 8 + j < a.n) dst[j] = out[j];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_on_cta(tmem_empty, rank & ~1u);   // this pair's leader gates its next tile's first MMA
            acc_phase ^= 1;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// Number of 4-CTA clusters that can be co-resident (GPCs whose SM count is not a multiple of four leave SMs idle); 0 if the
// query fails.
template <bool A_MN>
inline int max_clusters()
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, rotate_i8_tc4_kernel<A_MN>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nc;
}

// Same contract as tc2::launch.
inline int launch(cudaStream_t stream, int sm_count, const int8_t* x8, long long x8_rows, const int8_t* planes, int npad,
                  int ldk, int n, long long mb, const double* scale, double* xr, long long ldx,
                  const int8_t* xsm = nullptr, long long ld_sm = 0, const LevelInfo* info = nullptr,
                  const double* u1 = nullptr, int accumulate = 0)
{
    tc::EncodeTiledFn enc = tc::encode_tiled_fn();
    if (!enc) return -1;
    CUtensorMap mx, mp;
    if (xsm) {
        cuuint64_t dims[2] = {(cuuint64_t)mb, (cuuint64_t)n};
        cuuint64_t strides[1] = {(cuuint64_t)ld_sm};
        cuuint32_t box[2] = {128, (cuuint32_t)kStageK};
        cuuint32_t es[2] = {1, 1};
        if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)xsm, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    } else {
        cuuint64_t dims[2] = {(cuuint64_t)ldk, (cuuint64_t)x8_rows};
        cuuint64_t strides[1] = {(cuuint64_t)ldk};
        cuuint32_t box[2] = {(cuuint32_t)kStageK, 128};
        cuuint32_t es[2] = {1, 1};
        if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)x8, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)ldk, (cuuint64_t)kSlices, (cuuint64_t)npad};
        cuuint64_t strides[2] = {(cuuint64_t)ldk, (cuuint64_t)kSlices * (cuuint64_t)ldk};   // planes of an eigenvector are adjacent rows
        cuuint32_t box[3] = {(cuuint32_t)kStageK, (cuuint32_t)kSlices, (cuuint32_t)(kTileEig / 2)};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(&mp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -3;
    }
    Args a;
    a.mb = mb; a.n = n;
    a.snp_tiles = (int)((mb + kClusterSnps - 1) / kClusterSnps);
    a.eig_tiles = (n + kTileEig - 1) / kTileEig;
    a.scale = scale; a.xr = xr; a.ldx = ldx;
    a.info = info; a.u1 = u1; a.accumulate = accumulate;
    static const int eg_env = getenv("PG_TC2_EG") ? atoi(getenv("PG_TC2_EG")) : 0;
    a.eig_group = eg_env > 0 ? eg_env : kEigGroup;
    static const int hints_env = getenv("PG_TC2_HINTS") ? atoi(getenv("PG_TC2_HINTS")) : 2;
    a.hints = hints_env;
    if (cudaFuncSetAttribute(rotate_i8_tc4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(rotate_i8_tc4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)
        return -4;
    static int cap_mn = -1, cap_k = -1;   // co-resident 4-CTA clusters (per process; every B200 of a box is alike)
    int& cap = xsm ? cap_mn : cap_k;
    if (cap < 0) cap = xsm ? max_clusters<true>() : max_clusters<false>();
    const long long units = (long long)a.snp_tiles * ((a.eig_tiles + 1) / 2);
    const int hw = cap > 0 ? cap : sm_count / 4;
    const int clusters = (int)std::min<long long>(units, hw);
    if (xsm) rotate_i8_tc4_kernel<true><<<4 * clusters, kThreads, kSmemBytes, stream>>>(mx, mp, a);
    else rotate_i8_tc4_kernel<false><<<4 * clusters, kThreads, kSmemBytes, stream>>>(mx, mp, a);
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace tc4
}  // namespace pg
