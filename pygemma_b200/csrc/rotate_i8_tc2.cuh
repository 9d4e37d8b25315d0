// rotate_i8_tc2.cuh -- the fused int8-split rotation on CTA PAIRS (tcgen05 cta_group::2).
//
// Same mathematics and the same bits as rotate_i8_tc.cuh; the difference is the operand economy.  A pair of CTAs on
// one TPC executes M = 256 MMAs: each CTA supplies its own 128 SNP rows of A and only HALF of B (112 of the 224
// plane-eigen rows), the tensor cores exchange the B halves.  Per CTA and k-step the shared-memory reads drop from
// 2 x (4 + 7) KB to 2 x (4 + 3.5) KB and the TMA fill from 60 KB to 46 KB per 128 samples, which is what the
// single-CTA kernel was bound by.  Cluster tile: 512 SNPs (2 CTAs x 2 accumulators x 128) x 32 eigenvectors x 7 planes.
//
// Accumulator columns are eigen-major here (column e*7 + p) so that one CTA's half of B is a contiguous range of
// eigenvectors: the B tensor map walks (sample, plane, eigenvector) with box (128, 7, 16).
//
//   warp 0    TMA producer (both CTAs): own A0, A1, B-half into own shared memory, transaction bytes onto the LEADER's
//             `full[s]` (peer bit of the mbarrier address cleared, as cute::SM100_TMA_2SM_LOAD does)
//   warp 1    TMEM allocation (both CTAs, cta_group::2); MMA issue by the leader CTA only; tcgen05.commit multicast
//             to `empty[s]` / `tmem_full` of both CTAs
//   warps 2-9 epilogue on the CTA's own TMEM (warps 2-5: accumulator 0, warps 6-9: accumulator 1; a warp reads the
//             TMEM lane quarter warp % 4); arrival on the leader's `tmem_empty` (remote for the peer CTA)
//
// MOMENT FUSION (template flag FUSE, round 2): the rotated genotypes are only ever consumed by the eigenvalue-space
// compression (compress.cuh), whose linear moments  z_jk = sum_l L_k(d_l) w_jl (U^T x)_l = x^T (U V)_{jk}  are linear in
// the RAW genotypes.  G = U V (n x klin * nodes, built once per design) is sliced into digit planes exactly like U^T and
// its 32-column tiles are appended to the eigen tiles of this kernel: their epilogue writes the moments straight into the
// SNP's slab.  On the eigen tiles the epilogue squares the rotated value in registers and accumulates the x^2 moments of
// its 16 eigenvectors per (half tile, segment) "piece" (moments_reduce_kernel sums the pieces in a fixed order); isolated
// eigenvalues (COPY rows) store x_l w_jl and x_l^2 directly.  The rotated genotypes never reach HBM: -160 KB per SNP of
// traffic and the whole compress_dmma_kernel (17 % of the step) for +10-14 % tensor work.
#pragma once

#include "compress_plan.h"
#include "pg_debug.cuh"
#include "rotate_i8_tc.cuh"

namespace pg {
namespace tc2 {

using tc::kABytes;
using tc::kAcc1Col;
using tc::kStageK;
using tc::kTileEig;
using tc::kTileN;
using tc::kTmemCols;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tmem_ld8;
using tc::umma_desc_sw128;

constexpr int kClusterSnps = 512;
constexpr int kCtaSnps = 256;
constexpr int kBHalfRows = kTileN / 2;              // 112
constexpr int kBHalfBytes = kBHalfRows * kStageK;   // 14 336
constexpr int kStageBytes = 2 * kABytes + kBHalfBytes;   // 47 104 per CTA
constexpr int kStages = 4;
#ifndef PG_TC2_EIG_GROUP
#define PG_TC2_EIG_GROUP 24
#endif
constexpr int kEigGroup = PG_TC2_EIG_GROUP;
#ifndef PG_TC2_EPI_WARPS
#define PG_TC2_EPI_WARPS 16
#endif
constexpr int kEpiWarps = PG_TC2_EPI_WARPS;              // 8: one warp per accumulator and lane quarter; 16: two (column halves)
constexpr int kThreads = 64 + 32 * kEpiWarps;           // TMA warp, MMA warp, epilogue warps
// FUSE: what the epilogue needs per tile column (scales, interpolation weights, piece / node / slab offsets), staged into
// shared memory by the epilogue warps WHILE the tile's MMAs run -- read from global memory inside the epilogue, the two
// dependent L2 latencies per eigenvector cost 11 us per tile (measured: 59 ms instead of 42 per 100 k SNPs)
struct TileMeta {
    double lw[kTileEig][kCq];   // interpolation weights of the tile's eigenvectors (rows are 16-byte aligned)
    double sc[kTileEig];        // fixed-point scale of the column
    double u1[kTileEig];        // U^T 1 / G^T 1 (level-coded genotypes)
    int2 ei[kTileEig];          // {piece, kq} / {node, 0} / {_, -1}
    int goff[kTileEig];         // G tiles: slab offset of the column, -1 padding
    int chunk_piece[kTileEig / 8];   // piece of an 8-eigenvector chunk when all eight are COMPRESS rows of one piece, else -1
    int pad[4];
};
static_assert(sizeof(TileMeta) % 16 == 0, "TileMeta is copied around in 16-byte aligned shared memory");
constexpr size_t kMetaBytes = 2 * sizeof(TileMeta);   // double-buffered by tile parity
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 + 256;
constexpr size_t kSmemBytesFused = kSmemBytes + kMetaBytes;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // cute::Sm100MmaPeerBitMask: address of the even CTA of the pair

// S32 accumulate, S8 x S8, K-major, N = 224, M = 256 (pair)
constexpr uint32_t kInstrDesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// L2 eviction policies: the digit planes of an eigen-tile group are re-read by every SNP tile (keep), genotype tiles
// are streamed once per group (evict first)
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// TMA loads whose completion bytes land on the leader CTA's barrier
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_local(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cta(uint64_t* bar, uint32_t cta)
{
    asm volatile(
        "{\n\t.reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}

// MN-major A tile straight from the caller's sample-major block: 128 sample rows of 128 SNP bytes, 128 B swizzle.
// Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute make_umma_desc<Major::MN>): SBO = 8 rows of
// 128 B; LBO (next 128-element chunk along M) is not exercised by an M = 128 tile.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(kABytes >> 4) << 16;     // leading byte offset: one whole tile
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 sample rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}

// eigen-indices per x^2 piece: what one epilogue thread holds of a tile
constexpr int kPieceEig = kTileEig / (kEpiWarps / 8);

struct FuseArgs {
    int g_tiles = 0;                 // tiles of 32 columns of G = U V that follow the eigen tiles
    const double* gscale = nullptr;  // [32 g_tiles] power-of-two scale of each G column's fixed-point form
    const double* g1 = nullptr;      // [32 g_tiles] column sums of G (level-coded genotypes: x = v0 + s code)
    const int* goff = nullptr;       // [32 g_tiles] offset of the column's moment inside an SNP slab, -1 for padding columns
    const int2* einfo = nullptr;     // [32 eig_tiles] {piece, kq > 0}: COMPRESS row; {node, 0}: COPY row; {_, -1}: padding
    const double* Lw = nullptr;      // [n][kCq] interpolation weights (compress_plan.h)
    const double* wy = nullptr;      // rotated [W0, y..] columns, column j at wy + j * ldw (COPY rows: x_l w_jl)
    long long ldw = 0;
    int klin = 0;
    const int* jrow = nullptr;       // [klin] slab offset of linear column j: z_row(j) * Kcp
    double* Z = nullptr;             // moment slabs of the block
    long long ldz = 0;               // doubles per SNP slab (zrows * Kcp)
    int x2row = 0;                   // slab offset of the x.x row (c0 * Kcp)
    double* P2 = nullptr;            // x^2 partial sums [piece][kCq][ldp]
    long long ldp = 0;
};

struct Args {
    long long mb;
    int n;
    int snp_tiles, eig_tiles;   // cluster tiles of 512 SNPs, tiles of 32 eigenvectors (of this launch)
    int eig_tile0 = 0;          // first eigen tile of this launch (one launch per eigen-tile group when the planes are pinned in L2)
    int eig_group;              // eigen tiles swept together (L2 residency of the B panels)
    int hints;                  // bit 0: evict_first for genotype tiles, bit 1: evict_last for the planes, bit 2: streaming stores of the rotated block, bit 3: 32-byte stores (default: see launch)
    const double* scale;
    double* xr;
    long long ldx;
    // level-coded float genotypes (rotate_i8.cuh): xr = v0 * u1 + s * (U^T code); second pass: xr += eps * (U^T [code == 2])
    const LevelInfo* info;   // nullable
    const double* u1;
    int accumulate;
    FuseArgs f;                 // FUSE kernels only
};

// x^2 moments: node k of segment s = sum over the segment's pieces, in eigen order (deterministic)
struct SegRed {
    int pb, pe;   // piece range
    int kb, kq;   // first node, node count
};
__global__ void __launch_bounds__(128) moments_reduce_kernel(const double* __restrict__ P2, long long ldp, long long mb,
                                                             const SegRed* __restrict__ segs, double* __restrict__ Z,
                                                             long long ldz, int x2row)
{
    const long long snp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (snp >= mb) return;
    const SegRed sr = segs[blockIdx.y];
    double acc[kCq];
#pragma unroll
    for (int k = 0; k < kCq; ++k) acc[k] = 0.0;
    for (int p = sr.pb; p < sr.pe; ++p) {
        const double* src = P2 + (size_t)p * kCq * ldp + snp;
#pragma unroll
        for (int k = 0; k < kCq; ++k)
            if (k < sr.kq) acc[k] += src[(size_t)k * ldp];
    }
    double* dst = Z + (size_t)snp * ldz + x2row + sr.kb;
#pragma unroll
    for (int k = 0; k < kCq; ++k)
        if (k < sr.kq) dst[k] = acc[k];
}

// A_MN: the A tensor map walks the caller's sample-major block (SNP, sample) and A is an MN-major operand;
// otherwise it walks the staged SNP-major copy (sample, SNP) and A is K-major.
template <bool A_MN, bool FUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
rotate_i8_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_p,
                     const __grid_constant__ CUtensorMap map_g, Args a)
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
    TileMeta* meta = (TileMeta*)(smem + (size_t)kStages * kStageBytes + 256);   // FUSE only (kSmemBytesFused)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int all_tiles = a.eig_tiles + (FUSE ? a.f.g_tiles : 0);   // eigen tiles, then the tiles of G
    const long long total_tiles = (long long)a.snp_tiles * all_tiles;
    const int ksteps = (a.n + kStageK - 1) / kStageK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * 32 * kEpiWarps);   // the epilogue threads of both CTAs (only the leader's barrier is used)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / TMA signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    auto decode = [&](long long t, int& st, int& et) {
        const long long per_group = (long long)a.eig_group * a.snp_tiles;
        const int g = (int)(t / per_group);
        const long long r = t - (long long)g * per_group;
        const int e0 = g * a.eig_group;
        const int ecount = min(a.eig_group, all_tiles - e0);
        st = (int)(r / ecount);
        et = a.eig_tile0 + e0 + (int)(r % ecount);
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
        for (long long t = cluster_id; t < total_tiles; t += num_clusters) {
            int st, et;
            decode(t, st, et);
            const int snp0 = st * kClusterSnps + (int)rank * kCtaSnps;
            for (int k = 0; k < ksteps; ++k) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sA = smem + (size_t)stage * kStageBytes;
                    if (leader) mbar_expect_tx_local(&full[stage], 2 * kStageBytes);   // both CTAs' bytes land here
                    if (A_MN) {
                        tma2_load_2d(sA, &map_x, &full[stage], snp0, k * kStageK, pol_a);
                        tma2_load_2d(sA + kABytes, &map_x, &full[stage], snp0 + 128, k * kStageK, pol_a);
                    } else {
                        tma2_load_2d(sA, &map_x, &full[stage], k * kStageK, snp0, pol_a);
                        tma2_load_2d(sA + kABytes, &map_x, &full[stage], k * kStageK, snp0 + 128, pol_a);
                    }
                    if (FUSE && et >= a.eig_tiles)
                        tma2_load_3d(sA + 2 * kABytes, &map_g, &full[stage], k * kStageK, 0,
                                     (et - a.eig_tiles) * kTileEig + (int)rank * (kTileEig / 2), pol_b);
                    else
                        tma2_load_3d(sA + 2 * kABytes, &map_p, &full[stage], k * kStageK, 0, et * kTileEig + (int)rank * (kTileEig / 2), pol_b);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long t = cluster_id; t < total_tiles; t += num_clusters) {
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
                            const uint64_t adv = (uint64_t)((kk * 32) >> 4);                      // K-major: 32 bytes along the row
                            const uint64_t adv_a = A_MN ? (uint64_t)((kk * 32 * 128) >> 4) : adv;   // MN-major: 32 sample rows
                            umma2_i8(tmem_base, da0 + adv_a, db + adv, accum, idesc);
                            umma2_i8(tmem_base + kAcc1Col, da1 + adv_a, db + adv, accum, idesc);
                        }
                        umma2_commit(&empty[stage]);
                        if (k == ksteps - 1) umma2_commit(tmem_full);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        // the accumulators (and, with 16 warps, their column halves) drain concurrently: a short serial chain per tile
        const int part = (warp - 2) >> 2;
        const int acc = part & 1;
        constexpr int kCPer = (kTileEig / 8) / (kEpiWarps / 8);
        const int c_begin = (part >> 1) * kCPer;
        uint32_t acc_phase = 0;
        const int etid = (int)threadIdx.x - 64;   // index among the epilogue threads
        for (long long t = cluster_id; t < total_tiles; t += num_clusters) {
            int st, et;
            decode(t, st, et);
            const bool g_tile = FUSE && et >= a.eig_tiles;          // a tile of G = U V: its values ARE linear moments
            const int eig0 = (g_tile ? et - a.eig_tiles : et) * kTileEig;   // first eigenvector / first column of G
            const TileMeta& M = meta[acc_phase];
            if (FUSE) {
                // stage this tile's column metadata while its MMAs run.  The buffer alternates with the accumulator phase:
                // no warp can be more than one tile ahead of another (tmem_empty gates the next tile's MMAs).
                TileMeta& Mw = meta[acc_phase];
                for (int i = etid; i < kTileEig * kCq; i += 32 * kEpiWarps) {
                    const int e = eig0 + i / kCq;
                    (&Mw.lw[0][0])[i] = (!g_tile && e < a.n) ? __ldg(a.f.Lw + (size_t)eig0 * kCq + i) : 0.0;
                }
                for (int i = etid; i < kTileEig; i += 32 * kEpiWarps) {
                    const int e = eig0 + i;
                    if (g_tile) {
                        Mw.sc[i] = __ldg(a.f.gscale + e);
                        Mw.u1[i] = a.info ? __ldg(a.f.g1 + e) : 0.0;
                        Mw.goff[i] = __ldg(a.f.goff + e);
                        Mw.ei[i] = make_int2(0, -1);
                    } else {
                        Mw.sc[i] = e < a.n ? __ldg(a.scale + e) : 0.0;
                        Mw.u1[i] = (a.info && e < a.n) ? __ldg(a.u1 + e) : 0.0;
                        Mw.goff[i] = -1;
                        Mw.ei[i] = __ldg(a.f.einfo + e);
                    }
                }
                if (etid < kTileEig / 8) {
                    int cp = -1;
                    if (!g_tile) {
                        const int2 e0 = __ldg(a.f.einfo + eig0 + etid * 8);
                        cp = e0.y > 0 ? e0.x : -1;
                        for (int j = 1; j < 8; ++j) {
                            const int2 ej = __ldg(a.f.einfo + eig0 + etid * 8 + j);
                            if (ej.y <= 0 || ej.x != cp) cp = -1;
                        }
                    }
                    Mw.chunk_piece[etid] = cp;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            }
            mbar_wait(tmem_full, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const long long snp = (long long)st * kClusterSnps + (long long)rank * kCtaSnps + acc * 128 + quarter * 32 + lane;
                double lv0 = 0.0, ls = 1.0, leps = 0.0;
                if (a.info && snp < a.mb) { const LevelInfo li = a.info[snp]; lv0 = li.v0; ls = li.s; leps = li.eps; }
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAcc1Col);
                // FUSE, eigen tiles: x^2 moments of the current piece (one segment inside this thread's kPieceEig eigenvectors)
                double m2[FUSE ? kCq : 1];
                int piece = -1;
                if (FUSE) {
#pragma unroll
                    for (int k = 0; k < (FUSE ? kCq : 1); ++k) m2[k] = 0.0;
                }
                auto flush_piece = [&]() {
                    if (FUSE) {
                        if (piece >= 0 && snp < a.mb) {
                            PG_BOUNDS(snp < a.f.ldp, "x^2 partial store outside the piece buffer");
                            double* pp = a.f.P2 + (size_t)piece * kCq * a.f.ldp + snp;
#pragma unroll
                            for (int k = 0; k < (FUSE ? kCq : 1); ++k) __stcs(pp + (size_t)k * a.f.ldp, m2[k]);   // streaming: the operand tiles own the L2
                        }
#pragma unroll
                        for (int k = 0; k < (FUSE ? kCq : 1); ++k) m2[k] = 0.0;
                    }
                };
                double* Zs = FUSE ? a.f.Z + (size_t)(snp < a.mb ? snp : 0) * a.f.ldz : nullptr;
#pragma unroll 1
                for (int c = c_begin; c < c_begin + kCPer; ++c) {
                    // eight eigenvectors x seven planes = 56 consecutive columns (eigen-major)
                    uint32_t r[kSlices][8];
                    PG_BOUNDS(c * 8 * kSlices + kSlices * 8 <= kTileN, "TMEM column range of an accumulator");
#pragma unroll
                    for (int q = 0; q < kSlices; ++q) tmem_ld8(tbase + (uint32_t)(c * 8 * kSlices + q * 8), r[q]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (FUSE || snp < a.mb) {
                        double* dst = FUSE ? nullptr : a.xr + (size_t)snp * a.ldx + eig0 + c * 8;
                        if (!FUSE)
                            PG_BOUNDS(snp >= 0 && snp < a.mb && eig0 + c * 8 < a.ldx && (eig0 + c * 8 + 8 <= a.n ? eig0 + c * 8 + 8 <= a.ldx : true),
                                      "rotated-genotype store outside the block");
                        double out[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int e = min(eig0 + c * 8 + j, a.n - 1);   // (!FUSE)
                            // the digit planes are recombined as 64-bit integers (|plane sum| < 2^31, so hi < 2^48 and
                            // lo < 2^56 never overflow) and converted once each: 2 int -> double conversions per output
                            // instead of 7.  Same bits as the FP64 Horner form of combine_i8_kernel: both are the exact
                            // integer, rounded once if it exceeds 2^53 (it does not for dosages: |plane sum| <= 256 n)
#define PG_PL(P) ((long long)(int)r[(j * kSlices + (P)) >> 3][(j * kSlices + (P)) & 7])
                            const long long hi_i = (PG_PL(0) << 16) + (PG_PL(1) << 8) + PG_PL(2);
                            long long lo_i = PG_PL(3);
#pragma unroll
                            for (int p = 4; p < kSlices; ++p) lo_i = (lo_i << 8) + PG_PL(p);
#undef PG_PL
                            const double hi = (double)hi_i, lo = (double)lo_i;
                            const double v = fma(lo, kLoScale, hi);
                            out[j] = v * (FUSE ? M.sc[c * 8 + j] : __ldg(a.scale + e));
                            if (a.info) {
                                if (!FUSE && a.accumulate) out[j] = (leps != 0.0) ? fma(leps, out[j], dst[j]) : dst[j];
                                else out[j] = fma(ls, out[j], lv0 * (FUSE ? M.u1[c * 8 + j] : __ldg(a.u1 + e)));
                            }
                        }
                        if (FUSE) {
                            if (g_tile) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int off = M.goff[c * 8 + j];
                                    PG_BOUNDS(off < a.f.ldz, "linear-moment store outside the slab");
                                    if (off >= 0 && snp < a.mb) __stcs(Zs + off, out[j]);
                                }
                            } else if (M.chunk_piece[c] >= 0) {
                                // the common case: eight COMPRESS rows of one piece, straight-line code
                                if (M.chunk_piece[c] != piece) { flush_piece(); piece = M.chunk_piece[c]; }
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const double sq = out[j] * out[j];
                                    const double2* lw2 = reinterpret_cast<const double2*>(&M.lw[c * 8 + j][0]);
#pragma unroll
                                    for (int k2 = 0; k2 < (FUSE ? kCq / 2 : 0); ++k2) {
                                        const double2 w = lw2[k2];
                                        m2[2 * k2] = fma(w.x, sq, m2[2 * k2]);
                                        m2[2 * k2 + 1] = fma(w.y, sq, m2[2 * k2 + 1]);
                                    }
                                }
                            } else {
                                // segment boundaries, isolated eigenvalues, padding: one eigenvector at a time
#pragma unroll 1
                                for (int j = 0; j < 8; ++j) {
                                    double oj = 0.0;   // out[j] without dynamic register indexing
#pragma unroll
                                    for (int jj = 0; jj < 8; ++jj) oj = (jj == j) ? out[jj] : oj;
                                    const int2 ei = M.ei[c * 8 + j];
                                    const int e = eig0 + c * 8 + j;
                                    if (ei.y > 0) {
                                        if (ei.x != piece) { flush_piece(); piece = ei.x; }
                                        const double sq = oj * oj;
#pragma unroll
                                        for (int k = 0; k < (FUSE ? kCq : 1); ++k) m2[k] = fma(M.lw[c * 8 + j][k], sq, m2[k]);
                                    } else if (ei.y == 0 && snp < a.mb) {
                                        // isolated eigenvalue, its own node: x_l w_jl and x_l^2 (compress_copy_kernel)
                                        PG_BOUNDS(a.f.x2row + ei.x < a.f.ldz, "COPY-row moment store outside the slab");
                                        for (int jj = 0; jj < a.f.klin; ++jj)
                                            __stcs(Zs + __ldg(a.f.jrow + jj) + ei.x, oj * __ldg(a.f.wy + (size_t)jj * a.f.ldw + e));
                                        __stcs(Zs + a.f.x2row + ei.x, oj * oj);
                                    }
                                }
                            }
                        } else {
                            if (eig0 + c * 8 + 8 <= a.n) {   // rows are 128-byte aligned (ldx % 16 == 0): four 16-byte stores
                                if (a.hints & 8) {   // two 32-byte stores (STG.256): half the store instructions / LSU wavefronts
                                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(out[0]), "d"(out[1]),
                                                 "d"(out[2]), "d"(out[3]) : "memory");
                                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "d"(out[4]), "d"(out[5]),
                                                 "d"(out[6]), "d"(out[7]) : "memory");
                                } else if (a.hints & 4) {   // streaming stores: the rotated block is read back from HBM anyway
#pragma unroll
                                    for (int j = 0; j < 8; j += 2) __stcs(reinterpret_cast<double2*>(dst + j), make_double2(out[j], out[j + 1]));
                                } else {
#pragma unroll
                                    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(out[j], out[j + 1]);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (eig0 + c * 8 + j < a.n) dst[j] = out[j];
                            }
                        }
                    }
                }
                flush_piece();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_on_cta(tmem_empty, 0);   // the leader's barrier gates the next tile's first MMA
            acc_phase ^= 1;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // the peer may still be reading this CTA's shared memory / finishing its epilogue
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// L2 persistence of the digit planes (see launch).  Measured on the default bench command (tools/run_r2_s34.sh, back to back):
// groups of 16 eigen tiles (36 MB of planes) in a 40 MB set-aside 58.5-58.6 ms per step against 58.8-58.9 without, groups of
// 24 in 56-64 MB 58.7-59.9, groups of 32-40 in 80-100 MB 62 (too little L2 left for the streaming genotype tiles); DRAM reads of
// the rotation 3.5 GB per 16 384 SNPs instead of 10.5 (ncu).  PG_TC2_PERSIST=0 / 1 overrides the default.
constexpr int kPersistDefault = 1;
constexpr int kPersistMB = 40;      // L2 set-aside requested for the plane group
constexpr int kPersistGroup = 16;   // eigen tiles per launch
constexpr int kPersistMinWaves = 8; // a launch must hold this many waves of cluster tiles, else one launch over all tiles
inline bool persist_planes()
{
    static const int v = getenv("PG_TC2_PERSIST") ? atoi(getenv("PG_TC2_PERSIST")) : kPersistDefault;
    return v != 0;
}

// Eigen tiles per launch when the digit planes are pinned in the L2 set-aside, 0 = one launch over all tiles (host arithmetic;
// pg_probe_rotation_launches exposes it to the CPU tests).  ldk: row pitch of the planes; setaside / max_window: bytes.
inline int persist_group_tiles(int snp_tiles, int ldk, size_t setaside, size_t max_window, int sm_count, int want)
{
    if (setaside == 0 || max_window == 0 || want < 1) return 0;
    const size_t tile_bytes = (size_t)kTileEig * kSlices * (size_t)ldk;
    const int g = (int)std::min<size_t>({(size_t)want, setaside / tile_bytes, max_window / tile_bytes});
    if (g < 1 || (long long)snp_tiles * g < (long long)kPersistMinWaves * (sm_count / 2)) return 0;
    return g;
}

// xsm != nullptr: the caller's sample-major int8 block (element (sample j, SNP g) at xsm[j*ld_sm + g]) is used directly
// as an MN-major operand; otherwise x8 (SNP-major, staged) is the K-major operand.
inline int launch(cudaStream_t stream, int sm_count, const int8_t* x8, long long x8_rows, const int8_t* planes, int npad,
                  int ldk, int n, long long mb, const double* scale, double* xr, long long ldx,
                  const int8_t* xsm = nullptr, long long ld_sm = 0, const LevelInfo* info = nullptr,
                  const double* u1 = nullptr, int accumulate = 0, const FuseArgs* fuse = nullptr,
                  const int8_t* planes_g = nullptr, int* n_launched = nullptr)
{
    if (n_launched) *n_launched = 1;
    tc::EncodeTiledFn enc = tc::encode_tiled_fn();
    if (!enc) return -1;
    if (fuse && (accumulate || !planes_g)) return -6;
    CUtensorMap mx, mp, mg;
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
        // (sample, plane, eigenvector) over the [npad][kSlices][ldk] plane array: a box of 16 eigenvectors x 7 planes lands
        // eigen-major, 112 rows of 128 B
        cuuint64_t dims[3] = {(cuuint64_t)ldk, (cuuint64_t)kSlices, (cuuint64_t)npad};
        cuuint64_t strides[2] = {(cuuint64_t)ldk, (cuuint64_t)kSlices * (cuuint64_t)ldk};   // planes of an eigenvector are adjacent rows
        cuuint32_t box[3] = {(cuuint32_t)kStageK, (cuuint32_t)kSlices, (cuuint32_t)(kTileEig / 2)};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(&mp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -3;
        mg = mp;
        if (fuse) {   // the digit planes of G: [32 g_tiles][kSlices][ldk]
            const int npad_g = fuse->g_tiles * kTileEig;
            dims[2] = (cuuint64_t)npad_g;
            if (enc(&mg, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)planes_g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return -3;
        }
    }
    Args a;
    a.mb = mb; a.n = n;
    a.snp_tiles = (int)((mb + kClusterSnps - 1) / kClusterSnps);
    a.eig_tiles = (n + kTileEig - 1) / kTileEig;
    a.scale = scale; a.xr = xr; a.ldx = ldx;
    a.info = info; a.u1 = u1; a.accumulate = accumulate;
    if (fuse) a.f = *fuse;
    static const int eg_env = getenv("PG_TC2_EG") ? atoi(getenv("PG_TC2_EG")) : 0;
    a.eig_group = eg_env > 0 ? eg_env : kEigGroup;
    // measured at n = 10 000 per 25 088 SNPs: no hint 10.17 ms, evict_last(B) 10.13, evict_first(A) 10.86, both 10.81
    // rotated block written with two 32-byte stores per thread and chunk instead of four 16-byte ones (bit 3): 59.1-59.8 ms
    // per bench step against 60.1-60.6 (three pairs, back to back); streaming stores (bit 2): inside the noise
    static const int hints_env = getenv("PG_TC2_HINTS") ? atoi(getenv("PG_TC2_HINTS")) : 10;
    a.hints = hints_env;
    // per call: the attribute is per device, and a process may hold handles on several devices
    if (cudaFuncSetAttribute(rotate_i8_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(rotate_i8_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(rotate_i8_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesFused) != cudaSuccess ||
        cudaFuncSetAttribute(rotate_i8_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesFused) != cudaSuccess)
        return -4;
    // L2-resident digit planes (VERDICT r1 task 4): one launch per group of eigen tiles with the group's planes -- one contiguous
    // range of the [npad][kSlices][ldk] array -- pinned in the L2 set-aside by an access-policy window (it applies to the TMA
    // loads), so that only the genotype tiles stream through the rest of the L2.  In a single launch a wave of 74 K-deep cluster
    // tiles touches more operand bytes than the L2 keeps and every wave re-reads its planes from HBM (profiles/l2_probe_r02_dram.txt).
    // Only when every launch still holds kPersistMinWaves waves of tiles: short blocks (the ramp of host-resident input, one
    // rank's shard of a problem split over 8 GPUs) and large n (a group of planes no longer fits) keep the single launch.
    if (persist_planes() && !fuse) {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        static const int mb_env = getenv("PG_TC2_PERSIST_MB") ? atoi(getenv("PG_TC2_PERSIST_MB")) : kPersistMB;
        const size_t setaside = std::min<size_t>((size_t)max_persist, (size_t)mb_env << 20);
        {
            static const int g_env = getenv("PG_TC2_PERSIST_EG") ? atoi(getenv("PG_TC2_PERSIST_EG")) : kPersistGroup;
            const int g = persist_group_tiles(a.snp_tiles, ldk, setaside, (size_t)std::max(max_window, 0), sm_count, g_env);
            const int all = a.eig_tiles;
            int rc = 0;
            if (g >= 1) {
              static size_t configured[64] = {0};   // per device: the limit is set once (a benign race between host threads)
              if (dev < 0 || dev >= 64 || configured[dev] != setaside) {
                  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside);
                  if (dev >= 0 && dev < 64) configured[dev] = setaside;
              }
              for (int e0 = 0; e0 < all && rc == 0; e0 += g) {
                const int cnt = std::min(g, all - e0);
                cudaStreamAttrValue v;
                memset(&v, 0, sizeof v);
                const size_t first_row = (size_t)e0 * kTileEig, rows = std::min<size_t>((size_t)cnt * kTileEig, (size_t)npad - first_row);
                v.accessPolicyWindow.base_ptr = (void*)(planes + first_row * kSlices * ldk);
                v.accessPolicyWindow.num_bytes = rows * kSlices * ldk;
                v.accessPolicyWindow.hitRatio = 1.0f;
                v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { rc = -7; break; }
                a.eig_tile0 = e0; a.eig_tiles = cnt; a.eig_group = cnt;
                const long long tiles_g = (long long)a.snp_tiles * cnt;
                const int clusters_g = (int)std::min<long long>(tiles_g, sm_count / 2);
                if (xsm) rotate_i8_tc2_kernel<true, false><<<2 * clusters_g, kThreads, kSmemBytes, stream>>>(mx, mp, mg, a);
                else rotate_i8_tc2_kernel<false, false><<<2 * clusters_g, kThreads, kSmemBytes, stream>>>(mx, mp, mg, a);
                if (cudaGetLastError() != cudaSuccess) rc = -5;
              }
              if (n_launched) *n_launched = (all + g - 1) / g;
              cudaStreamAttrValue off;
              memset(&off, 0, sizeof off);
              cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &off);   // num_bytes = 0: no window
              return rc;
            }
        }
    }
    const long long tiles = (long long)a.snp_tiles * (a.eig_tiles + (fuse ? fuse->g_tiles : 0));
    const int clusters = (int)std::min<long long>(tiles, sm_count / 2);
    if (fuse) {
        if (xsm) rotate_i8_tc2_kernel<true, true><<<2 * clusters, kThreads, kSmemBytesFused, stream>>>(mx, mp, mg, a);
        else rotate_i8_tc2_kernel<false, true><<<2 * clusters, kThreads, kSmemBytesFused, stream>>>(mx, mp, mg, a);
    } else {
        if (xsm) rotate_i8_tc2_kernel<true, false><<<2 * clusters, kThreads, kSmemBytes, stream>>>(mx, mp, mg, a);
        else rotate_i8_tc2_kernel<false, false><<<2 * clusters, kThreads, kSmemBytes, stream>>>(mx, mp, mg, a);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace tc2
}  // namespace pg
