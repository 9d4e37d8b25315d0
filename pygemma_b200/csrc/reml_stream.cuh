// reml_stream.cuh -- the production per-SNP REML kernel (sm_100a).
//
// One warp owns one SNP at a time; a CTA of NW warps marches through the sample dimension in lock
// step so that the SNP-independent operands of every likelihood evaluation -- the eigenvalues d and the
// rotated [W0, y] columns -- are fetched from L2 once per CTA per pass (cp.async into a shared-memory
// ring) instead of once per SNP, while each warp streams its own rotated genotype vector through the
// same ring.  Every optimiser iteration (bracket scan, Brent, Newton, final likelihood) stays on chip:
//   pass    : x^T H^-p [W0, y, x],  p = 1,2(,3)        FP64 FMA, operands from shared memory
//   reduce  : warp shuffles -> x row of the packed (c0+2)^2 triangles in shared memory
//   tables  : [W0,y] block from the exact / Chebyshev lambda tables (pg_eval.cuh)
//   Pab     : projection recursion over covariates, warp-parallel (pg_eval.cuh)
//   solver  : SnpSolver state machine (pg_math.cuh) run by lane 0, state in shared memory
// Warps pull SNPs from a device-wide queue, so a CTA keeps running passes as long as any of its warps
// has a pending evaluation; different warps are usually at different optimiser stages (their lambda and
// number of powers differ), only the sample-dimension sweep is shared.
//
// Replaces: reference lmm/lmm.py:461-495 (calculate) and pygemma_model.pyx:64-194, :880-1053, :1349-1416,
// :1514-1537, :1631-1698, :1813-1830.
#pragma once

#include <cuda_runtime.h>

#include "reml_kernels.cuh"

namespace pg {

constexpr int kTile = 128;    // samples per shared-memory tile
constexpr int kStages = 3;    // cp.async ring depth

struct StreamCfg {
    int nw;          // warps per CTA
    int stages;
    size_t smem;     // dynamic shared memory bytes
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// 1/t for t in [1, 1e10]: hardware seed (2^-20) + two Newton steps; within 1 ulp, no special cases needed
// because t = lambda*d + 1 >= 1 is always a normal number here.
__device__ __forceinline__ double fast_rcp(double t)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    double e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// shared-memory layout of one ring stage: [d | w_0 .. w_{NCmax-1}] (kTile doubles each), then NW x-tiles
template <int NC, bool FULL, bool FIRST>
__device__ __forceinline__ void tile_compute(const double* __restrict__ sd, const double* __restrict__ sw,
                                             const double* __restrict__ sx, int lane, double lam, double (&a1)[NC],
                                             double (&a2)[NC], double (&a3)[NC], double& xx1, double& xx2,
                                             double& xx3)
{
    constexpr int kIt = kTile / 32;
    // phase 1: the serial chain d -> 1/(lam d + 1) -> x h^p for all steps of the tile at once (independent
    // chains overlap their latencies)
    double xh[kIt], xh2[kIt], xh3[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int l = it * 32 + lane;
        const double xv = sx[l];
        const double h = fast_rcp(fma(lam, sd[l], 1.0));
        xh[it] = xv * h;
        xh2[it] = xh[it] * h;
        xh3[it] = FULL ? xh2[it] * h : 0.0;
        if (FIRST) {
            xx1 = fma(xh[it], xv, xx1);
            xx2 = fma(xh2[it], xv, xx2);
            if (FULL) xx3 = fma(xh3[it], xv, xx3);
        }
    }
    // phase 2: rank-1 accumulation against the [W0, y] columns
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int l = it * 32 + lane;
        double wv[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) wv[j] = sw[j * kTile + l];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            a1[j] = fma(xh[it], wv[j], a1[j]);
            a2[j] = fma(xh2[it], wv[j], a2[j]);
            if (FULL) a3[j] = fma(xh3[it], wv[j], a3[j]);
        }
    }
}

// Operands of one lock-step sweep (passed by value: a by-reference struct would live in local memory)
struct PassArgs {
    const double* d;    // eigenvalues, zero-padded to a multiple of kTile
    const double* w;    // first column of this chunk; column j at w + j*ldw (zero-padded)
    long long ldw;
    long long ldx;      // padded length of a rotated genotype row (even, zero-filled)
    int n, c0, jb;
    double* ring;
    int stage_doubles, nw, ncmax;
};

// One lock-step sweep over the sample dimension for column chunk [jb, jb+NC) of [W0, y].
// active / lam / full are per-warp (warp-uniform); x is the warp's rotated genotype vector.
template <int NC, bool FIRST>
__device__ __noinline__ void stream_pass(PassArgs p, bool active, double lam, int full, const double* __restrict__ x,
                                         double* A, double* B, double* C)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = p.nw * 32;
    const int ntiles = (p.n + kTile - 1) / kTile;
    double a1[NC], a2[NC], a3[NC];
    double xx1 = 0.0, xx2 = 0.0, xx3 = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j) { a1[j] = 0.0; a2[j] = 0.0; a3[j] = 0.0; }

    constexpr int kChunksPerArr = kTile / 2;  // 16-byte chunks per array per tile
    constexpr int kTotalChunks = (NC + 1) * kChunksPerArr;
    double* const ring = p.ring;
    const int sd_off = 0, sw_off = kTile, sx_off = (1 + p.ncmax) * kTile + warp * kTile;

    auto load_tile = [&](int s, int t) {
        double* st = ring + (size_t)s * p.stage_doubles;
        const size_t off = (size_t)t * kTile;
        for (int c = tid; c < kTotalChunks; c += nthreads) {
            const int arr = c / kChunksPerArr, o = (c % kChunksPerArr) * 2;
            const double* src = (arr == 0) ? p.d + off + o : p.w + (size_t)(arr - 1) * p.ldw + off + o;
            cp_async16(st + (size_t)arr * kTile + o, src);
        }
        if (active) {
#pragma unroll
            for (int c = lane; c < kChunksPerArr; c += 32) {
                const long long pos = (long long)off + c * 2;
                const int valid = (pos + 2 <= p.ldx) ? 16 : 0;
                cp_async16_zfill(st + sx_off + c * 2, x + (valid ? pos : 0), valid);
            }
        }
    };

    for (int s = 0; s < kStages - 1; ++s) {
        if (s < ntiles) load_tile(s, s);
        cp_async_commit();
    }
    for (int t = 0; t < ntiles; ++t) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        const int tn = t + kStages - 1;
        if (tn < ntiles) load_tile(tn % kStages, tn);
        cp_async_commit();
        if (active) {
            const double* st = ring + (size_t)(t % kStages) * p.stage_doubles;
            if (full) tile_compute<NC, true, FIRST>(st + sd_off, st + sw_off, st + sx_off, lane, lam, a1, a2, a3, xx1, xx2, xx3);
            else tile_compute<NC, false, FIRST>(st + sd_off, st + sw_off, st + sx_off, lane, lam, a1, a2, a3, xx1, xx2, xx3);
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // ring is free for the next pass

    if (active) {
        const int c0 = p.c0;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const double s1 = warp_sum(a1[j]), s2 = warp_sum(a2[j]);
            const double s3 = full ? warp_sum(a3[j]) : 0.0;
            if (lane == 0) {
                const int col = p.jb + j;  // reduced column: < c0 -> W0 column, == c0 -> y
                const int dst = (col < c0) ? tri(c0, col) : tri(c0 + 1, c0);
                A[dst] = s1; B[dst] = s2;
                if (full) C[dst] = s3;
            }
        }
        if (FIRST) {
            const double s1 = warp_sum(xx1), s2 = warp_sum(xx2);
            const double s3 = full ? warp_sum(xx3) : 0.0;
            if (lane == 0) {
                const int dst = tri(c0, c0);
                A[dst] = s1; B[dst] = s2;
                if (full) C[dst] = s3;
            }
        }
    }
}

template <bool FIRST>
__device__ __forceinline__ void stream_pass_dispatch(int nc, const PassArgs& p, bool active, double lam, int full,
                                                     const double* x, double* A, double* B, double* C)
{
#define PG_SP(NCV)                                                          \
    case NCV:                                                               \
        stream_pass<NCV, FIRST>(p, active, lam, full, x, A, B, C);          \
        break;
    switch (nc) {
        PG_SP(1) PG_SP(2) PG_SP(3) PG_SP(4) PG_SP(5) PG_SP(6)
        PG_SP(7) PG_SP(8) PG_SP(9) PG_SP(10) PG_SP(11) PG_SP(12)
    default: break;
    }
#undef PG_SP
}

struct WarpSlot {
    SnpSolver s;
    long long g;  // SNP index within the block, -1: no work
};

__global__ void __launch_bounds__(256, 2) reml_stream_kernel(ScanArgs a, int nw, int ncmax)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.c0 + 2, TT = k * (k + 1) / 2, k0 = a.c0 + 1;
    const int stage_doubles = (1 + ncmax + nw) * kTile;
    double* ring = smem;
    double* tris = ring + (size_t)kStages * stage_doubles;
    double* A = tris + (size_t)warp * 3 * TT;
    double* B = A + TT;
    double* C = B + TT;
    WarpSlot* slots = reinterpret_cast<WarpSlot*>(tris + (size_t)nw * 3 * TT);
    WarpSlot& slot = slots[warp];

    auto fetch = [&]() {
        if (lane == 0) {
            const unsigned long long g = atomicAdd(a.counter, 1ULL);
            if (g < (unsigned long long)a.m) {
                slot.g = (long long)g;
                slot.s.init(a.n, a.c0, a.grid);
            } else {
                slot.g = -1;
            }
        }
        __syncwarp();
    };
    fetch();

    for (;;) {
        const bool active = slot.g >= 0;
        if (__syncthreads_count(active ? 1 : 0) == 0) break;
        double lam = 1.0;
        int full = 0, fixed_t = -1, need_ll = 0;
        const double* x = a.xr;
        if (active) {
            lam = slot.s.rq_lambda; full = slot.s.rq_full; fixed_t = slot.s.rq_fixed; need_ll = slot.s.rq_ll;
            x = a.xr + (size_t)slot.g * a.ldx;
        }
        // SNP-independent block of the triangles (global tables) while the first tiles are in flight
        Level0 l0{0.0, 0.0, 0.0};
        if (active) {
            if (full) assemble_w0y<true>(a.tab, lam, fixed_t, A, B, C, &l0);
            else assemble_w0y<false>(a.tab, lam, fixed_t, A, B, C, &l0);
        }
        // lock-step sweeps, one per chunk of [W0, y] columns
        for (int jb = 0; jb < k0; jb += kChunkCols) {
            const int nc = min(kChunkCols, k0 - jb);
            PassArgs p;
            p.d = a.d; p.w = a.wy + (size_t)jb * a.ldw; p.ldw = a.ldw; p.ldx = a.ldx; p.n = a.n; p.c0 = a.c0; p.jb = jb;
            p.ring = ring; p.stage_doubles = stage_doubles; p.nw = nw; p.ncmax = ncmax;
            if (jb == 0) stream_pass_dispatch<true>(nc, p, active, lam, full, x, A, B, C);
            else stream_pass_dispatch<false>(nc, p, active, lam, full, x, A, B, C);
        }
        if (active) {
            __syncwarp();
            EvalOut e;
            if (full) pab_recursion<true>(a.tab, A, B, C, l0, need_ll != 0, &e);
            else pab_recursion<false>(a.tab, A, B, C, l0, need_ll != 0, &e);
            if (lane == 0) {
                slot.s.feed(e);
                if (!slot.s.pending()) {
                    const long long row = a.row0 + slot.g;
                    a.out[0][row] = slot.s.beta; a.out[1][row] = slot.s.se; a.out[2][row] = slot.s.tau;
                    a.out[3][row] = slot.s.lambda; a.out[4][row] = slot.s.F; a.out[5][row] = slot.s.p;
                    if (a.status) a.status[row] = slot.s.status;
                    if (a.n_eval2) a.n_eval2[row] = slot.s.n_eval2;
                    if (a.n_eval3) a.n_eval3[row] = slot.s.n_eval3;
                }
            }
            __syncwarp();
            if (!slot.s.pending()) fetch();
        }
    }
}

inline StreamCfg stream_config(int c0)
{
    const int k = c0 + 2, TT = k * (k + 1) / 2;
    const int ncmax = (c0 + 1 < kChunkCols) ? c0 + 1 : kChunkCols;
    StreamCfg cfg;
    cfg.stages = kStages;
    for (int nw = 8; nw >= 1; nw >>= 1) {
        const size_t ring = (size_t)kStages * (1 + ncmax + nw) * kTile * sizeof(double);
        const size_t tri_bytes = (size_t)nw * 3 * TT * sizeof(double);
        const size_t slots = (size_t)nw * sizeof(WarpSlot);
        cfg.nw = nw;
        cfg.smem = ring + tri_bytes + slots + 16;
        if (cfg.smem <= 100 * 1024 || nw == 1) break;  // two CTAs per SM when possible
    }
    return cfg;
}

}  // namespace pg
