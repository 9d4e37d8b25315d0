// pg_api.cu -- C-ABI implementation (include/pygemma_b200.h): handle, eigendecomposition, design
// rotation, table building and the blocked scan pipeline (upload -> stage -> rotate -> REML -> download).
//
// Host-side mirror of the reference driver lmm/lmm.py:87-411; the reference's multiprocessing chunking
// (SampleIter, lmm/lmm.py:413-436) becomes SNP blocks streamed through one GPU.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pygemma_b200.h"
#include "compress.cuh"
#include "reml_kernels.cuh"
#include "reml_solve.cuh"
#include "reml_stream.cuh"
#include "rotate_kernels.cuh"

using namespace pg;

static thread_local std::string g_create_error;

// Dynamic shared memory above which a launch opts in with cudaFuncAttributeMaxDynamicSharedMemorySize.  The 48 KB default
// limit counts STATIC + dynamic bytes and the solver kernels carry up to ~10 KB of static state (one SnpSolver per tile), so
// the opt-in starts well below 48 KB (found by the round-2 fuzz sweep: c0 = 13 asked for 46.3 KB dynamic + static > 48 KB).
static constexpr size_t kDynSmemOptIn = 16 * 1024;

// SNP-block boundaries of a scan (see scan_impl, "Block boundaries"): plain blocks of `blk`, preceded for host-resident
// genotypes with automatic blocking by short blocks growing x 1.6 from 1 536 SNPs (multiples of 256).  PG_RAMP=0: none.
// Pageable host genotypes of at least this size go through the pinned bounce buffers (and are cut into about four blocks so
// that packing, upload and compute overlap).  16 MB: the driver's pageable 2-D copy runs at ~5-10 GB/s, so the 45 MB of the
// GD449 shape (449 x 100 000 int8) took 9 ms of an 11 ms call; round 2 started with 64 MB.
static constexpr size_t kBounceMinBytes = size_t(16) << 20;

static std::vector<long long> plan_blocks(long long m, long long blk, bool host_input, bool fixed_block)
{
    std::vector<long long> bstart;
    long long g = 0;
    static const bool ramp = !(getenv("PG_RAMP") && atoi(getenv("PG_RAMP")) == 0);
    if (ramp && host_input && !fixed_block && m >= 2 * blk) {
        for (long long cur = 1536; cur < blk && g + cur + 1024 < m; cur = (cur * 8 / 5 + 255) / 256 * 256) {
            bstart.push_back(g);
            g += cur;
        }
    }
    for (; g < m; g += blk) bstart.push_back(g);
    bstart.push_back(m);
    return bstart;
}

extern "C" int pg_probe_block_plan(int64_t m, int64_t blk, int host_input, int fixed_block, int64_t* starts, int cap)
{
    if (m < 0 || blk <= 0 || (!starts && cap > 0)) return PG_ERR_ARG;
    if (m == 0) return 0;
    const std::vector<long long> b = plan_blocks(m, blk, host_input != 0, fixed_block != 0);
    for (int i = 0; i < (int)b.size() && i < cap; ++i) starts[i] = b[i];
    return (int)b.size() - 1;
}

struct pg_handle {
    int n = 0, c0 = 0, device = 0;
    long long ldw = 0;  // padded leading dimension of d / wy (multiple of kTile, zero-filled)
    long long ldx = 0;  // padded row length of the rotated genotype block (multiple of 16, zero-filled)
    int engine = PG_REML_AUTO;  // REML stage engine (pg_set_reml_engine / env PG_REML_ENGINE)
    int scan_mode = PG_SCAN_WALD;  // pg_set_scan_mode
    int sm_count = 148;
    std::string err;
    cudaStream_t compute = nullptr, copy = nullptr;
    // scan pipeline: rotation GEMMs on `compute`, int8 recombination on `cmb`, compression + optimiser on `aux`,
    // so that block b's REML stage (FP64 pipes) runs under block b+1's rotation (int8 tensor pipe)
    cudaStream_t aux = nullptr, cmb = nullptr;
    cudaEvent_t ev_xr_ready[2] = {nullptr, nullptr}, ev_xr_free[2] = {nullptr, nullptr}, ev_aux_done = nullptr;
    cudaEvent_t ev_pv[2] = {nullptr, nullptr};   // brackets the scan's single p-value launch
    bool defer_pvalues = false;                  // set by pg_scan: one pvalue_kernel launch after the last block
    cudaEvent_t ev_z_ready[2] = {nullptr, nullptr}, ev_z_free[2] = {nullptr, nullptr};
    bool overlap = false;
    // per-call scratch kept across calls (grow-only): device result arrays, pinned host staging, timing events
    double* res_d = nullptr;   // [6][cap] doubles
    int* res_i = nullptr;      // [3][cap] ints
    char* res_host = nullptr;  // pinned: 6*8*cap + 3*4*cap bytes
    size_t res_cap = 0;
    std::vector<cudaEvent_t> ev_pool;
    // pageable host genotypes: blocks are packed into pinned bounce buffers by a few host threads while the GPU
    // works on the previous block (a pageable cudaMemcpy2D runs at ~10 GB/s, the bounce path at host-memcpy speed)
    char* bounce[2] = {nullptr, nullptr};
    size_t bounce_bytes = 0;
    cudaEvent_t ev_bounce_done[2] = {nullptr, nullptr};
    cudaStream_t own_compute = nullptr;  // compute defaults to this; pg_set_stream can point it at a caller stream
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t solver = nullptr;
    // eigen
    double* U = nullptr;  // n x n
    int u_op_t = 1;       // 1: stored column-major U (rotation uses U^T = op T); 0: stored buffer is U^T column-major
    bool have_U = false, have_d = false, have_design = false;
    double* d = nullptr;  // n
    // design
    double* wy = nullptr;  // n x (c0+1) column-major, rotated
    bool rotated_inputs = false;
    // tables
    double *fixtab = nullptr, *itab = nullptr, *basis = nullptr, *lambdas = nullptr;
    TriAB* tri_ab = nullptr;
    Tables tab{};
    // scan workspace
    int rotation = PG_ROT_AUTO;
    long long block_snps_opt = 0;
    long long blk = 0;  // allocated block size (SNPs)
    void* stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    double *xf = nullptr, *xr[2] = {nullptr, nullptr};
    size_t xbuf_elems = 0;
    const double* last_xr = nullptr;
    unsigned long long* counter = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    RotWorkspace rot;
    long long last_block_count = 0, last_block_row0 = 0;
    // internal eigen order: ascending d.  perm[l] = caller's index of the l-th smallest eigenvalue
    std::vector<int> perm;
    bool perm_identity = true;
    int* perm_dev = nullptr;
    // eigenvalue-space compression (compress_plan.h / compress.cuh)
    CompressPlan hplan;
    DevPlan plan;
    double* Z[2] = {nullptr, nullptr};  // [blk][k1p][Kcp], double-buffered like xr
    size_t z_elems = 0;
    double* F[2] = {nullptr, nullptr};  // [kNumFixed][zrows][ldF] double2: x rows of the fixed-lambda evaluations, per Z buffer
    double* FX[2] = {nullptr, nullptr}; // [kNumFixed][kFxOut][ldF] results of the fixed-lambda evaluations (fixed_phase_kernel)
    size_t f_elems = 0, fx_elems = 0;
    long long ldF = 0;                  // SNP stride of F / FX: the block size rounded up to 32
    int k1p = 0;          // c0+2 rounded up to a multiple of 4
    // several phenotypes on one eigen-system (pg_set_design_multi): each has its own rotated [W0, y] and lambda tables;
    // slot 0 owns the buffers allocated with the handle, `activate` points the handle's working fields at a slot
    // before its kernels are launched (launches copy the pointers, so slots can alternate per block)
    struct DesignSlot {
        double *wy, *fixtab, *itab, *fix2, *itab2;
        bool null_valid = false;                 // null-model ML fit of this phenotype (pg_null_model), cached per design
        double null_vals[4] = {0, 0, 0, 0};      // lambda_null, tau_null, l_null, status
    };
    int active_slot = 0;
    std::vector<DesignSlot> slots;
    int q = 1;
    double* design_raw = nullptr;   // staging of the unrotated [W, y] columns (pg_set_design), n x (c0+1)
    double* wy_all = nullptr;   // q > 1: rotated [W0, y_0 .. y_{q-1}], the linear columns of the shared compression
    int wy_all_cols = 0;
    int z_rows = 0;             // slab rows per SNP the Z buffers were last laid out for (k1p - 1 + q)
    // tables as one DGEMM per design (reml_kernels.cuh): powers of 1/(lambda d + 1) per eigen-system, pair products and the
    // product matrix per design
    double *hpow = nullptr, *hscal = nullptr, *pairp = nullptr, *tabc = nullptr;
    long long ldh = 0;
    bool hpow_valid = false;
    // table-2 rows (covariate levels eliminated per table lambda, pg_eval.cuh)
    double *fix2 = nullptr, *itab2 = nullptr, *t2work = nullptr;
    Tables2 tab2{};
    // Moments straight from the fused rotation (rotate_i8_tc2.cuh, FUSE).  Per eigen-system: which x^2 piece / node every
    // eigen-index feeds and the piece ranges of the segments.  Per design: G = U V (one column per COMPRESS node and linear
    // column), its digit planes, scales, column sums and slab offsets.
    int fuse_mode = -1;   // pg_set_moment_fusion: -1 automatic (fuse_wanted), 0 never, 1 wherever the engine allows it
    struct Fused {
        int2* einfo = nullptr;
        tc2::SegRed* segs = nullptr;
        int nsegs = 0, npieces = 0, cnodes = 0;       // cnodes: nodes of COMPRESS segments
        std::vector<Segment> csegs;                   // the COMPRESS segments, eigen order
        bool valid = false;                           // the per-design part below matches the current design
        int klin = 0, gcols = 0, g_tiles = 0;
        double *G = nullptr, *vdc = nullptr, *gscale = nullptr, *g1 = nullptr;
        size_t g_elems = 0, vdc_elems = 0;
        int8_t* planes_g = nullptr;
        size_t planes_bytes = 0;
        int *exps_g = nullptr, *goff = nullptr, *jrow = nullptr;
        int col_cap = 0, jrow_cap = 0;
        double* P2 = nullptr;
        size_t p2_elems = 0;
        long long ldp = 0;
        float build_ms = 0.f;
    } fz;
};

static void activate(pg_handle* h, int ph);

static void free_fused(pg_handle* h)
{
    pg_handle::Fused& F = h->fz;
    void* bufs[] = {F.einfo, F.segs, F.G, F.vdc, F.gscale, F.g1, F.planes_g, F.exps_g, F.goff, F.jrow, F.P2};
    for (void* b : bufs)
        if (b) cudaFree(b);
    F = pg_handle::Fused{};
}

static void free_plan(pg_handle* h)
{
    free_fused(h);
    DevPlan& P = h->plan;
    void* bufs[] = {P.nodes, P.H, P.Lw, P.seg_kq, P.V, P.items, P.copy_l, P.copy_node, h->Z[0], h->Z[1], h->F[0], h->F[1],
                    h->FX[0], h->FX[1]};
    for (void* b : bufs)
        if (b) cudaFree(b);
    P = DevPlan{};
    h->Z[0] = h->Z[1] = nullptr;
    h->F[0] = h->F[1] = nullptr;
    h->FX[0] = h->FX[1] = nullptr;
    h->fx_elems = 0;
    h->z_elems = 0;
    h->f_elems = 0;
    h->z_rows = 0;
}

static int fail(pg_handle* h, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(h, PG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CKB(call)                                                                                     \
    do {                                                                                              \
        cublasStatus_t s_ = (call);                                                                   \
        if (s_ != CUBLAS_STATUS_SUCCESS)                                                              \
            return fail(h, PG_ERR_CUBLAS, "%s:%d %s -> cublas status %d", __FILE__, __LINE__, #call, (int)s_); \
    } while (0)
#define CKS(call)                                                                                     \
    do {                                                                                              \
        cusolverStatus_t s_ = (call);                                                                 \
        if (s_ != CUSOLVER_STATUS_SUCCESS)                                                            \
            return fail(h, PG_ERR_CUSOLVER, "%s:%d %s -> cusolver status %d", __FILE__, __LINE__, #call, (int)s_); \
    } while (0)

// Scope guards for temporaries of one call: error paths return through the CK macros, the destructors release.
struct DevMem {
    void* p = nullptr;
    DevMem() = default;
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
    ~DevMem() { if (p) cudaFree(p); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
    void* release() { void* q = p; p = nullptr; return q; }
};
struct EventGuard {
    cudaEvent_t e = nullptr;
    EventGuard() = default;
    EventGuard(const EventGuard&) = delete;
    EventGuard& operator=(const EventGuard&) = delete;
    ~EventGuard() { if (e) cudaEventDestroy(e); }
};
struct SolverParamsGuard {
    cusolverDnParams_t p = nullptr;
    ~SolverParamsGuard() { if (p) cusolverDnDestroyParams(p); }
};

extern "C" int pg_abi_version(void) { return PG_ABI_VERSION; }
extern "C" int pg_rotation_planes(void) { return kSlices; }

extern "C" int pg_device_count(int* count)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (count) *count = (e == cudaSuccess) ? c : 0;
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        return PG_ERR_NO_DEVICE;
    }
    return PG_OK;
}

extern "C" const char* pg_last_error(const pg_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int free_all(pg_handle* h)
{
    cudaSetDevice(h->device);
    if (!h->slots.empty()) {
        activate(h, 0);   // the handle's own fields point at slot 0 again before they are freed below
        for (size_t ph = 1; ph < h->slots.size(); ++ph) {
            void* bufs[] = {h->slots[ph].wy, h->slots[ph].fixtab, h->slots[ph].itab, h->slots[ph].fix2, h->slots[ph].itab2};
            for (void* b : bufs)
                if (b) cudaFree(b);
        }
    }
    if (h->wy_all) cudaFree(h->wy_all);
    h->wy_all = nullptr;
    if (h->design_raw) cudaFree(h->design_raw);
    h->design_raw = nullptr;
    rot_free(&h->rot);
    for (int s = 0; s < 2; ++s) {
        if (h->stage[s]) cudaFree(h->stage[s]);
        if (h->ev_ready[s]) cudaEventDestroy(h->ev_ready[s]);
        if (h->ev_free[s]) cudaEventDestroy(h->ev_free[s]);
    }
    free_plan(h);
    for (int s = 0; s < 2; ++s) {
        if (h->ev_xr_ready[s]) cudaEventDestroy(h->ev_xr_ready[s]);
        if (h->ev_xr_free[s]) cudaEventDestroy(h->ev_xr_free[s]);
        if (h->ev_z_ready[s]) cudaEventDestroy(h->ev_z_ready[s]);
        if (h->ev_z_free[s]) cudaEventDestroy(h->ev_z_free[s]);
    }
    if (h->ev_aux_done) cudaEventDestroy(h->ev_aux_done);
    for (int s = 0; s < 2; ++s)
        if (h->ev_pv[s]) cudaEventDestroy(h->ev_pv[s]);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    h->ev_pool.clear();
    if (h->res_d) cudaFree(h->res_d);
    if (h->res_i) cudaFree(h->res_i);
    if (h->res_host) cudaFreeHost(h->res_host);
    for (int s = 0; s < 2; ++s) {
        if (h->bounce[s]) cudaFreeHost(h->bounce[s]);
        if (h->ev_bounce_done[s]) cudaEventDestroy(h->ev_bounce_done[s]);
    }
    if (h->aux) cudaStreamDestroy(h->aux);
    if (h->cmb) cudaStreamDestroy(h->cmb);
    void* bufs[] = {h->U, h->d, h->wy, h->fixtab, h->itab, h->basis, h->lambdas, h->tri_ab, h->xf, h->xr[0], h->xr[1], h->counter,
                    h->perm_dev, h->fix2, h->itab2, h->t2work, h->hpow, h->hscal, h->pairp, h->tabc};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (h->blas) cublasDestroy(h->blas);
    if (h->solver) cusolverDnDestroy(h->solver);
    if (h->own_compute) cudaStreamDestroy(h->own_compute);
    if (h->copy) cudaStreamDestroy(h->copy);
    return 0;
}

extern "C" int pg_destroy(pg_handle* h)
{
    if (!h) return PG_OK;
    free_all(h);
    delete h;
    return PG_OK;
}

extern "C" int pg_create(int n, int c0, int device, pg_handle** out)
{
    if (!out) return fail(nullptr, PG_ERR_ARG, "pg_create: out is NULL");
    *out = nullptr;
    if (n < 2 || c0 < 0 || c0 + 2 > kMaxCols || n - c0 - 1 < 1)
        return fail(nullptr, PG_ERR_ARG, "pg_create: need n >= 2, 0 <= c0 <= %d, n - c0 - 1 >= 1 (got n=%d c0=%d)",
                    kMaxCols - 2, n, c0);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, PG_ERR_NO_DEVICE, "pg_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, PG_ERR_ARG, "pg_create: device %d of %d", device, count);
    pg_handle* h = new pg_handle;
    h->n = n; h->c0 = c0; h->device = device;
    h->ldw = ((long long)n + kTile - 1) / kTile * kTile;
    h->ldx = ((long long)n + 31) / 32 * 32;
    if (const char* e = getenv("PG_REML_ENGINE")) {
        if (!strcmp(e, "stream")) h->engine = PG_REML_STREAM;
        else if (!strcmp(e, "warp")) h->engine = PG_REML_WARP;
        else if (!strcmp(e, "compressed")) h->engine = PG_REML_COMPRESSED;
    }
    int rc = [&]() -> int {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(h, PG_ERR_NO_DEVICE, "pg_create: device %d is sm_%d%d; this build targets sm_100a only", device,
                        prop.major, prop.minor);
        h->sm_count = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&h->own_compute, cudaStreamNonBlocking));
        h->compute = h->own_compute;
        CK(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
        {
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = highest priority
            // LOW priority: when block b's compression ends, the rotation clusters of block b+1 must be placed first
            // (they need 190 KB of shared memory per SM); the optimiser CTAs of block b then fill the room left beside them
            CK(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, lo));
            CK(cudaStreamCreateWithPriority(&h->cmb, cudaStreamNonBlocking, hi));
            for (int s = 0; s < 2; ++s) {
                CK(cudaEventCreateWithFlags(&h->ev_xr_ready[s], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_xr_free[s], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_z_ready[s], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_z_free[s], cudaEventDisableTiming));
            }
            CK(cudaEventCreateWithFlags(&h->ev_aux_done, cudaEventDisableTiming));
            // PG_OVERLAP=1: the optimiser kernel of block b runs on the auxiliary stream under the fused rotation of
            // block b+1.  Measured in the sustained bench: 67.1 vs 65.7 ms per step -- the step is power-capped (~900 W,
            // 1.7 GHz), so overlapping pipes does not buy time; off by default, stage timings stay disjoint.
            h->overlap = getenv("PG_OVERLAP") && atoi(getenv("PG_OVERLAP")) != 0;
        }
        CKB(cublasCreate(&h->blas));
        CKB(cublasSetStream(h->blas, h->compute));
        CKS(cusolverDnCreate(&h->solver));
        CKS(cusolverDnSetStream(h->solver, h->compute));
        for (int s = 0; s < 2; ++s) {
            CK(cudaEventCreateWithFlags(&h->ev_ready[s], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_free[s], cudaEventDisableTiming));
        }
        const int k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2, NF = 3 * T0 + 3;
        CK(cudaMalloc(&h->d, sizeof(double) * h->ldw));
        CK(cudaMalloc(&h->wy, sizeof(double) * (size_t)h->ldw * k0));
        CK(cudaMemset(h->d, 0, sizeof(double) * h->ldw));
        CK(cudaMemset(h->wy, 0, sizeof(double) * (size_t)h->ldw * k0));
        CK(cudaMalloc(&h->fixtab, sizeof(double) * (size_t)kNumFixed * NF));
        CK(cudaMalloc(&h->itab, sizeof(double) * (size_t)kNumIntervals * kNodes * NF));
        CK(cudaMalloc(&h->basis, sizeof(double) * kNodes * kNodes));
        CK(cudaMalloc(&h->lambdas, sizeof(double) * kNumTableRows));
        CK(cudaMalloc(&h->tri_ab, sizeof(TriAB) * kMaxTri));
        CK(cudaMalloc(&h->counter, 2 * sizeof(unsigned long long)));  // one work-queue counter per block parity
        std::vector<TriAB> tab(kMaxTri);
        fill_tri_ab(tab.data());
        CK(cudaMemcpy(h->tri_ab, tab.data(), sizeof(TriAB) * kMaxTri, cudaMemcpyHostToDevice));
        std::vector<double> basis(kNodes * kNodes), lams(kNumTableRows);
        fill_basis(basis.data());
        for (int r = 0; r < kNumTableRows; ++r) lams[r] = table_lambda(r);
        CK(cudaMemcpy(h->basis, basis.data(), sizeof(double) * basis.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->lambdas, lams.data(), sizeof(double) * lams.size(), cudaMemcpyHostToDevice));
        const int NF2 = t2_nf(c0);
        CK(cudaMalloc(&h->fix2, sizeof(double) * (size_t)kNumFixed * NF2));
        CK(cudaMalloc(&h->itab2, sizeof(double) * (size_t)kNumIntervals * kNodes * NF2));
        CK(cudaMalloc(&h->t2work, sizeof(double) * (size_t)kNumTableRows * 3 * T0));
        h->tab2.c0 = c0; h->tab2.Tp = t2_pairs(c0); h->tab2.NF2 = NF2;
        h->tab2.fix2 = h->fix2; h->tab2.itab2 = h->itab2; h->tab2.basis = h->basis;
        h->k1p = (c0 + 2 + 3) / 4 * 4;
        if (getenv("PG_FUSE_MOMENTS")) h->fuse_mode = std::max(-1, std::min(1, atoi(getenv("PG_FUSE_MOMENTS"))));   // pg_set_moment_fusion
        h->tab.c0 = c0; h->tab.k0 = k0; h->tab.T0 = T0; h->tab.NF = NF;
        h->tab.fixtab = h->fixtab; h->tab.itab = h->itab; h->tab.basis = h->basis; h->tab.tri_ab = h->tri_ab;
        return PG_OK;
    }();
    if (rc != PG_OK) {
        g_create_error = h->err;
        free_all(h);
        delete h;
        return rc;
    }
    *out = h;
    return PG_OK;
}

static int ensure_U(pg_handle* h)
{
    if (!h->U) CK(cudaMalloc(&h->U, sizeof(double) * (size_t)h->n * h->n));
    return PG_OK;
}


// U columns (eigenvectors) into ascending-eigenvalue order.  cols_contig: eigenvector i at src + i*n.
__global__ void permute_u_kernel(const double* __restrict__ src, double* __restrict__ dst, int n,
                                 const int* __restrict__ perm, int cols_contig)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int a = (int)(idx / n), b = (int)(idx % n);
    // cols_contig: dst[i*n + j] = src[perm[i]*n + j] (a = i, b = j); else dst[j*n + i] = src[j*n + perm[i]] (a = j, b = i)
    dst[idx] = cols_contig ? src[(size_t)perm[a] * n + b] : src[(size_t)a * n + perm[b]];
}

// Work items and the V operand of the compression for `klin` linear columns [W0, y_0 .. y_{q-1}] (klin = c0 + q):
// one item per COMPRESS segment and group of kJGroup columns.  The V contents are written by build_v_kernel.
static int build_groups(pg_handle* h, int klin)
{
    const CompressPlan& H = h->hplan;
    DevPlan& P = h->plan;
    if (P.items) cudaFree(P.items);
    if (P.V) cudaFree(P.V);
    P.items = nullptr; P.V = nullptr; P.nitems = 0;
    P.klin = klin;
    P.ngroups = (klin + kJGroup - 1) / kJGroup;
    P.vpitch = P.ngroups * kGroupCols;
    std::vector<int> order;
    for (int si = 0; si < (int)H.segs.size(); ++si)
        if (H.segs[si].type == kSegCompress) order.push_back(si);
    // longest segments first: the tail of the grid is made of short work items
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (H.segs[a].l1 - H.segs[a].l0) > (H.segs[b].l1 - H.segs[b].l0);
    });
    std::vector<CompItem> items;
    for (int si : order) {
        const Segment& sg = H.segs[si];
        for (int g = 0; g < P.ngroups; ++g) {
            CompItem it;
            it.l0 = sg.l0; it.l1 = sg.l1; it.kb = sg.kb; it.kq = sg.kq;
            it.j0 = g * kJGroup; it.nj = std::min<int>(kJGroup, klin - it.j0); it.vcol0 = g * kGroupCols; it.pad = 0;
            items.push_back(it);
        }
    }
    P.nitems = (int)items.size();
    if (P.nitems) {
        CK(cudaMalloc(&P.items, sizeof(CompItem) * items.size()));
        CK(cudaMemcpy(P.items, items.data(), sizeof(CompItem) * items.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&P.V, sizeof(double) * (size_t)P.npad16 * P.vpitch));
        CK(cudaMemset(P.V, 0, sizeof(double) * (size_t)P.npad16 * P.vpitch));
    }
    return PG_OK;
}

// Builds the compression plan for the handle's (ascending, clipped) eigenvalues and uploads it.
static int upload_plan(pg_handle* h, const std::vector<double>& d_sorted)
{
    const int n = h->n, c0 = h->c0, k0 = c0 + 1;
    free_plan(h);
    build_compress_plan(d_sorted.data(), n, &h->hplan);
    const CompressPlan& H = h->hplan;
    DevPlan& P = h->plan;
    P.Kc = H.Kc;
    P.Kcp = std::max(32, (H.Kc + 31) / 32 * 32);
    std::vector<double> nodes(P.Kcp, 0.0);
    std::copy(H.nodes.begin(), H.nodes.end(), nodes.begin());
    CK(cudaMalloc(&P.nodes, sizeof(double) * P.Kcp));
    CK(cudaMemcpy(P.nodes, nodes.data(), sizeof(double) * P.Kcp, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&P.H, sizeof(double) * (size_t)P.Kcp * kFxCols));
    build_h_kernel<<<(P.Kcp * kFxCols + 127) / 128, 128, 0, h->compute>>>(P.nodes, P.Kcp, P.H);
    CK(cudaGetLastError());
    CK(cudaMalloc(&P.Lw, sizeof(double) * (size_t)n * kCq));
    CK(cudaMemcpy(P.Lw, H.Lw.data(), sizeof(double) * (size_t)n * kCq, cudaMemcpyHostToDevice));
    std::vector<int> seg_kq(n, 0), copy_l, copy_node;
    P.npad16 = (n + 31) / 32 * 32;
    for (int si = 0; si < (int)H.segs.size(); ++si) {
        const Segment& sg = H.segs[si];
        if (sg.type == kSegCompress) {
            for (int l = sg.l0; l < sg.l1; ++l) seg_kq[l] = sg.kq;
        } else {
            for (int l = sg.l0; l < sg.l1; ++l) { copy_l.push_back(l); copy_node.push_back(sg.kb + (l - sg.l0)); }
        }
    }
    CK(cudaMalloc(&P.seg_kq, sizeof(int) * n));
    CK(cudaMemcpy(P.seg_kq, seg_kq.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    {
        // fused moments: x^2 piece (COMPRESS rows) or node (COPY rows) of every eigen-index, piece ranges per segment
        pg_handle::Fused& F = h->fz;
        FusedPlan fp;
        build_fused_plan(H, tc2::kTileEig, tc2::kPieceEig, &fp);
        static_assert(sizeof(FusedRow) == sizeof(int2) && sizeof(FusedSeg) == sizeof(tc2::SegRed), "plain int records");
        F.csegs = fp.csegs;
        F.npieces = fp.npieces; F.cnodes = fp.cnodes;
        F.nsegs = (int)fp.segs.size();
        CK(cudaMalloc(&F.einfo, sizeof(int2) * fp.rows.size()));
        CK(cudaMemcpy(F.einfo, fp.rows.data(), sizeof(int2) * fp.rows.size(), cudaMemcpyHostToDevice));
        if (F.nsegs) {
            CK(cudaMalloc(&F.segs, sizeof(tc2::SegRed) * F.nsegs));
            CK(cudaMemcpy(F.segs, fp.segs.data(), sizeof(tc2::SegRed) * F.nsegs, cudaMemcpyHostToDevice));
        }
        F.valid = false;
    }
    int rcg = build_groups(h, k0);
    if (rcg) return rcg;
    P.ncopy = (int)copy_l.size();
    if (P.ncopy) {
        CK(cudaMalloc(&P.copy_l, sizeof(int) * P.ncopy));
        CK(cudaMalloc(&P.copy_node, sizeof(int) * P.ncopy));
        CK(cudaMemcpy(P.copy_l, copy_l.data(), sizeof(int) * P.ncopy, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(P.copy_node, copy_node.data(), sizeof(int) * P.ncopy, cudaMemcpyHostToDevice));
    }
    return PG_OK;
}

// Brings the handle's eigen-system into ascending-eigenvalue order (h->d holds the clipped eigenvalues in the
// caller's order, h->U the eigenvectors when have_U) and builds the compression plan.
static int canonicalise_eigen(pg_handle* h)
{
    const int n = h->n;
    h->hpow_valid = false;   // the table powers belong to the eigenvalues
    std::vector<double> dh(n);
    CK(cudaMemcpyAsync(dh.data(), h->d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    h->perm.resize(n);
    for (int l = 0; l < n; ++l) h->perm[l] = l;
    std::stable_sort(h->perm.begin(), h->perm.end(), [&](int a, int b) { return dh[a] < dh[b]; });
    h->perm_identity = true;
    for (int l = 0; l < n; ++l)
        if (h->perm[l] != l) { h->perm_identity = false; break; }
    std::vector<double> ds(n);
    for (int l = 0; l < n; ++l) ds[l] = dh[h->perm[l]];
    if (!h->perm_dev) CK(cudaMalloc(&h->perm_dev, sizeof(int) * n));
    CK(cudaMemcpy(h->perm_dev, h->perm.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    if (!h->perm_identity) {
        CK(cudaMemcpy(h->d, ds.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        if (h->have_U) {
            DevMem tmp;
            CK(cudaMalloc(&tmp.p, sizeof(double) * (size_t)n * n));
            const size_t total = (size_t)n * n;
            permute_u_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->compute>>>(h->U, tmp.as<double>(), n, h->perm_dev, h->u_op_t);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(h->compute));
            cudaFree(h->U);
            h->U = static_cast<double*>(tmp.release());
        }
    }
    return upload_plan(h, ds);
}

static int set_kinship_common(pg_handle* h, const double* K, cudaMemcpyKind kind, double* d_out_host, float* eig_ms);

extern "C" int pg_set_kinship(pg_handle* h, const double* K_host, double* d_out_host, float* eig_ms)
{
    return set_kinship_common(h, K_host, cudaMemcpyHostToDevice, d_out_host, eig_ms);
}

// K already on the device (pg_grm)
static int set_kinship_device(pg_handle* h, const double* K_dev, double* d_out_host, float* eig_ms)
{
    return set_kinship_common(h, K_dev, cudaMemcpyDeviceToDevice, d_out_host, eig_ms);
}

static int set_kinship_common(pg_handle* h, const double* K_host, cudaMemcpyKind kind, double* d_out_host, float* eig_ms)
{
    if (!h || !K_host) return fail(h, PG_ERR_ARG, "pg_set_kinship: NULL argument");
    CK(cudaSetDevice(h->device));
    const int n = h->n;
    int rc = ensure_U(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->U, K_host, sizeof(double) * (size_t)n * n, kind, h->compute));
    EventGuard g0, g1;
    CK(cudaEventCreate(&g0.e));
    CK(cudaEventCreate(&g1.e));
    const cudaEvent_t e0 = g0.e, e1 = g1.e;
    SolverParamsGuard pg;
    CKS(cusolverDnCreateParams(&pg.p));
    cusolverDnParams_t params = pg.p;
    size_t wdev = 0, whost = 0;
    CKS(cusolverDnXsyevd_bufferSize(h->solver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F,
                                    h->U, n, CUDA_R_64F, h->d, CUDA_R_64F, &wdev, &whost));
    DevMem dwork_g, info_g;
    CK(cudaMalloc(&dwork_g.p, wdev ? wdev : 8));
    CK(cudaMalloc(&info_g.p, sizeof(int)));
    void* dwork = dwork_g.p;
    int* info = info_g.as<int>();
    std::vector<char> hwork(whost ? whost : 8);
    CK(cudaEventRecord(e0, h->compute));
    cusolverStatus_t st = cusolverDnXsyevd(h->solver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                           CUDA_R_64F, h->U, n, CUDA_R_64F, h->d, CUDA_R_64F, dwork, wdev,
                                           hwork.data(), whost, info);
    CK(cudaEventRecord(e1, h->compute));
    int hinfo = -1;
    cudaError_t ce = cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, h->compute);
    cudaError_t se = cudaStreamSynchronize(h->compute);
    if (st != CUSOLVER_STATUS_SUCCESS) return fail(h, PG_ERR_CUSOLVER, "cusolverDnXsyevd status %d", (int)st);
    if (ce != cudaSuccess || se != cudaSuccess)
        return fail(h, PG_ERR_CUDA, "syevd: %s", cudaGetErrorString(ce != cudaSuccess ? ce : se));
    if (hinfo != 0) return fail(h, PG_ERR_CUSOLVER, "syevd did not converge (info=%d)", hinfo);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (eig_ms) *eig_ms = ms;
    clip_nonneg_kernel<<<(n + 255) / 256, 256, 0, h->compute>>>(h->d, n);
    CK(cudaGetLastError());
    if (d_out_host) CK(cudaMemcpyAsync(d_out_host, h->d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    h->u_op_t = 1; h->have_U = true; h->have_d = true; h->rotated_inputs = false; h->have_design = false;
    rot_invalidate(&h->rot);
    return canonicalise_eigen(h);
}

static int set_eigen_common(pg_handle* h, const double* U, int u_row_major, const double* d, cudaMemcpyKind kind)
{
    if (!h || !d) return fail(h, PG_ERR_ARG, "pg_set_eigen: NULL argument");
    CK(cudaSetDevice(h->device));
    const int n = h->n;
    if (U) {
        int rc = ensure_U(h);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->U, U, sizeof(double) * (size_t)n * n, kind, h->compute));
        // a row-major U buffer read as column-major is U^T: the rotation then uses it untransposed
        h->u_op_t = u_row_major ? 0 : 1;
        h->have_U = true;
    } else {
        h->have_U = false;
    }
    CK(cudaMemcpyAsync(h->d, d, sizeof(double) * n, kind, h->compute));
    clip_nonneg_kernel<<<(n + 255) / 256, 256, 0, h->compute>>>(h->d, n);  // lmm/lmm.py:157 / :166
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->compute));
    h->have_d = true; h->have_design = false;
    rot_invalidate(&h->rot);
    return canonicalise_eigen(h);
}

extern "C" int pg_set_eigen(pg_handle* h, const double* U_host, int u_row_major, const double* d_host)
{
    return set_eigen_common(h, U_host, u_row_major, d_host, cudaMemcpyHostToDevice);
}

extern "C" int pg_set_eigen_device(pg_handle* h, const double* U_dev, int u_row_major, const double* d_dev)
{
    return set_eigen_common(h, U_dev, u_row_major, d_dev, cudaMemcpyDeviceToDevice);
}

extern "C" int pg_get_eigen_device(pg_handle* h, double* U_dev_out, double* d_dev_out)
{
    if (!h || !h->have_U || !h->have_d) return fail(h, PG_ERR_ARG, "pg_get_eigen_device: no eigendecomposition held");
    if (!h->u_op_t) return fail(h, PG_ERR_ARG, "pg_get_eigen_device: U is held transposed");
    CK(cudaSetDevice(h->device));
    const int n = h->n;
    if (U_dev_out)
        CK(cudaMemcpyAsync(U_dev_out, h->U, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, h->compute));
    if (d_dev_out) CK(cudaMemcpyAsync(d_dev_out, h->d, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    return PG_OK;
}

// Hands the eigen-system of `src` (U, clipped d, already in ascending order) to `dst` device-to-device: a caller that
// keeps one decomposition for many designs (lmm.factorize) re-creates the handle when the covariate count changes.
extern "C" int pg_copy_eigen(pg_handle* h, const pg_handle* src)
{
    if (!h || !src) return fail(h, PG_ERR_ARG, "pg_copy_eigen: NULL argument");
    if (h == src) return PG_OK;
    if (h->n != src->n) return fail(h, PG_ERR_ARG, "pg_copy_eigen: n differs (%d vs %d)", h->n, src->n);
    if (!src->have_d) return fail(h, PG_ERR_ARG, "pg_copy_eigen: the source handle holds no eigen-system");
    CK(cudaSetDevice(src->device));
    CK(cudaStreamSynchronize(src->compute));
    CK(cudaSetDevice(h->device));
    if (h->device != src->device) {
        // direct peer-to-peer copies (NVLink / NVSwitch) instead of staging through host memory, where the devices allow it
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, h->device, src->device) == cudaSuccess && can) {
            const cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0);
            if (pe != cudaSuccess) cudaGetLastError();   // cudaErrorPeerAccessAlreadyEnabled is fine; anything else falls back to staging
        } else {
            cudaGetLastError();
        }
    }
    const int n = h->n;
    if (src->have_U) {
        int rc = ensure_U(h);
        if (rc) return rc;
        if (h->device == src->device)
            CK(cudaMemcpyAsync(h->U, src->U, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, h->compute));
        else
            CK(cudaMemcpyPeerAsync(h->U, h->device, src->U, src->device, sizeof(double) * (size_t)n * n, h->compute));
        h->u_op_t = src->u_op_t;
        h->have_U = true;
    } else {
        h->have_U = false;
    }
    if (h->device == src->device)
        CK(cudaMemcpyAsync(h->d, src->d, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->compute));
    else
        CK(cudaMemcpyPeerAsync(h->d, h->device, src->d, src->device, sizeof(double) * n, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    h->have_d = true; h->have_design = false; h->rotated_inputs = false;
    rot_invalidate(&h->rot);
    // src is canonical (ascending); its caller-order permutation is kept so that rotated inputs / probes keep their meaning
    int rc = canonicalise_eigen(h);
    if (rc) return rc;
    h->perm = src->perm;
    h->perm_identity = src->perm_identity;
    CK(cudaMemcpy(h->perm_dev, h->perm.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    return PG_OK;
}

static void activate(pg_handle* h, int ph)
{
    const pg_handle::DesignSlot& S = h->slots[ph];
    h->active_slot = ph;
    h->wy = S.wy; h->fixtab = S.fixtab; h->itab = S.itab; h->fix2 = S.fix2; h->itab2 = S.itab2;
    h->tab.fixtab = S.fixtab; h->tab.itab = S.itab; h->tab2.fix2 = S.fix2; h->tab2.itab2 = S.itab2;
}

// makes sure slots 0..q-1 exist (slot 0 = the handle's own buffers)
static int ensure_slots(pg_handle* h, int q)
{
    const int c0 = h->c0, k0 = c0 + 1, T0 = k0 * (k0 + 1) / 2, NF = 3 * T0 + 3, NF2 = t2_nf(c0);
    if (h->slots.empty()) h->slots.push_back(pg_handle::DesignSlot{h->wy, h->fixtab, h->itab, h->fix2, h->itab2});
    while ((int)h->slots.size() < q) {
        pg_handle::DesignSlot S{nullptr, nullptr, nullptr, nullptr, nullptr};
        CK(cudaMalloc(&S.wy, sizeof(double) * (size_t)h->ldw * k0));
        CK(cudaMemset(S.wy, 0, sizeof(double) * (size_t)h->ldw * k0));
        CK(cudaMalloc(&S.fixtab, sizeof(double) * (size_t)kNumFixed * NF));
        CK(cudaMalloc(&S.itab, sizeof(double) * (size_t)kNumIntervals * kNodes * NF));
        CK(cudaMalloc(&S.fix2, sizeof(double) * (size_t)kNumFixed * NF2));
        CK(cudaMalloc(&S.itab2, sizeof(double) * (size_t)kNumIntervals * kNodes * NF2));
        h->slots.push_back(S);
    }
    return PG_OK;
}

// powers of 1/(lambda d + 1) for every table lambda and the design-independent scalars: once per eigen-system
static int ensure_hpow(pg_handle* h)
{
    if (h->hpow_valid) return PG_OK;
    const int n = h->n, R3 = 3 * kNumTableRows, T0 = h->tab.T0;
    h->ldh = h->ldw;
    if (!h->hpow) CK(cudaMalloc(&h->hpow, sizeof(double) * (size_t)h->ldh * R3));
    if (!h->hscal) CK(cudaMalloc(&h->hscal, sizeof(double) * 3 * kNumTableRows));
    if (!h->pairp) CK(cudaMalloc(&h->pairp, sizeof(double) * (size_t)h->ldh * T0));
    if (!h->tabc) CK(cudaMalloc(&h->tabc, sizeof(double) * (size_t)R3 * T0));
    build_hpow_kernel<<<dim3((unsigned)((h->ldh + 255) / 256), kNumTableRows), 256, 0, h->compute>>>(n, h->ldh, h->d, h->lambdas, h->hpow);
    CK(cudaGetLastError());
    build_hscal_kernel<<<kNumTableRows, 256, 0, h->compute>>>(n, h->d, h->lambdas, h->hscal);
    CK(cudaGetLastError());
    h->hpow_valid = true;
    return PG_OK;
}

static int build_tables(pg_handle* h)
{
    const int T0 = h->tab.T0;
    // PG_TABLES_GEMM=0: the per-row reduction kernel of round 1 (cross-check); memory: 24 n x 971 bytes for the powers
    static const bool use_gemm = !(getenv("PG_TABLES_GEMM") && atoi(getenv("PG_TABLES_GEMM")) == 0);
    if (use_gemm) {
        int rc = ensure_hpow(h);
        if (rc) return rc;
        const int n = h->n, R3 = 3 * kNumTableRows;
        pair_products_kernel<<<dim3((unsigned)((h->ldh + 255) / 256), T0), 256, 0, h->compute>>>(n, h->ldh, T0, h->wy, h->ldw,
                                                                                              h->tri_ab, h->pairp);
        CK(cudaGetLastError());
        const double one = 1.0, zero = 0.0;
        // C (R3 x T0) = Hpow^T (R3 x n) . P (n x T0), all column-major
        CKB(cublasDgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, R3, T0, n, &one, h->hpow, (int)h->ldh, h->pairp, (int)h->ldh, &zero,
                        h->tabc, R3));
        scatter_tables_kernel<<<kNumTableRows, 128, 0, h->compute>>>(T0, h->tab.NF, R3, h->tabc, h->hscal, h->fixtab, h->itab);
        CK(cudaGetLastError());
        // one CTA per table lambda; PG_ELIM_SERIAL=1: the one-thread-per-lambda kernel (same bits, 19 ms at c0 = 40)
        static const bool elim_serial = getenv("PG_ELIM_SERIAL") && atoi(getenv("PG_ELIM_SERIAL")) != 0;
        if (elim_serial)
            eliminate_tables_kernel<<<(kNumTableRows + 63) / 64, 64, 0, h->compute>>>(h->c0, h->tab.NF, h->tab2.NF2, h->fixtab,
                                                                                      h->itab, h->fix2, h->itab2, h->t2work);
        else
            eliminate_tables_cta_kernel<<<kNumTableRows, 128, 0, h->compute>>>(h->c0, h->tab.NF, h->tab2.NF2, h->fixtab, h->itab,
                                                                               h->fix2, h->itab2, h->t2work);
        CK(cudaGetLastError());
        return PG_OK;
    }
    const int nchunks = (T0 + kTablePairs - 1) / kTablePairs;
    dim3 grid(nchunks + 1, kNumTableRows);
    build_tables_kernel<<<grid, 256, 0, h->compute>>>(h->n, h->c0, h->d, h->wy, h->ldw, h->lambdas, h->fixtab, h->itab,
                                                      h->tri_ab);
    CK(cudaGetLastError());
    eliminate_tables_kernel<<<(kNumTableRows + 63) / 64, 64, 0, h->compute>>>(h->c0, h->tab.NF, h->tab2.NF2, h->fixtab, h->itab,
                                                                              h->fix2, h->itab2, h->t2work);
    CK(cudaGetLastError());
    return PG_OK;
}

static int set_design_active(pg_handle* h, const double* W_host, const double* y_host, int already_rotated, float* ms);

// (Re)builds the compression operand V for the handle's current phenotypes: linear columns [W0, y_0 .. y_{q-1}].
static int build_v(pg_handle* h, int q)
{
    const int n = h->n, c0 = h->c0, klin = c0 + q;
    h->fz.valid = false;   // G = U V belongs to the design
    if (h->plan.klin != klin) {
        int rc = build_groups(h, klin);
        if (rc) return rc;
    }
    const double* cols = h->slots[0].wy;   // q == 1: the slot's own [W0, y]
    if (q > 1) {
        if (h->wy_all_cols < klin) {
            if (h->wy_all) cudaFree(h->wy_all);
            h->wy_all = nullptr; h->wy_all_cols = 0;
            CK(cudaMalloc(&h->wy_all, sizeof(double) * (size_t)h->ldw * klin));
            h->wy_all_cols = klin;
        }
        CK(cudaMemcpyAsync(h->wy_all, h->slots[0].wy, sizeof(double) * (size_t)h->ldw * c0, cudaMemcpyDeviceToDevice,
                           h->compute));
        for (int ph = 0; ph < q; ++ph)
            CK(cudaMemcpyAsync(h->wy_all + (size_t)h->ldw * (c0 + ph), h->slots[ph].wy + (size_t)h->ldw * c0,
                               sizeof(double) * (size_t)h->ldw, cudaMemcpyDeviceToDevice, h->compute));
        cols = h->wy_all;
    }
    if (h->plan.nitems) {
        build_v_kernel<<<n, 128, 0, h->compute>>>(n, klin, h->plan.Lw, h->plan.seg_kq, cols, h->ldw, h->plan.V,
                                                  h->plan.vpitch, h->plan.ngroups);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->compute));
    return PG_OK;
}

// ---- moments straight from the fused rotation -------------------------------------------------------------------
static bool fuse_env_tc2()
{
    static const bool ok = !(getenv("PG_ROT_DEFAULT") && !strcmp(getenv("PG_ROT_DEFAULT"), "cublas")) &&
                           !(getenv("PG_TC_CLUSTER") && atoi(getenv("PG_TC_CLUSTER")) != 2);
    return ok;
}

// Whether scans of the current design take the linear moments from tiles of G = U V and the x^2 moments from the rotation
// epilogue (rotate_i8_tc2.cuh) instead of writing the rotated genotypes and compressing them.  The extra tensor work is
// (COMPRESS nodes x linear columns) / n of the rotation.  MEASURED (profiles/experiments_r02.md, one B200, n = 10 000, 131
// nodes, per 100 k SNPs): with 11 linear columns (c0 = 10) the fused kernel takes what rotation + compression take together
// (53.3-53.8 ms against 42.7 + 11.0; the step is power-capped and the tensor pipe, busy for the whole step, pulls the SM clock
// from 1.67 to 1.57 GHz), and end to end it loses the ~1 ms G costs per design; with 41 linear columns (c0 = 40) it wins 7 %
// (73.8 ms against 40.6 + 37.5: 1.23 M instead of 1.15 M SNPs/s).  The extra tiles grow with nodes x columns, the compression
// with the columns only (in groups of 11): at 206 nodes a 32-trait pass (42 columns) LOSES 7 % fused (85.9 ms against
// 40.2 + 38.7).  So the automatic mode fuses from 24 linear columns on (c0 >= 23, or several traits per pass) when the plan has
// at most 140 COMPRESS nodes; pg_set_moment_fusion(h, 1) / PG_FUSE_MOMENTS=1 fuses wherever possible.
static bool fuse_wanted(const pg_handle* h)
{
    if (h->fuse_mode == 0 || !fuse_env_tc2()) return false;
    if (!(h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED) || h->overlap) return false;
    if (!h->have_U || h->rotated_inputs || h->fz.nsegs == 0) return false;
    const int klin = h->c0 + h->q;
    const long long gcols = (long long)h->fz.cnodes * klin;
    if (h->fuse_mode == 1)
        return (h->rotation == PG_ROT_AUTO || h->rotation == PG_ROT_I8TC) && gcols <= 4LL * h->n + 4096;
    return h->rotation == PG_ROT_AUTO && klin >= 24 && h->n >= 4096 && h->fz.cnodes <= 140 && gcols <= h->n;
}

// G = U V for the current design, sliced into digit planes like U^T (slice_u_kernel), with its scales, column sums
// and the slab offset of every column.  Built lazily by the first scan that fuses; pg_set_design* invalidates it.
static int ensure_fused_operand(pg_handle* h)
{
    pg_handle::Fused& F = h->fz;
    if (F.valid) return PG_OK;
    CK(cudaSetDevice(h->device));
    const int n = h->n, c0 = h->c0, q = h->q, klin = c0 + q, Kcp = h->plan.Kcp;
    const int gcols = F.cnodes * klin, g_tiles = (gcols + tc2::kTileEig - 1) / tc2::kTileEig, npad_g = g_tiles * tc2::kTileEig;
    const int ldk = (n + 127) / 128 * 128;   // row pitch of the digit planes: rot_prepare_i8's
    const double* cols = q > 1 ? h->wy_all : h->slots[0].wy;
    EventGuard g0, g1;
    CK(cudaEventCreate(&g0.e));
    CK(cudaEventCreate(&g1.e));
    CK(cudaEventRecord(g0.e, h->compute));
    const size_t g_need = (size_t)n * std::max(gcols, 1), vdc_need = (size_t)n * kCq * klin;
    if (g_need > F.g_elems) {
        if (F.G) cudaFree(F.G);
        F.G = nullptr; F.g_elems = 0;
        CK(cudaMalloc(&F.G, sizeof(double) * g_need));
        F.g_elems = g_need;
    }
    if (vdc_need > F.vdc_elems) {
        if (F.vdc) cudaFree(F.vdc);
        F.vdc = nullptr; F.vdc_elems = 0;
        CK(cudaMalloc(&F.vdc, sizeof(double) * vdc_need));
        F.vdc_elems = vdc_need;
    }
    const size_t pl_need = (size_t)kSlices * npad_g * ldk;
    if (pl_need > F.planes_bytes) {
        if (F.planes_g) cudaFree(F.planes_g);
        F.planes_g = nullptr; F.planes_bytes = 0;
        CK(cudaMalloc(&F.planes_g, pl_need));
        F.planes_bytes = pl_need;
    }
    if (npad_g > F.col_cap) {
        void* old[] = {F.gscale, F.g1, F.exps_g, F.goff};
        for (void* b : old)
            if (b) cudaFree(b);
        F.gscale = F.g1 = nullptr; F.exps_g = F.goff = nullptr; F.col_cap = 0;
        CK(cudaMalloc(&F.gscale, sizeof(double) * npad_g));
        CK(cudaMalloc(&F.g1, sizeof(double) * npad_g));
        CK(cudaMalloc(&F.exps_g, sizeof(int) * npad_g));
        CK(cudaMalloc(&F.goff, sizeof(int) * npad_g));
        F.col_cap = npad_g;
    }
    if (klin > F.jrow_cap) {
        if (F.jrow) cudaFree(F.jrow);
        F.jrow = nullptr; F.jrow_cap = 0;
        CK(cudaMalloc(&F.jrow, sizeof(int) * klin));
        F.jrow_cap = klin;
    }
    build_vc_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)(kCq * klin)), 256, 0, h->compute>>>(
        n, klin, h->plan.Lw, h->plan.seg_kq, cols, h->ldw, F.vdc);
    CK(cudaGetLastError());
    std::vector<int> goff(npad_g, -1), jrow(klin);
    for (int j = 0; j < klin; ++j) jrow[j] = z_row(j, c0, h->k1p) * Kcp;
    const double one = 1.0, zero = 0.0;
    int col0 = 0;
    for (const Segment& sg : F.csegs) {
        const int len = sg.l1 - sg.l0, ncs = sg.kq * klin;
        // G[:, col0 : col0 + ncs] = U[:, l0:l1] . Vc[l0:l1, .]   (column-major; columns (j, k), node k fastest; a one-node
        // segment takes the k = 0 column of every j: leading dimension n kCq)
        const int ldb = sg.kq == kCq ? n : n * kCq;
        if (h->u_op_t)
            CKB(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, n, ncs, len, &one, h->U + (size_t)sg.l0 * n, n, F.vdc + sg.l0, ldb,
                            &zero, F.G + (size_t)col0 * n, n));
        else
            CKB(cublasDgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, n, ncs, len, &one, h->U + sg.l0, n, F.vdc + sg.l0, ldb, &zero,
                            F.G + (size_t)col0 * n, n));
        for (int j = 0; j < klin; ++j)
            for (int k = 0; k < sg.kq; ++k) goff[col0 + j * sg.kq + k] = jrow[j] + sg.kb + k;
        col0 += ncs;
    }
    CK(cudaMemsetAsync(F.gscale, 0, sizeof(double) * npad_g, h->compute));
    CK(cudaMemsetAsync(F.g1, 0, sizeof(double) * npad_g, h->compute));
    if (gcols) {
        slice_u_kernel<<<npad_g, 256, 0, h->compute>>>(F.G, 1, n, npad_g, ldk, F.planes_g, F.exps_g, gcols);
        CK(cudaGetLastError());
        tc::plane_scale_kernel<<<(gcols + 255) / 256, 256, 0, h->compute>>>(F.exps_g, gcols, F.gscale);
        CK(cudaGetLastError());
        column_sums_kernel<<<gcols, 256, 0, h->compute>>>(F.G, 1, n, F.g1);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(F.goff, goff.data(), sizeof(int) * npad_g, cudaMemcpyHostToDevice, h->compute));
    CK(cudaMemcpyAsync(F.jrow, jrow.data(), sizeof(int) * klin, cudaMemcpyHostToDevice, h->compute));
    CK(cudaEventRecord(g1.e, h->compute));
    CK(cudaStreamSynchronize(h->compute));   // goff / jrow are local buffers
    cudaEventElapsedTime(&F.build_ms, g0.e, g1.e);
    F.klin = klin; F.gcols = gcols; F.g_tiles = g_tiles;
    F.valid = true;
    return PG_OK;
}

extern "C" int pg_set_moment_fusion(pg_handle* h, int mode)
{
    if (!h || mode < -1 || mode > 1) return fail(h, PG_ERR_ARG, "pg_set_moment_fusion: mode %d", mode);
    h->fuse_mode = mode;
    return PG_OK;
}

extern "C" int pg_probe_rotation_launches(int n, int64_t mb, int64_t setaside_bytes, int64_t max_window_bytes, int sm_count,
                                          int32_t* group_tiles)
{
    if (n < 1 || mb < 0 || setaside_bytes < 0 || max_window_bytes < 0 || sm_count < 2) return PG_ERR_ARG;
    const int ldk = (n + 127) / 128 * 128, eig_tiles = (n + tc2::kTileEig - 1) / tc2::kTileEig;
    const int snp_tiles = (int)((mb + tc2::kClusterSnps - 1) / tc2::kClusterSnps);
    const size_t setaside = std::min<size_t>((size_t)setaside_bytes, (size_t)tc2::kPersistMB << 20);
    const int g = tc2::persist_planes()
                      ? tc2::persist_group_tiles(snp_tiles, ldk, setaside, (size_t)max_window_bytes, sm_count, tc2::kPersistGroup)
                      : 0;
    if (group_tiles) *group_tiles = g;
    return g > 0 ? (eig_tiles + g - 1) / g : 1;
}

extern "C" int pg_probe_fusion(pg_handle* h, int32_t* fused, int32_t* g_columns, int32_t* pieces, float* build_ms)
{
    if (!h) return PG_ERR_ARG;
    if (fused) *fused = (h->have_design && fuse_wanted(h)) ? 1 : 0;
    if (g_columns) *g_columns = h->fz.cnodes * (h->c0 + h->q);
    if (pieces) *pieces = h->fz.npieces;
    if (build_ms) *build_ms = h->fz.build_ms;
    return PG_OK;
}

extern "C" int pg_set_design(pg_handle* h, const double* W_host, const double* y_host, int already_rotated, float* ms)
{
    if (!h) return PG_ERR_ARG;
    int rc = ensure_slots(h, 1);
    if (rc) return rc;
    activate(h, 0);
    h->have_design = false;
    rc = set_design_active(h, W_host, y_host, already_rotated, ms);
    if (rc) return rc;
    rc = build_v(h, 1);
    if (rc) return rc;
    h->q = 1;
    h->have_design = true;
    return PG_OK;
}

extern "C" int pg_set_design_multi(pg_handle* h, const double* W_host, const double* Y_host, int q, int already_rotated,
                                   float* ms)
{
    if (!h || !Y_host || q < 1) return fail(h, PG_ERR_ARG, "pg_set_design_multi: bad argument");
    if (q > PG_MAX_TRAITS)
        return fail(h, PG_ERR_ARG, "pg_set_design_multi: %d traits, at most %d per pass (scan in groups)", q, PG_MAX_TRAITS);
    int rc = ensure_slots(h, q);
    if (rc) return rc;
    h->have_design = false;
    std::vector<double> y(h->n);
    float total = 0.f;
    for (int ph = 0; ph < q; ++ph) {
        for (int l = 0; l < h->n; ++l) y[l] = Y_host[(size_t)l * q + ph];   // (n, q) C-order
        activate(h, ph);
        float t = 0.f;
        rc = set_design_active(h, W_host, y.data(), already_rotated, &t);
        if (rc) { activate(h, 0); return rc; }
        total += t;
    }
    activate(h, 0);
    rc = build_v(h, q);
    if (rc) return rc;
    h->q = q;
    h->have_design = true;
    if (ms) *ms = total;
    return PG_OK;
}

static int set_design_active(pg_handle* h, const double* W_host, const double* y_host, int already_rotated, float* ms)
{
    if (!h || !y_host || (!W_host && h->c0 > 0)) return fail(h, PG_ERR_ARG, "pg_set_design: NULL argument");
    if (!h->have_d) return fail(h, PG_ERR_ARG, "pg_set_design: call pg_set_kinship / pg_set_eigen first");
    if (!already_rotated && !h->have_U) return fail(h, PG_ERR_ARG, "pg_set_design: no U held and inputs are not rotated");
    CK(cudaSetDevice(h->device));
    const int n = h->n, c0 = h->c0, k0 = c0 + 1;
    // (n, c0) C-order + y -> column-major n x (c0+1)
    // rotated inputs arrive in the caller's eigen order: bring them into the handle's ascending order
    std::vector<double> col((size_t)n * k0);
    const bool gather = already_rotated && !h->perm_identity;
    for (int j = 0; j < c0; ++j)
        for (int l = 0; l < n; ++l) col[(size_t)j * n + l] = W_host[(size_t)(gather ? h->perm[l] : l) * c0 + j];
    for (int l = 0; l < n; ++l) col[(size_t)c0 * n + l] = y_host[gather ? h->perm[l] : l];
    EventGuard g0, g1;
    CK(cudaEventCreate(&g0.e));
    CK(cudaEventCreate(&g1.e));
    const cudaEvent_t e0 = g0.e, e1 = g1.e;
    CK(cudaEventRecord(e0, h->compute));
    if (already_rotated) {
        CK(cudaMemcpy2DAsync(h->wy, sizeof(double) * h->ldw, col.data(), sizeof(double) * n, sizeof(double) * n, k0,
                             cudaMemcpyHostToDevice, h->compute));
    } else {
        if (!h->design_raw) CK(cudaMalloc(&h->design_raw, sizeof(double) * col.size()));
        double* raw = h->design_raw;
        CK(cudaMemcpyAsync(raw, col.data(), sizeof(double) * col.size(), cudaMemcpyHostToDevice, h->compute));
        const double one = 1.0, zero = 0.0;
        // wy = U^T [W, y]  (lmm/lmm.py:245-246)
        cublasStatus_t s = cublasDgemm(h->blas, h->u_op_t ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, n, k0, n, &one, h->U,
                                       n, raw, n, &zero, h->wy, (int)h->ldw);
        cudaStreamSynchronize(h->compute);   // col is a local buffer: the copy must have left it
        if (s != CUBLAS_STATUS_SUCCESS) return fail(h, PG_ERR_CUBLAS, "pg_set_design: dgemm status %d", (int)s);
    }
    h->rotated_inputs = already_rotated != 0;
    h->slots[h->active_slot].null_valid = false;
    int rc = build_tables(h);
    if (rc) return rc;
    CK(cudaEventRecord(e1, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms = t;
    return PG_OK;
}

// ML fit of the null model [W0] of phenotype slot `ph` (null_model_kernel), cached until the design changes
static int ensure_null(pg_handle* h, int ph)
{
    if (ph < 0 || ph >= (int)h->slots.size()) return fail(h, PG_ERR_ARG, "null model: phenotype %d of %d", ph, (int)h->slots.size());
    pg_handle::DesignSlot& S = h->slots[ph];
    if (S.null_valid) return PG_OK;
    CK(cudaSetDevice(h->device));
    double* dout = nullptr;
    CK(cudaMalloc(&dout, sizeof(double) * 4));
    Tables2 t2 = h->tab2;
    t2.fix2 = S.fix2; t2.itab2 = S.itab2;
    null_model_kernel<<<1, 32, 0, h->compute>>>(h->n, t2, dout);
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaMemcpyAsync(S.null_vals, dout, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->compute);
    cudaError_t e3 = cudaStreamSynchronize(h->compute);
    cudaFree(dout);
    for (cudaError_t e : {e1, e2, e3})
        if (e != cudaSuccess) return fail(h, PG_ERR_CUDA, "null model: %s", cudaGetErrorString(e));
    S.null_valid = true;
    return PG_OK;
}

extern "C" int pg_null_model(pg_handle* h, int trait, double* lambda_null, double* tau_null, double* loglik_null)
{
    if (!h) return PG_ERR_ARG;
    if (!h->have_design) return fail(h, PG_ERR_ARG, "pg_null_model: call pg_set_design first");
    if (trait < 0 || trait >= h->q) return fail(h, PG_ERR_ARG, "pg_null_model: trait %d of %d", trait, h->q);
    int rc = ensure_null(h, trait);
    if (rc) return rc;
    const double* v = h->slots[trait].null_vals;
    if (lambda_null) *lambda_null = v[0];
    if (tau_null) *tau_null = v[1];
    if (loglik_null) *loglik_null = v[2];
    return PG_OK;
}

extern "C" int pg_set_stream(pg_handle* h, void* stream)
{
    if (!h) return PG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->compute));
    h->compute = stream ? (cudaStream_t)stream : h->own_compute;
    CKB(cublasSetStream(h->blas, h->compute));
    CKS(cusolverDnSetStream(h->solver, h->compute));
    return PG_OK;
}

extern "C" int pg_set_options(pg_handle* h, int rotation, int64_t block_snps)
{
    if (!h) return PG_ERR_ARG;
    if (rotation < PG_ROT_AUTO || rotation > PG_ROT_I8TC) return fail(h, PG_ERR_ARG, "pg_set_options: rotation %d", rotation);
    if (block_snps < 0) return fail(h, PG_ERR_ARG, "pg_set_options: block_snps %lld", (long long)block_snps);
    h->rotation = rotation;
    h->block_snps_opt = block_snps;
    return PG_OK;
}

extern "C" int pg_set_bed_options(pg_handle* h, int count_a1, int standardize)
{
    if (!h) return PG_ERR_ARG;
    h->rot.bed_count_a1 = count_a1 != 0;
    h->rot.bed_standardize = standardize != 0;
    return PG_OK;
}

extern "C" int pg_set_scan_mode(pg_handle* h, int mode)
{
    if (!h) return PG_ERR_ARG;
    if (mode != PG_SCAN_WALD && mode != PG_SCAN_DE) return fail(h, PG_ERR_ARG, "pg_set_scan_mode: mode %d", mode);
    h->scan_mode = mode;
    return PG_OK;
}

extern "C" int pg_set_reml_engine(pg_handle* h, int engine)
{
    if (!h) return PG_ERR_ARG;
    if (engine < PG_REML_AUTO || engine > PG_REML_WARP) return fail(h, PG_ERR_ARG, "pg_set_reml_engine: engine %d", engine);
    h->engine = engine;
    return PG_OK;
}

static size_t xdtype_size(int t) { return (t == PG_X_I8 || t == PG_X_BED) ? 1 : (t == PG_X_F32 ? 4 : 8); }

static int ensure_workspace(pg_handle* h, long long m, int xdtype, bool host_input)
{
    const int n = h->n;
    long long blk = h->block_snps_opt;
    if (blk <= 0) {
        blk = (long long)((size_t(1) << 31) / (sizeof(double) * (size_t)n));  // ~2 GiB of fp64 per buffer
        blk = std::max<long long>(256, (blk / 256) * 256);
        blk = std::min<long long>(blk, 32768);
        // keep at least four blocks in flight on large inputs so uploads overlap compute
        if (m >= 4 * 8192) blk = std::min<long long>(blk, std::max<long long>(8192, ((m + 3) / 4 + 255) / 256 * 256));
        // a host-resident shard that is worth pipelining (>= kBounceMinBytes) but shorter than that, e.g. one rank's 12 544 SNPs of a
        // problem split over 8 GPUs: still about four blocks of whole 512-SNP tiles, so that packing, upload and compute overlap
        else if (host_input && (size_t)m * n * xdtype_size(xdtype) >= kBounceMinBytes && m >= 4 * 2048)
            blk = std::min<long long>(blk, std::max<long long>(2048, ((m + 3) / 4 + 511) / 512 * 512));
        // compressed moments of a block: at most 2 GiB
        const size_t zrow = sizeof(double) * (size_t)(h->k1p - 1 + h->q) * std::max(h->plan.Kcp, 32);
        blk = std::min<long long>(blk, std::max<long long>(256, (long long)((size_t(1) << 31) / zrow) / 256 * 256));
    }
    blk = std::min(blk, std::max<long long>(m, 1));
    blk = ((blk + 31) / 32) * 32;
    const size_t need = (size_t)blk * h->ldx;
    if (need > h->xbuf_elems) {
        if (h->xf) cudaFree(h->xf);
        h->xf = nullptr;
        for (int s = 0; s < 2; ++s) {
            if (h->xr[s]) cudaFree(h->xr[s]);
            h->xr[s] = nullptr;
        }
        h->xbuf_elems = 0;
        h->last_xr = nullptr;
        CK(cudaMalloc(&h->xf, sizeof(double) * need));
        for (int s = 0; s < 2; ++s) {
            CK(cudaMalloc(&h->xr[s], sizeof(double) * need));
            CK(cudaMemset(h->xr[s], 0, sizeof(double) * need));
        }
        h->xbuf_elems = need;
    }
    if (h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED) {
        const int zrows = h->k1p - 1 + h->q;
        const size_t zneed = (size_t)blk * zrows * h->plan.Kcp;
        if (zneed <= h->z_elems && zrows != h->z_rows) {
            // another slab layout than the one the buffers last held: padding rows / nodes must read as zero again
            for (int s = 0; s < 2; ++s) CK(cudaMemsetAsync(h->Z[s], 0, sizeof(double) * h->z_elems, h->compute));
        }
        h->z_rows = zrows;
        if (zneed > h->z_elems) {
            for (int s = 0; s < 2; ++s) {
                if (h->Z[s]) cudaFree(h->Z[s]);
                h->Z[s] = nullptr;
            }
            h->z_elems = 0;
            for (int s = 0; s < 2; ++s) {
                CK(cudaMalloc(&h->Z[s], sizeof(double) * zneed));
                CK(cudaMemset(h->Z[s], 0, sizeof(double) * zneed));  // padding nodes / rows stay zero for ever
            }
            h->z_elems = zneed;
        }
        h->ldF = (blk + 31) / 32 * 32;
        const size_t fneed = (size_t)2 * kNumFixed * zrows * h->ldF;   // double2 per (lambda, slab row, SNP)
        if (fneed > h->f_elems) {
            for (int s = 0; s < 2; ++s) {
                if (h->F[s]) cudaFree(h->F[s]);
                h->F[s] = nullptr;
            }
            h->f_elems = 0;
            for (int s = 0; s < 2; ++s) CK(cudaMalloc(&h->F[s], sizeof(double) * fneed));
            h->f_elems = fneed;
        }
        const size_t fxneed = (size_t)kNumFixed * kFxOut * h->ldF;
        if (fxneed > h->fx_elems) {
            for (int s = 0; s < 2; ++s) {
                if (h->FX[s]) cudaFree(h->FX[s]);
                h->FX[s] = nullptr;
            }
            h->fx_elems = 0;
            for (int s = 0; s < 2; ++s) CK(cudaMalloc(&h->FX[s], sizeof(double) * fxneed));
            h->fx_elems = fxneed;
        }
    }
    const size_t sbytes = need * xdtype_size(xdtype);
    if (sbytes > h->stage_bytes) {
        for (int s = 0; s < 2; ++s) {
            if (h->stage[s]) cudaFree(h->stage[s]);
            h->stage[s] = nullptr;
        }
        h->stage_bytes = 0;
        for (int s = 0; s < 2; ++s) CK(cudaMalloc(&h->stage[s], sbytes));
        h->stage_bytes = sbytes;
    }
    h->blk = blk;
    return PG_OK;
}

static int launch_stage(pg_handle* h, const void* src, int xdtype, long long ld, int layout, long long mb, double* dst)
{
    if (stage_to_snp_major(h->compute, h->n, src, xdtype, ld, layout, mb, dst, h->ldx,
                           h->perm_identity ? nullptr : h->perm_dev)) {
        cudaError_t e_ = cudaGetLastError();
        return fail(h, PG_ERR_CUDA, "staging kernel: %s", cudaGetErrorString(e_));
    }
    return PG_OK;
}

// REML stage of one block on stream `st`: reads xr, uses moment buffer Zbuf.  `ev_xr_done` (nullable) is recorded once
// the rotated genotypes are no longer needed (after the compression, or after the kernel for the streaming engines).
// `st` carries the compression, `st_solve` the optimiser + p-value kernels (the same stream, or the auxiliary stream
// when the optimiser of this block is to run under the next block's rotation: then `ev_z[0]` is recorded after the
// compression, `ev_z[1]` after the optimiser, and the optimiser is limited to one CTA per SM so that a rotation CTA
// still fits beside it).
static int launch_reml(pg_handle* h, cudaStream_t st, const double* xr, double* Zbuf, long long mb, long long row0,
                       int grid_mode, double* const out[6], int* status, int* e2, int* e3, cudaEvent_t* ev_mid = nullptr,
                       cudaEvent_t ev_xr_done = nullptr, cudaStream_t st_solve = nullptr, cudaEvent_t* ev_z = nullptr,
                       int parity = 0, int ph = 0, double* const* lrt = nullptr, bool have_moments = false)
{
    // have_moments: the fused rotation already wrote this block's moments into Zbuf (no compression, xr is not read)
    // ph: phenotype slot of this launch (the caller has activated it).  The compressed engine compresses the block for
    // all h->q phenotypes at ph == 0; the direct engines read xr for every phenotype.
    const int q = h->q;
    if (!st_solve) st_solve = st;
    const bool split = st_solve != st;
    unsigned long long* counter = h->counter + (parity & 1);
    double* Fbuf = Zbuf == h->Z[1] ? h->F[1] : h->F[0];
    double* FXbuf = Zbuf == h->Z[1] ? h->FX[1] : h->FX[0];
    ScanArgs a;
    a.n = h->n; a.c0 = h->c0; a.grid = grid_mode; a.m = mb; a.row0 = row0;
    a.d = h->d; a.wy = h->wy; a.ldw = h->ldw; a.xr = xr; a.ldx = h->ldx; a.tab = h->tab;
    for (int i = 0; i < 6; ++i) a.out[i] = out[i];
    a.status = status; a.n_eval2 = e2; a.n_eval3 = e3; a.counter = counter;
    if (h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED) {
        const DevPlan& P = h->plan;
        const int ntiles = (int)((mb + kCtSnps - 1) / kCtSnps);
        const int zrows = h->k1p - 1 + q;
        if (ph == 0) {
            if (ev_mid) CK(cudaEventRecord(ev_mid[0], st));
            if (P.nitems && !have_moments) {
                CK(cudaFuncSetAttribute(compress_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtSmemBytes));
                compress_dmma_kernel<<<(unsigned)((long long)P.nitems * ntiles), 256, kCtSmemBytes, st>>>(
                    xr, h->ldx, mb, P.items, P.V, P.vpitch, h->c0, h->k1p, zrows, P.Kcp, Zbuf, ntiles);
                CK(cudaGetLastError());
            }
            if (P.ncopy && !have_moments) {
                dim3 grid((unsigned)((P.ncopy + 127) / 128), (unsigned)((mb + 7) / 8));
                compress_copy_kernel<<<grid, 128, 0, st>>>(xr, h->ldx, mb, P.copy_l, P.copy_node, P.ncopy,
                                                           q > 1 ? h->wy_all : h->wy, h->ldw, h->c0, P.klin, h->k1p, zrows,
                                                           P.Kcp, Zbuf);
                CK(cudaGetLastError());
            }
            if (ev_mid) CK(cudaEventRecord(ev_mid[1], st));
            if (ev_xr_done) CK(cudaEventRecord(ev_xr_done, st));
            {
                // x rows of the kNumFixed fixed-lambda evaluations of every slab row, all phenotypes at once
                const long long rows = mb * zrows;
                // H goes through shared memory in slabs of at most kFxSlabMax nodes (Kcp is a multiple of 32)
                const int kslab = std::min(P.Kcp, kFxSlabMax);
                const size_t hs = sizeof(double) * (size_t)kslab * kFxCols;
                if (hs > kDynSmemOptIn)
                    CK(cudaFuncSetAttribute(fixed_xrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
                fixed_xrow_kernel<<<(unsigned)((rows + 255) / 256), 256, hs, st>>>(Zbuf, rows, P.Kcp, P.H,
                                                                                    reinterpret_cast<double2*>(Fbuf), kslab, zrows,
                                                                                    h->ldF);
                CK(cudaGetLastError());
            }
            if (split) {
                CK(cudaEventRecord(ev_z[0], st));
                CK(cudaStreamWaitEvent(st_solve, ev_z[0], 0));
            }
        }
        CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st_solve));
        {
            // the kNumFixed fixed-lambda evaluations of this phenotype, one thread per SNP (reml_solve.cuh)
            FixedArgs fa;
            fa.c0 = h->c0; fa.zrows = zrows; fa.yrow = h->k1p - 1 + ph; fa.swap = (h->scan_mode == PG_SCAN_DE) ? 1 : 0;
            fa.m = mb; fa.ldF = h->ldF; fa.F2 = reinterpret_cast<const double2*>(Fbuf); fa.t2 = h->tab2; fa.FX = FXbuf;
            const size_t fsm = sizeof(double) * ((size_t)((h->tab2.NF2 + 1) & ~1) + (size_t)2 * (h->c0 + 2) * 128);
            if (fsm > kDynSmemOptIn)
                CK(cudaFuncSetAttribute(fixed_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
            fixed_phase_kernel<<<(unsigned)((mb + 127) / 128), 128, fsm, st_solve>>>(fa);
            CK(cudaGetLastError());
        }
        SolveArgs sa;
        sa.n = h->n; sa.c0 = h->c0; sa.grid = grid_mode; sa.m = mb; sa.row0 = row0;
        sa.nodes = P.nodes; sa.Kcp = P.Kcp; sa.Kc = P.Kc; sa.Z = Zbuf; sa.k1p = h->k1p; sa.t2 = h->tab2;
        sa.zrows = zrows; sa.yrow = h->k1p - 1 + ph; sa.FX = FXbuf; sa.ldF = h->ldF; sa.swap = (h->scan_mode == PG_SCAN_DE) ? 1 : 0;
        for (int i = 0; i < 4; ++i) sa.lrt[i] = lrt ? lrt[i] : nullptr;
        sa.l_null = lrt ? h->slots[ph].null_vals[2] : 0.0;
        for (int i = 0; i < 6; ++i) sa.out[i] = out[i];
        sa.status = status; sa.n_eval2 = e2; sa.n_eval3 = e3; sa.counter = counter;
        // per-tile shared memory (a tile = the 32 or 16 lanes that own one SNP): x-row scratch + an interpolated table-2 row,
        // and -- when it pays -- the SNP's whole moment slab (k1p rows x Kcp nodes), staged once per SNP
        const size_t scratch_d = (size_t)((3 * h->k1p + h->tab2.NF2 + 1) & ~1);
        const size_t slab_d = (size_t)h->k1p * P.Kcp;
        static const bool zsm_env = !(getenv("PG_SOLVE_ZSM") && atoi(getenv("PG_SOLVE_ZSM")) == 0);
        // two SNPs per warp when the x row has 9..16 entries (7 <= c0 <= 14): the covariate recursion and the optimiser's
        // scalar code never use more lanes.  Measured (profiles/bench_r02_solver_tile_ab.txt): c0 = 10 solve stage 6.5 -> 5.7 ms
        // per 100 k SNPs, c0 = 11 unchanged, c0 = 6 (8 entries, slab staged in shared memory) 4.9 -> 5.3 ms, so short rows keep
        // a warp per SNP.  PG_SOLVE_TILE=32 forces a warp per SNP everywhere.
        static const int tile_env = getenv("PG_SOLVE_TILE") ? atoi(getenv("PG_SOLVE_TILE")) : 16;
        const int T = (tile_env == 16 && h->c0 + 2 <= 16 && h->k1p > 8 && !split) ? 16 : 32;
        const int tpw = 32 / T;   // tiles per warp
        const size_t budget = 200 * 1024;   // per SM
        int warps = 8, ctas_fit = 2;
        bool zsm = false;
        // Measured (profiles/bench_r02_solver_slab_ab.txt): with short x rows (k1p <= 8, e.g. the GD449 shape, c0 = 6) the slab is
        // 8 KB and 16 warps per SM keep it resident: solve stage 5.9 -> 5.0 ms per 100 k SNPs; for k1p = 12 / 16 only 10-11
        // warps fit and the lost occupancy costs more than the L2 latency saved (c0 = 10: 6.6 -> 8.9 ms), so those stay on
        // the read-only cache path.  Grid mode without LRT has no SNP-specific evaluation at all.
        if (zsm_env && !split && h->k1p <= 8 && (!grid_mode || lrt)) {
            const size_t pw = sizeof(double) * (scratch_d + slab_d) * tpw;
            const int fit = (int)(budget / pw);          // warps per SM with the slabs resident
            if (fit >= 16) { warps = 8; ctas_fit = 2; zsm = true; }
            else if (fit >= 12) { warps = 6; ctas_fit = 2; zsm = true; }
            else if (fit >= 10) { warps = 5; ctas_fit = 2; zsm = true; }
            else if (fit >= 8) { warps = 8; ctas_fit = 1; zsm = true; }
            else if (fit >= 6) { warps = 6; ctas_fit = 1; zsm = true; }
        }
        const size_t per_warp = sizeof(double) * (scratch_d + (zsm ? slab_d : 0)) * tpw;
        if (!zsm) while (warps > 1 && per_warp * warps > 100 * 1024) warps >>= 1;
        // under the next block's rotation (PG_OVERLAP): the rotation CTA leaves 10 K registers and ~37 KB of shared memory
        // per SM, i.e. room for two solver warps (128 registers each) beside it
        static const int ov_warps = getenv("PG_OVERLAP_WARPS") ? std::max(1, atoi(getenv("PG_OVERLAP_WARPS"))) : 2;
        if (split) warps = std::min(warps, ov_warps);
        const size_t smem = per_warp * warps;
        sa.zsm = zsm ? 1 : 0;
        const bool two = (h->c0 + 2) > 32;
        auto launch = [&](auto kern, int ctas) -> int {
            if (smem > kDynSmemOptIn) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(ctas, budget / std::max<size_t>(smem, 1)));
            const long long per_cta = (long long)warps * tpw;   // SNPs in flight per CTA
            long long want = (mb + per_cta - 1) / per_cta;
            int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->sm_count * ctas_per_sm));
            kern<<<grid, warps * 32, smem, st_solve>>>(sa);
            return PG_OK;
        };
        // 2 CTAs of 8 warps per SM (128 registers) for x rows of 12 or more entries: 3 and 4 CTAs/SM spill in the x-row
        // pass and were measured 5-25 % slower (c0 = 10, 11); short x rows (c0 <= 6, 8 entries) run 27 % faster at 4
        // CTAs/SM (64 registers).  4-6 warps per CTA were measured slower.
        // Under the next block's rotation: 1 CTA per SM.
        const bool small = h->k1p <= 8;
        const int per_sm = split ? 1 : (zsm ? ctas_fit : (small ? 4 : 2));
        // with the slab resident the shared memory caps the SM at two CTAs anyway: the 128-register variant (no spills)
        int lr;
        if (two) lr = launch(reml_solve_kernel<2, 2, 32>, per_sm);
        else if (T == 16) lr = (small && !zsm) ? launch(reml_solve_kernel<1, 4, 16>, per_sm) : launch(reml_solve_kernel<1, 2, 16>, per_sm);
        else lr = (small && !zsm) ? launch(reml_solve_kernel<1, 4, 32>, per_sm) : launch(reml_solve_kernel<1, 2, 32>, per_sm);
        if (lr) return lr;
        CK(cudaGetLastError());
        // p-values, one thread per SNP (NaN F -> NaN p, so failed rows stay NaN).  The kernel is one long serial chain
        // per thread (0.18 ms whatever the row count), so a scan launches it once over all rows after its last block.
        if (!h->defer_pvalues) {
            const long double nu = (long double)(h->n - h->c0 - 1);
            const double lnbeta = (double)(lgammal(0.5L * nu + 0.5L) - lgammal(0.5L * nu) - lgammal(0.5L));
            pvalue_kernel<<<(unsigned)((mb + 31) / 32), 32, 0, st_solve>>>(out[4], out[5], row0, mb, (double)nu, lnbeta);
        }
        CK(cudaGetLastError());
        if (split && ph == q - 1) CK(cudaEventRecord(ev_z[1], st_solve));
        return PG_OK;
    }
    if (ph != q - 1) ev_xr_done = nullptr;   // the direct engines read xr for every phenotype
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    if (h->engine == PG_REML_STREAM) {
        const StreamCfg cfg = stream_config(h->c0);
        const int ncmax = std::min(h->c0 + 1, (int)kChunkCols);
        CK(cudaFuncSetAttribute(reml_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        const int ctas_per_sm = cfg.smem <= 110 * 1024 ? 2 : 1;
        long long want = (mb + cfg.nw - 1) / cfg.nw;
        int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->sm_count * ctas_per_sm));
        reml_stream_kernel<<<grid, cfg.nw * 32, cfg.smem, st>>>(a, cfg.nw, ncmax);
        CK(cudaGetLastError());
        if (ev_xr_done) CK(cudaEventRecord(ev_xr_done, st));
        return PG_OK;
    }
    const int k = h->c0 + 2, TT = k * (k + 1) / 2;
    const size_t per_warp = sizeof(double) * 3 * TT;
    int warps = 8;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    const size_t smem = per_warp * warps;
    if (smem > kDynSmemOptIn)
        CK(cudaFuncSetAttribute(reml_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1)));
    long long want = (mb + warps - 1) / warps;
    int grid = (int)std::min<long long>(want, (long long)h->sm_count * ctas_per_sm);
    grid = std::max(grid, 1);
    reml_scan_kernel<<<grid, warps * 32, smem, st>>>(a);
    CK(cudaGetLastError());
    if (ev_xr_done) CK(cudaEventRecord(ev_xr_done, st));
    return PG_OK;
}

// host_pack.cpp: packs a strided host block into a pinned bounce buffer (host threads, streaming stores)
namespace pg { void pack_block_threads(char* dst, const char* src, size_t pitch, size_t width, size_t rows); }

static bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

struct EvPair {
    cudaEvent_t a, b;
};

static int scan_impl(pg_handle* h, const void* X, int xdtype, long long ld, int layout, long long m, int grid_mode,
                     double* const out_user[6], int32_t* status, int32_t* e2, int32_t* e3, pg_timing* timing,
                     bool on_device, double* const lrt_user[4] = nullptr)
{
    if (!h) return PG_ERR_ARG;
    if (!X || m < 0) return fail(h, PG_ERR_ARG, "pg_scan: NULL X or negative m");
    for (int i = 0; i < 6; ++i)
        if (!out_user[i] && m > 0) return fail(h, PG_ERR_ARG, "pg_scan: NULL output %d", i);
    if (xdtype < PG_X_I8 || xdtype > PG_X_BED) return fail(h, PG_ERR_ARG, "pg_scan: xdtype %d", xdtype);
    if (layout != PG_X_SAMPLE_MAJOR && layout != PG_X_SNP_MAJOR) return fail(h, PG_ERR_ARG, "pg_scan: layout %d", layout);
    if (!h->have_design) return fail(h, PG_ERR_ARG, "pg_scan: call pg_set_design first");
    if (h->scan_mode == PG_SCAN_DE && !(h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED))
        return fail(h, PG_ERR_ARG, "pg_scan: PG_SCAN_DE runs on the compressed REML engine only");
    const bool want_lrt = lrt_user != nullptr;
    if (want_lrt) {
        for (int i = 0; i < 4; ++i)
            if (!lrt_user[i] && m > 0) return fail(h, PG_ERR_ARG, "pg_scan_lrt: NULL likelihood-ratio output %d", i);
        if (!(h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED))
            return fail(h, PG_ERR_ARG, "pg_scan_lrt: the likelihood-ratio outputs need the compressed REML engine");
        if (h->scan_mode == PG_SCAN_DE) return fail(h, PG_ERR_ARG, "pg_scan_lrt: not available in PG_SCAN_DE mode");
    }
    const int n = h->n;
    const bool bed = xdtype == PG_X_BED;
    if (bed && layout != PG_X_SNP_MAJOR) return fail(h, PG_ERR_ARG, "pg_scan: PLINK .bed data is SNP-major");
    if (bed && h->rotated_inputs) return fail(h, PG_ERR_ARG, "pg_scan: PLINK .bed genotypes cannot be pre-rotated");
    if ((layout == PG_X_SAMPLE_MAJOR && ld < m) || (layout == PG_X_SNP_MAJOR && ld < (bed ? (n + 3) / 4 : n)))
        return fail(h, PG_ERR_ARG, "pg_scan: ld %lld too small", ld);
    if (timing) memset(timing, 0, sizeof *timing);
    if (m == 0) return PG_OK;
    CK(cudaSetDevice(h->device));
    int rc = ensure_workspace(h, m, xdtype, !on_device);
    if (rc) return rc;
    const long long blk = h->blk;
    // worth it for large pageable inputs only; very wide element types would need multi-GB pinned buffers
    const size_t total_bytes = (size_t)m * n * xdtype_size(xdtype), block_bytes = (size_t)blk * n * xdtype_size(xdtype);
    const bool use_bounce = !on_device && total_bytes >= kBounceMinBytes && block_bytes <= (size_t(1) << 30) &&
                            !getenv("PG_NO_BOUNCE") && !host_pointer_is_pinned(X);
    if (use_bounce) {
        const size_t need_b = (size_t)blk * n * xdtype_size(xdtype);
        if (need_b > h->bounce_bytes) {
            for (int s = 0; s < 2; ++s) {
                if (h->bounce[s]) cudaFreeHost(h->bounce[s]);
                h->bounce[s] = nullptr;
            }
            h->bounce_bytes = 0;
            for (int s = 0; s < 2; ++s) CK(cudaMallocHost(&h->bounce[s], need_b));
            h->bounce_bytes = need_b;
        }
        for (int s = 0; s < 2; ++s)
            if (!h->ev_bounce_done[s]) CK(cudaEventCreateWithFlags(&h->ev_bounce_done[s], cudaEventDisableTiming));
    }
    // Block boundaries.  Host input with automatic blocking: the blocks grow geometrically (x 1.6 from 1 536 SNPs) up to the
    // full block, so that compute starts after ~0.6 ms of packing + upload instead of a whole block's worth and every later
    // block is packed and uploaded in less time than the block before it computes (n = 10 000: 0.35 us per SNP to pack and
    // upload, 0.6 us to compute; the ratio is >= 1.6 at every n measured).  Round 1 used two short blocks (1/16, 1/4): the
    // third, full-sized block then waited ~4 ms for its 250 MB while the GPU had 5 ms of work.  PG_RAMP=0: no short blocks.
    const std::vector<long long> bstart = plan_blocks(m, blk, !on_device, h->block_snps_opt > 0);
    const long long nblocks = (long long)bstart.size() - 1;
    const size_t esz = xdtype_size(xdtype);
    const size_t snp_row_bytes = bed ? (size_t)((n + 3) / 4) : (size_t)n * esz;   // one SNP of an SNP-major block
    const bool rotate = !h->rotated_inputs;

    // outputs: device arrays (user's when on_device, temporaries otherwise)
    double* dout[6];
    int *dstatus = nullptr, *de2 = nullptr, *de3 = nullptr;
    double* dtmp = nullptr;
    int* itmp = nullptr;
    const int q = h->q;                       // phenotypes of the current design (pg_set_design_multi)
    const size_t mq = (size_t)m * (size_t)q;  // every output array holds q * m values, phenotype-major
    double* dlrt[4] = {nullptr, nullptr, nullptr, nullptr};
    constexpr int kResCols = 10;   // six Wald columns + four likelihood-ratio columns
    if (on_device) {
        for (int i = 0; i < 6; ++i) dout[i] = out_user[i];
        if (want_lrt) for (int i = 0; i < 4; ++i) dlrt[i] = lrt_user[i];
        dstatus = status; de2 = e2; de3 = e3;
    } else {
        if (mq > h->res_cap) {
            if (h->res_d) cudaFree(h->res_d);
            if (h->res_i) cudaFree(h->res_i);
            if (h->res_host) cudaFreeHost(h->res_host);
            h->res_d = nullptr; h->res_i = nullptr; h->res_host = nullptr; h->res_cap = 0;
            const size_t cap = (mq + 1023) / 1024 * 1024;
            CK(cudaMalloc(&h->res_d, sizeof(double) * kResCols * cap));
            CK(cudaMalloc(&h->res_i, sizeof(int) * 3 * cap));
            CK(cudaMallocHost(&h->res_host, (sizeof(double) * kResCols + sizeof(int) * 3) * cap));
            h->res_cap = cap;
        }
        dtmp = h->res_d; itmp = h->res_i;
        for (int i = 0; i < 6; ++i) dout[i] = dtmp + (size_t)i * mq;
        if (want_lrt) for (int i = 0; i < 4; ++i) dlrt[i] = dtmp + (size_t)(6 + i) * mq;
        dstatus = itmp; de2 = itmp + mq; de3 = itmp + 2 * mq;
    }

    // timing events come from a pool owned by the handle (creating / destroying ~40 events per call costs host time)
    size_t ev_next = 0;
    bool ev_ok = true;
    auto take = [&]() -> cudaEvent_t {
        if (ev_next == h->ev_pool.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) { ev_ok = false; return nullptr; }
            h->ev_pool.push_back(e);
        }
        return h->ev_pool[ev_next++];
    };
    std::vector<EvPair> ev_conv(nblocks), ev_rot(nblocks), ev_reml(nblocks), ev_h2d(nblocks), ev_cmp(nblocks);
    auto mk = [&](EvPair& p) { p.a = take(); p.b = take(); };
    for (long long b = 0; b < nblocks; ++b) { mk(ev_conv[b]); mk(ev_rot[b]); mk(ev_reml[b]); mk(ev_cmp[b]); if (!on_device) mk(ev_h2d[b]); }
    const bool compressed = (h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED);
    cudaEvent_t t0 = take(), t1 = take(), t2 = take();
    if (!ev_ok) return fail(h, PG_ERR_CUDA, "pg_scan: cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    int n_rot_launch = 0, last_engine = 0;

    if (want_lrt)
        for (int ph = 0; ph < q; ++ph) {
            rc = ensure_null(h, ph);
            if (rc) return rc;
        }
    // moments straight from the fused rotation where it pays (fuse_wanted); rot_run still decides per block
    const bool want_fuse = rotate && fuse_wanted(h);
    if (want_fuse) {
        rc = ensure_fused_operand(h);
        if (rc) return rc;
        pg_handle::Fused& F = h->fz;
        const long long ldp = (blk + 511) / 512 * 512;
        const size_t need_p = (size_t)std::max(F.npieces, 1) * kCq * ldp;
        if (need_p > F.p2_elems) {
            if (F.P2) cudaFree(F.P2);
            F.P2 = nullptr; F.p2_elems = 0;
            CK(cudaMalloc(&F.P2, sizeof(double) * need_p));
            F.p2_elems = need_p;
        }
        F.ldp = ldp;
    }
    int n_fused_blocks = 0;
    const int ncols = want_lrt ? kResCols : 6;
    h->defer_pvalues = compressed;   // launch_reml leaves p = NaN; one pvalue_kernel launch follows the last block
    rc = [&]() -> int {
        CK(cudaEventRecord(t0, h->compute));
        CK(cudaStreamWaitEvent(h->copy, t0, 0));
        if (h->overlap) {
            CK(cudaStreamWaitEvent(h->aux, t0, 0));
            CK(cudaStreamWaitEvent(h->cmb, t0, 0));
        }
        for (long long b = 0; b < nblocks; ++b) {
            const long long g0 = bstart[b], mb = bstart[b + 1] - g0;
            const int s = (int)(b & 1);
            const void* src_dev;
            long long ld_dev;
            if (on_device) {
                src_dev = (layout == PG_X_SAMPLE_MAJOR) ? (const char*)X + (size_t)g0 * esz
                                                        : (const char*)X + (size_t)g0 * ld * esz;
                ld_dev = ld;
            } else {
                if (b >= 2) CK(cudaStreamWaitEvent(h->copy, h->ev_free[s], 0));
                if (use_bounce) {
                    // host threads pack the block while the GPU is busy with the previous one
                    if (b >= 2) CK(cudaEventSynchronize(h->ev_bounce_done[s]));  // the upload that last read bounce[s]
                    const size_t width = layout == PG_X_SAMPLE_MAJOR ? (size_t)mb * esz : snp_row_bytes;
                    const size_t rows = layout == PG_X_SAMPLE_MAJOR ? (size_t)n : (size_t)mb;
                    const char* src0 = layout == PG_X_SAMPLE_MAJOR ? (const char*)X + (size_t)g0 * esz
                                                                    : (const char*)X + (size_t)g0 * ld * esz;
                    pack_block_threads(h->bounce[s], src0, (size_t)ld * esz, width, rows);
                    CK(cudaEventRecord(ev_h2d[b].a, h->copy));
                    CK(cudaMemcpyAsync(h->stage[s], h->bounce[s], width * rows, cudaMemcpyHostToDevice, h->copy));
                    CK(cudaEventRecord(h->ev_bounce_done[s], h->copy));
                    ld_dev = layout == PG_X_SAMPLE_MAJOR ? mb : (long long)(snp_row_bytes / esz);
                } else {
                CK(cudaEventRecord(ev_h2d[b].a, h->copy));
                if (layout == PG_X_SAMPLE_MAJOR) {
                    // n rows of mb elements, host pitch ld
                    CK(cudaMemcpy2DAsync(h->stage[s], (size_t)mb * esz, (const char*)X + (size_t)g0 * esz,
                                         (size_t)ld * esz, (size_t)mb * esz, (size_t)n, cudaMemcpyHostToDevice, h->copy));
                    ld_dev = mb;
                } else {
                    CK(cudaMemcpy2DAsync(h->stage[s], snp_row_bytes, (const char*)X + (size_t)g0 * ld * esz,
                                         (size_t)ld * esz, snp_row_bytes, (size_t)mb, cudaMemcpyHostToDevice, h->copy));
                    ld_dev = (long long)(snp_row_bytes / esz);
                }
                }
                CK(cudaEventRecord(ev_h2d[b].b, h->copy));
                CK(cudaEventRecord(h->ev_ready[s], h->copy));
                CK(cudaStreamWaitEvent(h->compute, h->ev_ready[s], 0));
                src_dev = h->stage[s];
            }
            // ---- rotation of block b into xr[s] (main stream + recombination stream)
            // compression on the main stream; with `overlap` the optimiser goes to the auxiliary stream and runs under
            // the rotation of block b+1 (it needs the fused rotation kernel as neighbour: the cuBLAS GEMM fills the SM)
            const bool compressed_engine = (h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED);
            cudaStream_t st_reml = h->compute;
            cudaStream_t st_solve = (h->overlap && compressed_engine) ? h->aux : h->compute;
            cudaStream_t st_cmb = h->compute;
            if (st_solve != st_reml && b >= 2) CK(cudaStreamWaitEvent(st_reml, h->ev_z_free[s], 0));  // Z[s] is free again
            double* xr_block = h->xr[s];
            if (b >= 2) {  // the REML stage of block b-2 has finished reading xr[s]
                CK(cudaStreamWaitEvent(h->compute, h->ev_xr_free[s], 0));
                CK(cudaStreamWaitEvent(st_cmb, h->ev_xr_free[s], 0));
            }
            CK(cudaEventRecord(ev_conv[b].a, h->compute));
            int used_i8 = 0, moments_done = 0;
            if (rotate) {
                FuseLaunch fl;
                if (want_fuse) {
                    const pg_handle::Fused& F = h->fz;
                    tc2::FuseArgs& fa = fl.args;
                    fa.g_tiles = F.g_tiles; fa.gscale = F.gscale; fa.g1 = F.g1; fa.goff = F.goff; fa.einfo = F.einfo;
                    fa.Lw = h->plan.Lw; fa.wy = q > 1 ? h->wy_all : h->slots[0].wy; fa.ldw = h->ldw; fa.klin = F.klin;
                    fa.jrow = F.jrow; fa.Z = h->Z[s]; fa.ldz = (long long)(h->k1p - 1 + q) * h->plan.Kcp;
                    fa.x2row = h->c0 * h->plan.Kcp; fa.P2 = F.P2; fa.ldp = F.ldp;
                    fl.planes_g = F.planes_g; fl.segs = F.segs; fl.nsegs = F.nsegs;
                }
                int r2 = rot_run(&h->rot, h->blas, h->compute, st_cmb, h->rotation, h->U, h->u_op_t, n, src_dev, xdtype,
                                 ld_dev, layout, mb, blk, h->xf, xr_block, h->ldx, &used_i8, &n_rot_launch, ev_conv[b].b,
                                 ev_rot[b].a, ev_rot[b].b, want_fuse ? &fl : nullptr, &moments_done);
                if (r2 != 0) return fail(h, r2, "rotation failed: %s", rot_error(&h->rot));
                n_fused_blocks += moments_done;
                last_engine = used_i8;  // rot_run reports PG_ROT_I8SPLIT / PG_ROT_I8TC, 0 for the FP64 GEMM
                if (!last_engine) last_engine = PG_ROT_FP64;
                CK(cudaEventRecord(h->ev_xr_ready[s], (used_i8 == PG_ROT_I8SPLIT) ? st_cmb : h->compute));
            } else {
                int r2 = launch_stage(h, src_dev, xdtype, ld_dev, layout, mb, xr_block);
                if (r2) return r2;
                CK(cudaEventRecord(ev_conv[b].b, h->compute));
                CK(cudaEventRecord(ev_rot[b].a, h->compute));
                CK(cudaEventRecord(ev_rot[b].b, h->compute));
                CK(cudaEventRecord(h->ev_xr_ready[s], h->compute));
            }
            if (!on_device) CK(cudaEventRecord(h->ev_free[s], h->compute));
            // ---- REML stage of block b (aux stream): runs under the rotation of block b+1
            CK(cudaStreamWaitEvent(st_reml, h->ev_xr_ready[s], 0));
            CK(cudaEventRecord(ev_reml[b].a, st_reml));
            cudaEvent_t mid[2] = {ev_cmp[b].a, ev_cmp[b].b};
            cudaEvent_t evz[2] = {h->ev_z_ready[s], h->ev_z_free[s]};
            for (int ph = 0; ph < q; ++ph) {
                // one solve per phenotype on the block's rotated genotypes; rotation and compression are shared
                activate(h, ph);
                double* outp[6];
                for (int i = 0; i < 6; ++i) outp[i] = dout[i] + (size_t)ph * m;
                double* lrtp[4];
                for (int i = 0; i < 4; ++i) lrtp[i] = want_lrt ? dlrt[i] + (size_t)ph * m : nullptr;
                int r3 = launch_reml(h, st_reml, xr_block, h->Z[s], mb, g0, grid_mode, outp,
                                     dstatus ? dstatus + (size_t)ph * m : nullptr, de2 ? de2 + (size_t)ph * m : nullptr,
                                     de3 ? de3 + (size_t)ph * m : nullptr, mid, h->ev_xr_free[s], st_solve, evz, s, ph,
                                     want_lrt ? lrtp : nullptr, moments_done != 0);
                if (r3) { activate(h, 0); return r3; }
            }
            activate(h, 0);
            CK(cudaEventRecord(ev_reml[b].b, st_solve));
            h->last_block_count = mb;
            h->last_block_row0 = g0;
            h->last_xr = moments_done ? nullptr : xr_block;   // a fused block leaves no rotated genotypes behind
        }
        if (h->overlap) {
            CK(cudaEventRecord(h->ev_aux_done, h->aux));
            CK(cudaStreamWaitEvent(h->compute, h->ev_aux_done, 0));
        }
        if (h->defer_pvalues && mq > 0) {
            const long double nu = (long double)(h->n - h->c0 - 1);
            const double lnbeta = (double)(lgammal(0.5L * nu + 0.5L) - lgammal(0.5L * nu) - lgammal(0.5L));
            for (int s = 0; s < 2; ++s)
                if (!h->ev_pv[s]) CK(cudaEventCreate(&h->ev_pv[s]));
            CK(cudaEventRecord(h->ev_pv[0], h->compute));
            pvalue_kernel<<<(unsigned)((mq + 31) / 32), 32, 0, h->compute>>>(dout[4], dout[5], 0, (long long)mq, (double)nu, lnbeta);
            CK(cudaGetLastError());
            CK(cudaEventRecord(h->ev_pv[1], h->compute));
        }
        CK(cudaEventRecord(t1, h->compute));
        if (!on_device) {
            // two D2H copies into pinned staging, then host memcpy into the caller's (pageable) arrays
            CK(cudaMemcpyAsync(h->res_host, dtmp, sizeof(double) * ncols * mq, cudaMemcpyDeviceToHost, h->compute));
            CK(cudaMemcpyAsync(h->res_host + sizeof(double) * ncols * mq, itmp, sizeof(int) * 3 * mq,
                               cudaMemcpyDeviceToHost, h->compute));
        }
        CK(cudaEventRecord(t2, h->compute));
        CK(cudaStreamSynchronize(h->compute));
        CK(cudaStreamSynchronize(h->copy));
        // the rotation pinned its digit planes in the L2 set-aside (rotate_i8_tc2.cuh): hand the lines back
        if (rotate && tc2::persist_planes() && n_rot_launch > 0) cudaCtxResetPersistingL2Cache();
        if (!on_device) {
            const double* hd = reinterpret_cast<const double*>(h->res_host);
            const int* hi = reinterpret_cast<const int*>(h->res_host + sizeof(double) * ncols * mq);
            for (int i = 0; i < 6; ++i) memcpy(out_user[i], hd + (size_t)i * mq, sizeof(double) * mq);
            if (want_lrt) for (int i = 0; i < 4; ++i) memcpy(lrt_user[i], hd + (size_t)(6 + i) * mq, sizeof(double) * mq);
            if (status) memcpy(status, hi, sizeof(int) * mq);
            if (e2) memcpy(e2, hi + mq, sizeof(int) * mq);
            if (e3) memcpy(e3, hi + 2 * mq, sizeof(int) * mq);
        }
        return PG_OK;
    }();
    h->defer_pvalues = false;

    if (rc == PG_OK && timing) {
        float v = 0;
        cudaEventElapsedTime(&timing->total_ms, t0, t2);
        cudaEventElapsedTime(&timing->d2h_ms, t1, t2);
        for (long long b = 0; b < nblocks; ++b) {
            cudaEventElapsedTime(&v, ev_conv[b].a, ev_conv[b].b); timing->convert_ms += v;
            cudaEventElapsedTime(&v, ev_rot[b].a, ev_rot[b].b); timing->rotate_ms += v;
            cudaEventElapsedTime(&v, ev_reml[b].a, ev_reml[b].b); timing->reml_ms += v;
            if (b == 0 && compressed && mq > 0) { cudaEventElapsedTime(&v, h->ev_pv[0], h->ev_pv[1]); timing->reml_ms += v; }  // the single p-value launch
            if (compressed) { cudaEventElapsedTime(&v, ev_cmp[b].a, ev_cmp[b].b); timing->compress_ms += v; }
            if (!on_device) { cudaEventElapsedTime(&v, ev_h2d[b].a, ev_h2d[b].b); timing->h2d_ms += v; }
        }
        timing->n_blocks = (int32_t)nblocks;
        timing->block_snps = (int32_t)blk;
        timing->reml_launches = (int32_t)nblocks;
        timing->rotate_launches = n_rot_launch;
        timing->convert_launches = rotate ? 0 : (int32_t)nblocks;  // staging kernels of the rotation are in rotate_launches
        timing->rot_engine = (n_fused_blocks == nblocks && nblocks > 0) ? PG_ROT_I8TC_MOMENTS : last_engine;
        timing->reml_engine = compressed ? PG_REML_COMPRESSED : h->engine;
        timing->n_nodes = compressed ? h->plan.Kc : h->n;
        // per block: compress (dmma / copy), fixed-lambda x rows, fixed-lambda evaluations + solve per phenotype; one
        // p-value launch per scan
        if (compressed)
            timing->reml_launches = (int32_t)(nblocks * (1 + 2 * q) + (nblocks - n_fused_blocks) *   // fused blocks: no compression
                                              ((h->plan.nitems ? 1 : 0) + (h->plan.ncopy ? 1 : 0)) + 1);
        else
            timing->reml_launches = (int32_t)(nblocks * q);
    }
    if (rc != PG_OK) {
        cudaStreamSynchronize(h->compute);
        cudaStreamSynchronize(h->copy);
        cudaStreamSynchronize(h->aux);
        cudaStreamSynchronize(h->cmb);
    }
    return rc;
}

extern "C" int pg_scan(pg_handle* h, const void* X, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                       double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
                       int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing)
{
    double* out[6] = {beta, se_beta, tau, lambda, F_wald, p_wald};
    return scan_impl(h, X, xdtype, ld, layout, m, grid, out, status, n_eval2, n_eval3, timing, false);
}

extern "C" int pg_scan_lrt(pg_handle* h, const void* X, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                           double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
                           double* lambda_ml, double* loglik_ml, double* D_lrt, double* p_lrt, int32_t* status,
                           int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing)
{
    double* out[6] = {beta, se_beta, tau, lambda, F_wald, p_wald};
    double* lrt[4] = {lambda_ml, loglik_ml, D_lrt, p_lrt};
    return scan_impl(h, X, xdtype, ld, layout, m, grid, out, status, n_eval2, n_eval3, timing, false, lrt);
}

extern "C" int pg_scan_device(pg_handle* h, const void* X_dev, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                              double* beta, double* se_beta, double* tau, double* lambda, double* F_wald,
                              double* p_wald, int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing)
{
    double* out[6] = {beta, se_beta, tau, lambda, F_wald, p_wald};
    return scan_impl(h, X_dev, xdtype, ld, layout, m, grid, out, status, n_eval2, n_eval3, timing, true);
}

// ------------------------------------------------------------------------------------------------
// Genetic relatedness matrix on the device ("next" row of the scope table: the step right before the path).
// Reference: calculate_genetic_relatedness_matrix, experiments/animal_gwas/run_gwas.py:45-55 --
//   sd = np.std(X, axis=0); sd[sd == 0] = 1; Z = (X - X.mean(axis=0)) / sd; K = Z @ Z.T / X.shape[1]
// (the same K = X_s X_s^T / p the other callers build after StandardScaler, experiments/wtccc/run_pygemma.py:432,:447).
// SNP blocks are standardised on the device (two-pass mean / variance like NumPy) and accumulated with FP64 SYRK.
// ------------------------------------------------------------------------------------------------
constexpr int kGrmChunk = 512;

// partial sums of one SNP column over a chunk of samples: pass 0 -> sum(x), pass 1 -> sum((x - mean)^2)
template <typename T>
__global__ void __launch_bounds__(128) grm_partial_kernel(const T* __restrict__ src, long long ld, int layout, int n,
                                                           long long pb, const double* __restrict__ mean, int pass,
                                                           double* __restrict__ part)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pb) return;
    const int j0 = blockIdx.y * kGrmChunk, j1 = min(n, j0 + kGrmChunk);
    const size_t step = layout == 0 ? (size_t)ld : 1, base = layout == 0 ? (size_t)g : (size_t)g * ld;
    const double mu = pass ? mean[g] : 0.0;
    double acc = 0.0;
    for (int j = j0; j < j1; ++j) {
        const double x = (double)src[base + (size_t)j * step];
        const double d = x - mu;
        acc += pass ? d * d : x;
    }
    part[(size_t)blockIdx.y * pb + g] = acc;
}

// pass 0: mean[g] = sum / n; pass 1: isd[g] = 1 / sqrt(sum / n), 1 when the column is constant (sd[sd == 0] = 1)
__global__ void grm_reduce_kernel(const double* __restrict__ part, int nchunks, long long pb, int n, int pass,
                                  double* __restrict__ out)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pb) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[(size_t)c * pb + g];
    if (pass == 0) {
        out[g] = s / (double)n;
    } else {
        const double sd = sqrt(s / (double)n);
        out[g] = sd == 0.0 ? 1.0 : 1.0 / sd;
    }
}

// Z^T block, column-major (pb x n, ld = pb): element (g, j) = (x_jg - mean_g) * isd_g
template <typename T>
__global__ void __launch_bounds__(128) grm_standardise_kernel(const T* __restrict__ src, long long ld, int layout, int n,
                                                               long long pb, const double* __restrict__ mean,
                                                               const double* __restrict__ isd, double* __restrict__ zt)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pb) return;
    const int j0 = blockIdx.y * 64, j1 = min(n, j0 + 64);
    const double mu = mean[g], is = isd[g];
    for (int j = j0; j < j1; ++j) {
        const size_t si = layout == 0 ? (size_t)j * ld + g : (size_t)g * ld + j;
        zt[(size_t)j * pb + g] = ((double)src[si] - mu) * is;
    }
}

__global__ void grm_symmetrise_kernel(double* __restrict__ K, int n)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int c = (int)(idx / n), r = (int)(idx % n);   // column-major element (r, c)
    if (r < c) K[idx] = K[(size_t)r * n + c];           // upper <- lower
}

template <typename T>
static int grm_block(pg_handle* h, const T* dsrc, long long ld_dev, int layout, long long pb, double* part, double* mean,
                     double* isd, double* zt)
{
    const int n = h->n;
    const int nchunks = (n + kGrmChunk - 1) / kGrmChunk;
    const unsigned gb = (unsigned)((pb + 127) / 128);
    for (int pass = 0; pass < 2; ++pass) {
        grm_partial_kernel<T><<<dim3(gb, nchunks), 128, 0, h->compute>>>(dsrc, ld_dev, layout, n, pb, mean, pass, part);
        CK(cudaGetLastError());
        grm_reduce_kernel<<<gb, 128, 0, h->compute>>>(part, nchunks, pb, n, pass, pass ? isd : mean);
        CK(cudaGetLastError());
    }
    grm_standardise_kernel<T><<<dim3(gb, (unsigned)((n + 63) / 64)), 128, 0, h->compute>>>(dsrc, ld_dev, layout, n, pb, mean, isd, zt);
    CK(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_grm(pg_handle* h, const void* X, int xdtype, int64_t ld, int layout, int64_t p, double* K_host_out,
                      int set_kinship, double* d_out_host, float* grm_ms, float* eig_ms)
{
    if (!h || !X || p <= 0) return fail(h, PG_ERR_ARG, "pg_grm: NULL X or p <= 0");
    if (xdtype < PG_X_I8 || xdtype > PG_X_F64) return fail(h, PG_ERR_ARG, "pg_grm: xdtype %d", xdtype);
    if (layout != PG_X_SAMPLE_MAJOR && layout != PG_X_SNP_MAJOR) return fail(h, PG_ERR_ARG, "pg_grm: layout %d", layout);
    const int n = h->n;
    if ((layout == PG_X_SAMPLE_MAJOR && ld < p) || (layout == PG_X_SNP_MAJOR && ld < n)) return fail(h, PG_ERR_ARG, "pg_grm: ld too small");
    CK(cudaSetDevice(h->device));
    const size_t esz = xdtype_size(xdtype);
    long long pb = (long long)((size_t(1) << 30) / (sizeof(double) * (size_t)n));
    pb = std::max<long long>(256, std::min<long long>(pb, p));
    const int nchunks = (n + kGrmChunk - 1) / kGrmChunk;
    double *Kd = nullptr, *zt = nullptr, *part = nullptr, *mean = nullptr, *isd = nullptr;
    void* raw = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = [&]() -> int {
        CK(cudaMalloc(&Kd, sizeof(double) * (size_t)n * n));
        CK(cudaMalloc(&zt, sizeof(double) * (size_t)n * pb));
        CK(cudaMalloc(&part, sizeof(double) * (size_t)nchunks * pb));
        CK(cudaMalloc(&mean, sizeof(double) * pb));
        CK(cudaMalloc(&isd, sizeof(double) * pb));
        CK(cudaMalloc(&raw, esz * (size_t)n * pb));
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, h->compute));
        const double alpha = 1.0 / (double)p;
        for (long long g0 = 0; g0 < p; g0 += pb) {
            const long long cnt = std::min(pb, p - g0);
            long long ld_dev;
            if (layout == PG_X_SAMPLE_MAJOR) {
                CK(cudaMemcpy2DAsync(raw, (size_t)cnt * esz, (const char*)X + (size_t)g0 * esz, (size_t)ld * esz,
                                     (size_t)cnt * esz, (size_t)n, cudaMemcpyHostToDevice, h->compute));
                ld_dev = cnt;
            } else {
                CK(cudaMemcpy2DAsync(raw, (size_t)n * esz, (const char*)X + (size_t)g0 * ld * esz, (size_t)ld * esz,
                                     (size_t)n * esz, (size_t)cnt, cudaMemcpyHostToDevice, h->compute));
                ld_dev = n;
            }
            int r2;
            if (xdtype == PG_X_I8) r2 = grm_block<int8_t>(h, (const int8_t*)raw, ld_dev, layout, cnt, part, mean, isd, zt);
            else if (xdtype == PG_X_F32) r2 = grm_block<float>(h, (const float*)raw, ld_dev, layout, cnt, part, mean, isd, zt);
            else r2 = grm_block<double>(h, (const double*)raw, ld_dev, layout, cnt, part, mean, isd, zt);
            if (r2) return r2;
            // zt is Z^T (cnt x n, column-major): K(lower) += (1/p) Z Z^T = (1/p) (Z^T)^T (Z^T)
            const double beta = g0 == 0 ? 0.0 : 1.0;
            CKB(cublasDsyrk(h->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n, (int)cnt, &alpha, zt, (int)cnt, &beta, Kd, n));
        }
        const size_t total = (size_t)n * n;
        grm_symmetrise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->compute>>>(Kd, n);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, h->compute));
        if (K_host_out) CK(cudaMemcpyAsync(K_host_out, Kd, sizeof(double) * total, cudaMemcpyDeviceToHost, h->compute));
        CK(cudaStreamSynchronize(h->compute));
        if (grm_ms) cudaEventElapsedTime(grm_ms, e0, e1);
        return PG_OK;
    }();
    if (rc == PG_OK && set_kinship) rc = set_kinship_device(h, Kd, d_out_host, eig_ms);
    void* bufs[] = {Kd, zt, part, mean, isd, raw};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

extern "C" int pg_probe_precompute(pg_handle* h, const double* x_rot_host, double lam, int fixed_index, int full,
                                   double* out9)
{
    if (!h || !x_rot_host || !out9) return fail(h, PG_ERR_ARG, "pg_probe_precompute: NULL argument");
    if (!h->have_design) return fail(h, PG_ERR_ARG, "pg_probe_precompute: call pg_set_design first");
    CK(cudaSetDevice(h->device));
    const int n = h->n;
    DevMem dx_g, dout_g, dz_g;
    CK(cudaMalloc(&dx_g.p, sizeof(double) * h->ldx));
    CK(cudaMalloc(&dout_g.p, sizeof(double) * 9));
    double *dx = dx_g.as<double>(), *dout = dout_g.as<double>();
    CK(cudaMemsetAsync(dx, 0, sizeof(double) * h->ldx, h->compute));
    std::vector<double> xs(n);
    for (int l = 0; l < n; ++l) xs[l] = x_rot_host[h->perm_identity ? l : h->perm[l]];
    CK(cudaMemcpyAsync(dx, xs.data(), sizeof(double) * n, cudaMemcpyHostToDevice, h->compute));
    if (h->engine == PG_REML_AUTO || h->engine == PG_REML_COMPRESSED) {
        const DevPlan& P = h->plan;
        const int k1p = h->k1p, zrows = k1p - 1 + h->q;   // probes phenotype 0
        CK(cudaMalloc(&dz_g.p, sizeof(double) * (size_t)zrows * P.Kcp));
        double* dz = dz_g.as<double>();
        CK(cudaMemsetAsync(dz, 0, sizeof(double) * (size_t)zrows * P.Kcp, h->compute));
        if (P.nitems) {
            CK(cudaFuncSetAttribute(compress_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtSmemBytes));
            compress_dmma_kernel<<<(unsigned)P.nitems, 256, kCtSmemBytes, h->compute>>>(dx, h->ldx, 1, P.items, P.V, P.vpitch,
                                                                                     h->c0, k1p, zrows, P.Kcp, dz, 1);
            CK(cudaGetLastError());
        }
        if (P.ncopy) {
            compress_copy_kernel<<<dim3((unsigned)((P.ncopy + 127) / 128), 1), 128, 0, h->compute>>>(
                dx, h->ldx, 1, P.copy_l, P.copy_node, P.ncopy, h->q > 1 ? h->wy_all : h->wy, h->ldw, h->c0, P.klin, k1p, zrows,
                P.Kcp, dz);
            CK(cudaGetLastError());
        }
        SolveArgs sa{};
        sa.n = n; sa.c0 = h->c0; sa.grid = 0; sa.m = 1; sa.row0 = 0; sa.nodes = P.nodes; sa.Kcp = P.Kcp; sa.Kc = P.Kc; sa.Z = dz;
        sa.k1p = k1p; sa.t2 = h->tab2; sa.zrows = zrows; sa.yrow = k1p - 1; sa.FX = nullptr; sa.ldF = 0; sa.swap = 0; sa.zsm = 0;
        for (int i = 0; i < 4; ++i) sa.lrt[i] = nullptr;
        sa.l_null = 0.0;
        const size_t smemc = sizeof(double) * (3 * (size_t)k1p + h->tab2.NF2);
        const bool two = (h->c0 + 2) > 32;
        if (smemc > kDynSmemOptIn) {
            if (two) CK(cudaFuncSetAttribute(probe_precompute_compressed_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc));
            else CK(cudaFuncSetAttribute(probe_precompute_compressed_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc));
        }
        if (two) probe_precompute_compressed_kernel<2><<<1, 32, smemc, h->compute>>>(sa, lam, fixed_index, full, dout);
        else probe_precompute_compressed_kernel<1><<<1, 32, smemc, h->compute>>>(sa, lam, fixed_index, full, dout);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out9, dout, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->compute));
        CK(cudaStreamSynchronize(h->compute));
        return PG_OK;
    }
    ScanArgs a{};
    a.n = n; a.c0 = h->c0; a.grid = 0; a.m = 1; a.row0 = 0; a.d = h->d; a.wy = h->wy; a.ldw = h->ldw; a.xr = dx; a.ldx = h->ldx; a.tab = h->tab;
    const int k = h->c0 + 2, TT = k * (k + 1) / 2;
    const size_t smem = sizeof(double) * 3 * TT;
    if (smem > kDynSmemOptIn)
        CK(cudaFuncSetAttribute(probe_precompute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_precompute_kernel<<<1, 32, smem, h->compute>>>(a, lam, fixed_index, full, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out9, dout, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    return PG_OK;
}

extern "C" int pg_probe_f_sf(pg_handle* h, const double* F_host, double nu, int64_t k, double* p_host)
{
    if (!h || !F_host || !p_host || k < 0) return fail(h, PG_ERR_ARG, "pg_probe_f_sf: bad argument");
    if (k == 0) return PG_OK;
    CK(cudaSetDevice(h->device));
    DevMem dF_g, dp_g;
    CK(cudaMalloc(&dF_g.p, sizeof(double) * k));
    CK(cudaMalloc(&dp_g.p, sizeof(double) * k));
    double *dF = dF_g.as<double>(), *dp = dp_g.as<double>();
    CK(cudaMemcpyAsync(dF, F_host, sizeof(double) * k, cudaMemcpyHostToDevice, h->compute));
    probe_f_sf_kernel<<<(unsigned)((k + 127) / 128), 128, 0, h->compute>>>(dF, nu, k, dp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(p_host, dp, sizeof(double) * k, cudaMemcpyDeviceToHost, h->compute));
    CK(cudaStreamSynchronize(h->compute));
    return PG_OK;
}

extern "C" int pg_probe_rotated(pg_handle* h, double* xr_host, int64_t count, int64_t* row0)
{
    if (!h || !xr_host || count < 0) return fail(h, PG_ERR_ARG, "pg_probe_rotated: bad argument");
    if (!h->last_xr || h->last_block_count == 0) return fail(h, PG_ERR_ARG, "pg_probe_rotated: no scan has run");
    CK(cudaSetDevice(h->device));
    const long long c = std::min<long long>(count, h->last_block_count);
    CK(cudaMemcpy2D(xr_host, sizeof(double) * h->n, h->last_xr, sizeof(double) * h->ldx, sizeof(double) * h->n, (size_t)c,
                    cudaMemcpyDeviceToHost));
    if (!h->perm_identity) {  // report in the caller's eigen order
        std::vector<double> row(h->n);
        for (long long g = 0; g < c; ++g) {
            double* r = xr_host + (size_t)g * h->n;
            for (int l = 0; l < h->n; ++l) row[h->perm[l]] = r[l];
            std::copy(row.begin(), row.end(), r);
        }
    }
    if (row0) *row0 = h->last_block_row0;
    return PG_OK;
}


// ------------------------------------------------------------------------------------------------
// Several GPUs, one process (include/pygemma_b200.h: pg_multi_*).  SNP shards are independent units: one host thread per
// device runs the ordinary blocked scan on its contiguous column range and writes straight into the caller's arrays.
// ------------------------------------------------------------------------------------------------
struct pg_multi {
    std::vector<pg_handle*> hs;
    std::string err;
};

static int mfail(pg_multi* mh, int code, const std::string& msg)
{
    if (mh) mh->err = msg; else g_create_error = msg;
    return code;
}

extern "C" const char* pg_multi_last_error(const pg_multi* mh) { return mh ? mh->err.c_str() : g_create_error.c_str(); }
extern "C" int pg_multi_count(const pg_multi* mh) { return mh ? (int)mh->hs.size() : 0; }
extern "C" pg_handle* pg_multi_handle(pg_multi* mh, int i) { return (mh && i >= 0 && i < (int)mh->hs.size()) ? mh->hs[i] : nullptr; }

extern "C" int pg_multi_destroy(pg_multi* mh)
{
    if (!mh) return PG_OK;
    for (pg_handle* h : mh->hs) pg_destroy(h);
    delete mh;
    return PG_OK;
}

extern "C" int pg_multi_create(int n, int c0, int ngpu, const int* devices, pg_multi** out)
{
    if (!out) return mfail(nullptr, PG_ERR_ARG, "pg_multi_create: out is NULL");
    *out = nullptr;
    if (ngpu < 1) return mfail(nullptr, PG_ERR_ARG, "pg_multi_create: ngpu must be >= 1");
    pg_multi* mh = new pg_multi;
    for (int i = 0; i < ngpu; ++i) {
        pg_handle* h = nullptr;
        const int rc = pg_create(n, c0, devices ? devices[i] : i, &h);
        if (rc != PG_OK) {
            const std::string msg = g_create_error;
            pg_multi_destroy(mh);
            return mfail(nullptr, rc, "pg_multi_create: device " + std::to_string(devices ? devices[i] : i) + ": " + msg);
        }
        mh->hs.push_back(h);
    }
    *out = mh;
    return PG_OK;
}

// runs fn(i) for every handle on its own host thread; the first failure is reported
template <typename F>
static int for_each_device(pg_multi* mh, F fn)
{
    const int g = (int)mh->hs.size();
    std::vector<int> rcs(g, PG_OK);
    std::vector<std::thread> th;
    for (int i = 1; i < g; ++i) th.emplace_back([&, i] { rcs[i] = fn(i); });
    rcs[0] = fn(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < g; ++i)
        if (rcs[i] != PG_OK) return mfail(mh, rcs[i], "device " + std::to_string(mh->hs[i]->device) + ": " + mh->hs[i]->err);
    return PG_OK;
}

static int multi_share_eigen(pg_multi* mh, float* bcast_ms)
{
    const auto t0 = std::chrono::steady_clock::now();
    // U (8 n^2 bytes) and d from the first device to every other one, peer to peer, concurrently
    const int rc = for_each_device(mh, [&](int i) { return i == 0 ? PG_OK : pg_copy_eigen(mh->hs[i], mh->hs[0]); });
    if (bcast_ms) *bcast_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

extern "C" int pg_multi_set_kinship(pg_multi* mh, const double* K_host, double* d_out_host, float* eig_ms, float* bcast_ms)
{
    if (!mh) return PG_ERR_ARG;
    const int rc = pg_set_kinship(mh->hs[0], K_host, d_out_host, eig_ms);
    if (rc != PG_OK) return mfail(mh, rc, mh->hs[0]->err);
    return multi_share_eigen(mh, bcast_ms);
}

extern "C" int pg_multi_set_eigen(pg_multi* mh, const double* U_host, int u_row_major, const double* d_host)
{
    if (!mh) return PG_ERR_ARG;
    const int rc = pg_set_eigen(mh->hs[0], U_host, u_row_major, d_host);
    if (rc != PG_OK) return mfail(mh, rc, mh->hs[0]->err);
    return multi_share_eigen(mh, nullptr);
}

extern "C" int pg_multi_set_design(pg_multi* mh, const double* W_host, const double* y_host, int already_rotated, float* ms)
{
    if (!mh) return PG_ERR_ARG;
    std::vector<float> t(mh->hs.size(), 0.f);
    const int rc = for_each_device(mh, [&](int i) { return pg_set_design(mh->hs[i], W_host, y_host, already_rotated, &t[i]); });
    if (ms) *ms = *std::max_element(t.begin(), t.end());
    return rc;
}

extern "C" int pg_multi_scan(pg_multi* mh, const void* X, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                             double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
                             int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing)
{
    if (!mh) return PG_ERR_ARG;
    if (!X || m < 0) return mfail(mh, PG_ERR_ARG, "pg_multi_scan: NULL X or negative m");
    if (xdtype < PG_X_I8 || xdtype > PG_X_BED) return mfail(mh, PG_ERR_ARG, "pg_multi_scan: xdtype");
    const int g = (int)mh->hs.size();
    for (pg_handle* h : mh->hs)
        if (h->q != 1) return mfail(mh, PG_ERR_ARG, "pg_multi_scan: one phenotype per pass");
    // contiguous shards, ceil(m / g) columns each (SampleIter), rounded up to 128 columns once shards are that long
    long long per = (m + g - 1) / g;
    if (g > 1 && per >= 128) {
        const long long al = (per + 127) / 128 * 128;
        if (al * (g - 1) < m) per = al;
    }
    const size_t esz = xdtype_size(xdtype);
    return for_each_device(mh, [&](int i) -> int {
        const long long a = std::min<long long>((long long)i * per, m), b = std::min<long long>((long long)(i + 1) * per, m);
        if (timing) memset(&timing[i], 0, sizeof(pg_timing));
        if (b <= a) return PG_OK;
        const char* Xi = layout == PG_X_SAMPLE_MAJOR ? (const char*)X + (size_t)a * esz : (const char*)X + (size_t)a * (size_t)ld * esz;
        double* out[6] = {beta + a, se_beta + a, tau + a, lambda + a, F_wald + a, p_wald + a};
        return scan_impl(mh->hs[i], Xi, xdtype, ld, layout, b - a, grid, out, status ? status + a : nullptr,
                         n_eval2 ? n_eval2 + a : nullptr, n_eval3 ? n_eval3 + a : nullptr, timing ? &timing[i] : nullptr, false);
    });
}
