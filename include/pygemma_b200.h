/*
 * pygemma_b200.h -- C ABI of the B200-native per-SNP LMM association scan.
 *
 * This is the drop-in boundary for the hot path of rlangefe/pygemma:
 *
 *   lmm.pygemma(Y, X, W, K, ...)                      reference lmm/lmm.py:87
 *     eigh(K), clip d >= 0                            lmm/lmm.py:151-162   -> pg_set_kinship / pg_set_eigen
 *     U.T @ Y, U.T @ W                                lmm/lmm.py:245-246   -> pg_set_design
 *     U.T @ X                                         lmm/lmm.py:244       -> pg_scan (rotation stage)
 *     Pool.imap(calculate, SampleIter(...))           lmm/lmm.py:378-403   -> pg_scan (REML stage)
 *       calc_lambda_restricted                        pygemma_model.pyx:64
 *       precompute_mat                                pygemma_model.pyx:880
 *       newton, brentq, derivative/likelihood formulas  pyx:1349, :176, :1656-1698, :1813
 *       calc_beta_vg_ve_restricted_overload           pyx:1514
 *       F_wald, stats.f.sf                            lmm/lmm.py:471,:482
 *
 * Conventions
 *   - every function returns 0 on success and a negative pg_status on failure;
 *     pg_last_error() gives the message (CUDA / cuSOLVER / cuBLAS text included).
 *   - all pointers are caller-owned; host entry points are synchronous at return.
 *   - a handle is bound to one CUDA device and is not thread-safe.
 *   - there is no CPU fallback: without a usable CUDA device pg_create fails.
 *   - per-SNP failure (status[g] & 1) yields a NaN row, never an error return,
 *     like the reference's LinAlgError branch (lmm/lmm.py:484-493).  Bits 1 and 2 of status are notes, not
 *     failures: 2 = the ML Newton iteration of the LRT outputs was stopped at the edge of the tabulated lambda range,
 *     4 = a bracket with a NaN derivative at one end was skipped (the other candidates still competed).
 */
#ifndef PYGEMMA_B200_H
#define PYGEMMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 8

typedef struct pg_handle pg_handle;

typedef enum {
    PG_OK = 0,
    PG_ERR_ARG = -1,      /* bad argument / call order */
    PG_ERR_CUDA = -2,     /* CUDA runtime error */
    PG_ERR_CUSOLVER = -3, /* eigendecomposition failed */
    PG_ERR_CUBLAS = -4,
    PG_ERR_NO_DEVICE = -5,
    PG_ERR_ALLOC = -6
} pg_status;

/* element type of the genotype matrix handed to pg_scan */
typedef enum {
    PG_X_I8 = 0,
    PG_X_F32 = 1,
    PG_X_F64 = 2,
    PG_X_BED = 3 /* packed PLINK .bed body (2 bits per genotype): SNP-major only, ld = bytes per SNP >= ceil(n/4);
                    missing genotypes are mean-imputed on the device, see pg_set_bed_options */
} pg_xdtype;

/* memory layout of the genotype matrix */
typedef enum {
    PG_X_SAMPLE_MAJOR = 0, /* (n, m) C-order as NumPy passes it: element (j, g) at X[j*ld + g], ld >= m */
    PG_X_SNP_MAJOR = 1     /* (m, n): element (j, g) at X[g*ld + j], ld >= n */
} pg_xlayout;

/* rotation engine (stage 1 of pg_scan) */
typedef enum {
    PG_ROT_AUTO = 0,  /* integer dosages -> exact int8-split tensor-core path when available, else FP64 */
    PG_ROT_FP64 = 1,  /* FP64 GEMM */
    PG_ROT_I8SPLIT = 2, /* exact int8 split: cuBLAS int8 GEMM + recombination kernel */
    PG_ROT_I8TC = 3,    /* exact int8 split as one hand-written TMA + tcgen05 kernel with the recombination fused in */
    PG_ROT_I8TC_MOMENTS = 4 /* reported in pg_timing.rot_engine only (not selectable): the PG_ROT_I8TC kernel with the
                               eigenvalue-space moments fused in -- see pg_set_moment_fusion */
} pg_rotation;

/* REML stage engine (stage 2 of pg_scan) */
typedef enum {
    PG_REML_AUTO = 0,       /* = PG_REML_COMPRESSED */
    PG_REML_COMPRESSED = 1, /* eigenvalue-space compression (FP64 tensor-pipe contraction) + optimiser on the moments */
    PG_REML_STREAM = 2,     /* optimiser streams the full rotated genotype vector on every evaluation (CTA lock step) */
    PG_REML_WARP = 3        /* as STREAM, one independent warp per SNP (cross-check engine) */
} pg_reml_engine;

/* what a scanned column is (pg_set_scan_mode) */
typedef enum {
    PG_SCAN_WALD = 0, /* default: the column is the tested regressor x, the design's y the phenotype (lmm/lmm.py:461 calculate) */
    PG_SCAN_DE = 1    /* "differential expression": the column is the PHENOTYPE and the design's y the tested regressor --
                         the reference's de=True / calculate_de (lmm/lmm.py:498-532): lambda from
                         calc_lambda_restricted(d, X[:, g], [W, Y]), beta / se / tau from
                         calc_beta_vg_ve_restricted_overload(d, W, Y, lambda, X[:, g]).  (The reference's own driver raises before
                         reaching it -- SampleIter yields a 5-tuple, lmm/lmm.py:434, calculate_de unpacks 4, :499 -- this is the
                         computation those lines spell out.) */
} pg_scan_mode;

/* device-side timings of the last pg_scan call, milliseconds (CUDA events) */
typedef struct {
    float total_ms;   /* whole call, first H2D to last D2H */
    float h2d_ms;     /* genotype uploads (overlapped with compute after the first block) */
    float convert_ms; /* dosage -> fp64 / layout change */
    float rotate_ms;  /* U^T X */
    float reml_ms;    /* per-SNP REML kernel */
    float d2h_ms;     /* result download */
    int32_t n_blocks; /* SNP blocks processed */
    int32_t block_snps;
    int32_t reml_launches, rotate_launches, convert_launches;
    int32_t rot_engine; /* engine the rotation used: PG_ROT_FP64 / PG_ROT_I8SPLIT, 0 when no rotation ran */
    float compress_ms;  /* part of reml_ms spent collapsing genotype vectors onto the compression nodes */
    int32_t reml_engine; /* engine the REML stage used */
    int32_t n_nodes;    /* compression nodes per SNP (n when the engine streams full vectors) */
} pg_timing;

int pg_abi_version(void);
/* digit planes of the int8-split rotation this library was built with (7; 6 with -DPG_SLICES=6): the rotation costs
 * planes * 2 * n^2 int8 operations per SNP */
int pg_rotation_planes(void);
int pg_device_count(int* count);

/* last error text of a handle (or of the failed pg_create when h == NULL) */
const char* pg_last_error(const pg_handle* h);

/* n samples, c0 covariate columns of W (intercept included by the caller, lmm/lmm.py:94) */
int pg_create(int n, int c0, int device, pg_handle** out);
int pg_destroy(pg_handle* h);

/*
 * Eigendecomposition K = U diag(d) U^T on the device (cuSOLVER syevd, FP64), eigenvalues clipped at 0.
 * Replaces scipy.linalg.eigh + np.maximum(0, .) of lmm/lmm.py:151-157.  K_host is n*n doubles
 * (symmetric, so row/column order is immaterial).  d_out_host (nullable) receives the n eigenvalues;
 * eig_ms (nullable) the device time.  This is setup: it is not part of the scan metric.
 */
int pg_set_kinship(pg_handle* h, const double* K_host, double* d_out_host, float* eig_ms);

/*
 * Genetic relatedness matrix on the device (the step right before the scan in the reference's callers):
 *   sd = std(X, axis=0); sd[sd == 0] = 1; Z = (X - mean(X, axis=0)) / sd; K = Z Z^T / p
 * (calculate_genetic_relatedness_matrix, experiments/animal_gwas/run_gwas.py:45-55; K = X_s X_s^T / p after
 * StandardScaler in experiments/wtccc/run_pygemma.py:432,:447).  X holds p marker columns for the handle's n samples
 * (host pointer; dtype / layout / ld as in pg_scan).  K_host_out (nullable) receives the n*n matrix; with
 * set_kinship != 0 the matrix stays on the device and is eigendecomposed in place as by pg_set_kinship
 * (d_out_host / eig_ms as there).  grm_ms (nullable): device time of the standardisation + FP64 SYRK.
 */
int pg_grm(pg_handle* h, const void* X, int xdtype, int64_t ld, int layout, int64_t p, double* K_host_out,
           int set_kinship, double* d_out_host, float* grm_ms, float* eig_ms);

/*
 * Supply an eigendecomposition computed elsewhere.  u_row_major = 1: U[j*n + i] is component j of
 * eigenvector i (NumPy C-order result of eigh); 0: column-major (LAPACK / cuSOLVER order).
 * U may be NULL when the inputs of pg_set_design / pg_scan are already rotated (eigen=False,
 * lmm/lmm.py:164-167,:243): then d is the eigenvalue vector and is clipped at 0 like the reference does.
 */
int pg_set_eigen(pg_handle* h, const double* U_host, int u_row_major, const double* d_host);
int pg_set_eigen_device(pg_handle* h, const double* U_dev, int u_row_major, const double* d_dev);
/* copy the handle's U (column-major) and d into caller-provided device buffers (multi-GPU broadcast) */
int pg_get_eigen_device(pg_handle* h, double* U_dev_out, double* d_dev_out);

/*
 * Take over the eigen-system another handle holds (same n; same or peer device), device to device.  A caller that keeps
 * one eigendecomposition for many calls -- the reference's callers repeat eigh(K) for every phenotype and covariate set
 * (lmm/lmm.py:151-162 runs inside every lmm.pygemma call; experiments/animal_gwas/run_gwas.py:167-175 loops it) -- creates
 * a handle per covariate count c0 and copies U, d instead of decomposing again (pygemma_b200.lmm.factorize).
 */
int pg_copy_eigen(pg_handle* h, const pg_handle* src);

/*
 * Covariates and phenotype: W_host is (n, c0) C-order, y_host is n doubles.  Computes U^T W and U^T y
 * (lmm/lmm.py:245-246) unless already_rotated, and builds the SNP-independent lambda tables.
 */
int pg_set_design(pg_handle* h, const double* W_host, const double* y_host, int already_rotated, float* ms);

/*
 * Several phenotypes on one eigen-system and one covariate matrix: Y_host is (n, q) C-order (the reference's 2-D Y,
 * lmm/lmm.py:108-110 takes one column; a caller loops lmm.pygemma over traits, e.g.
 * experiments/benchmarks/benchmarks.py).  Every later pg_scan / pg_scan_device rotates each genotype block ONCE and
 * runs the REML stage per phenotype: all output arrays then hold q * m values, phenotype-major
 * (out[ph * m + g]); each phenotype's rows are bit-identical to a pg_set_design + pg_scan of that column alone.
 * pg_set_design resets the handle to one phenotype.  1 <= q <= PG_MAX_TRAITS per pass (the moment slab of a SNP
 * grows by one row per trait); callers with more traits scan the genotypes once per group of PG_MAX_TRAITS
 * (pygemma_b200.lmm.pygemma_multi does), which keeps the rotation below 10 % of the pass.
 */
#define PG_MAX_TRAITS 64
int pg_set_design_multi(pg_handle* h, const double* W_host, const double* Y_host, int q, int already_rotated,
                        float* ms);

/*
 * Run all device work of this handle on a caller-owned CUDA stream (a cudaStream_t passed as void*),
 * e.g. torch's current stream, so the caller can bracket calls with its own CUDA events.  NULL restores
 * the handle's private stream.  Genotype uploads still use a private copy stream, ordered by events.
 */
int pg_set_stream(pg_handle* h, void* stream);

/* rotation engine selection and block size (0 = automatic) */
int pg_set_options(pg_handle* h, int rotation, int64_t block_snps);
/*
 * Decoding of PG_X_BED input.  count_a1 != 0: dosage = number of A1 alleles (00 -> 2, 10 -> 1, 11 -> 0), else number of
 * A2 alleles (pysnptools Bed(count_A1=...)); the code 01 is missing and is replaced by the column mean of the observed
 * dosages (SimpleImputer(strategy='mean'), experiments/benchmarks/benchmarks.py:22).  standardize != 0: the imputed
 * column is centred and divided by its population standard deviation (1 when it is constant), as StandardScaler does
 * (experiments/wtccc/run_pygemma.py:432).  Default: count_a1 = 0, standardize = 0.
 */
int pg_set_bed_options(pg_handle* h, int count_a1, int standardize);
/* Role of the scanned columns for every later pg_scan / pg_scan_device (default PG_SCAN_WALD); PG_SCAN_DE needs the
 * compressed REML engine.  Outputs keep their names: beta / se_beta are the effect of y on each column. */
int pg_set_scan_mode(pg_handle* h, int mode);
/* REML stage engine selection (default PG_REML_AUTO); all engines give the same results to rounding */
int pg_set_reml_engine(pg_handle* h, int engine);
/*
 * Moment fusion.  The rotated genotypes U^T x (lmm/lmm.py:244) are consumed only by the per-SNP moment sums the reference
 * forms in precompute_mat (pyx:938-943, :1002); on the compressed REML engine those are linear (x.w_j, x.y) or quadratic
 * (x.x) in U^T x.  With fusion the PG_ROT_I8TC kernel produces them itself: the linear moments as extra exact int8 tiles
 * against G = U V (V = the compression operand, built once per design), the x^2 moments in its epilogue; rotated genotypes
 * are never written and the compression kernels do not run.  Results agree with the unfused path to rounding (different
 * summation order), and are deterministic.  mode: 1 = wherever the engine allows it (PG_ROT_AUTO or PG_ROT_I8TC, int8 /
 * level-coded / .bed genotypes, the compressed REML engine); 0 = never; -1 (default) = where it was measured to pay on a
 * power-capped B200 (profiles/experiments_r02.md): PG_ROT_AUTO, n >= 4096, at least 24 linear columns (c0 + traits) and
 * at most 140 compression nodes (+7 % SNPs/s at c0 = 40 with 131 nodes).  With fewer columns the fused kernel takes as
 * long as rotation + compression together; with more nodes its extra tiles cost more than the compression.  Blocks that need the second,
 * eps-weighted rotation pass (unequally spaced levels, missing .bed calls), the FP64 rotation or another engine are
 * compressed as before.
 * pg_probe_rotated is not available after a fused scan.
 */
int pg_set_moment_fusion(pg_handle* h, int mode);
/* What the next scan of the current design would do (host-only probe, all outputs nullable): *fused != 0 when the moments
 * come from the fused rotation; *g_columns = columns of G = U V appended to the rotation (COMPRESS nodes x linear columns),
 * *pieces = x^2 partial sums per SNP, *build_ms = device time of the last build of G and its digit planes (0 before it). */
int pg_probe_fusion(pg_handle* h, int32_t* fused, int32_t* g_columns, int32_t* pieces, float* build_ms);

/*
 * The scan: for each of the m genotype columns, rotate (unless the handle holds rotated inputs),
 * find the REML lambda (grid != 0: 12-point grid search, pyx:99-132; else bracket scan + Brent +
 * Newton, pyx:135-194) and compute the Wald statistics.  Outputs are m doubles each, in input
 * column order, no filtering (lmm/lmm.py:401-411).  status / n_eval2 / n_eval3 are nullable:
 * n_eval2 / n_eval3 count the passes over the rotated genotype vector with 2 / 3 powers of H^-1.
 * pg_scan takes host pointers; pg_scan_device takes device pointers for X and for every output.
 */
int pg_scan(pg_handle* h, const void* X, int xdtype, int64_t ld, int layout, int64_t m, int grid,
            double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
            int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing);
int pg_scan_device(pg_handle* h, const void* X_dev, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                   double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
                   int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing);

/*
 * Likelihood-ratio outputs next to the Wald scan.  The reference driver carries them as commented-out scaffolding
 * (result columns 'D_lrt', 'p_lrt', 'likelihood' at lmm/lmm.py:137-141; null model :176-190; per-SNP loop :278-300); the
 * functions that scaffolding calls are live in the reference and are what is computed here:
 *   lambda_ml  = lmm.calc_lambda(d, y, [W, x])                    ML lambda of the alternative model (lmm/lmm.py:22-84)
 *   loglik_ml  = likelihood_lambda(lambda_ml, d, y, [W, x])       (pyx:1542; = likelihood(lam, n / yPy, beta_hat, ..), :1736)
 *   D_lrt      = 2 (loglik_ml - l_null)                           (lmm/lmm.py:282)
 *   p_lrt      = chi-square(1) upper tail of D_lrt                (lmm/lmm.py:300, evaluated as a survival function)
 * with the null model of pg_null_model.  Host pointers, m (x q) doubles each; otherwise exactly pg_scan.
 */
int pg_scan_lrt(pg_handle* h, const void* X, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
                double* lambda_ml, double* loglik_ml, double* D_lrt, double* p_lrt,
                int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing);
/*
 * Null model [W] of phenotype `trait` of the current design (lmm/lmm.py:176-190): lambda_null = lmm.calc_lambda(d, y, W),
 * tau_null = n / y^T P y, loglik_null = likelihood(lambda_null, tau_null, beta_hat, d, y, W).  Outputs are nullable.
 */
int pg_null_model(pg_handle* h, int trait, double* lambda_null, double* tau_null, double* loglik_null);

/*
 * Several GPUs of one node driven by ONE process (SURVEY 8b's pg_create_multi): a pg_multi owns one handle per device.
 * The eigendecomposition runs on the first device and U, d travel to the others peer to peer (cudaMemcpyPeer: NVLink /
 * NVSwitch where the devices are connected), the design goes to every device, and pg_multi_scan cuts the m columns into
 * contiguous shards -- SampleIter's chunks (lmm/lmm.py:427-434), chunk length rounded up to 128 columns -- scanned
 * concurrently by one host thread per device; every shard writes its rows of the caller's output arrays directly, so the
 * result is in input order and bit-identical to a single-GPU scan.  One phenotype per pass (pg_set_design semantics).
 * The one-process-per-GPU route (torch.distributed + NCCL, pygemma_b200/multi.py) is the other way to use several GPUs.
 */
typedef struct pg_multi pg_multi;
/* devices: ngpu CUDA device ordinals, or NULL for 0 .. ngpu-1 */
int pg_multi_create(int n, int c0, int ngpu, const int* devices, pg_multi** out);
int pg_multi_destroy(pg_multi* mh);
const char* pg_multi_last_error(const pg_multi* mh);
int pg_multi_count(const pg_multi* mh);
pg_handle* pg_multi_handle(pg_multi* mh, int i);   /* the i-th device's handle (options, probes); owned by mh */
/* as pg_set_kinship on the first device + peer copies; bcast_ms (nullable): wall time of the copies */
int pg_multi_set_kinship(pg_multi* mh, const double* K_host, double* d_out_host, float* eig_ms, float* bcast_ms);
int pg_multi_set_eigen(pg_multi* mh, const double* U_host, int u_row_major, const double* d_host);
int pg_multi_set_design(pg_multi* mh, const double* W_host, const double* y_host, int already_rotated, float* ms);
/* as pg_scan; timing (nullable) receives pg_multi_count(mh) entries, one per device */
int pg_multi_scan(pg_multi* mh, const void* X, int xdtype, int64_t ld, int layout, int64_t m, int grid,
                  double* beta, double* se_beta, double* tau, double* lambda, double* F_wald, double* p_wald,
                  int32_t* status, int32_t* n_eval2, int32_t* n_eval3, pg_timing* timing);

/*
 * Unit-level probes used by the parity tests (mirrors of the reference's Python-callable cpdefs).
 * pg_probe_precompute: one precompute_mat(lam, ...) (pyx:880) for one rotated genotype vector x (host,
 * n doubles): out9 = {yPy, yPPy, yPPPy, trP, trPP, logdet_H, logdet_WtHinvW, xPx, yPx}.
 * fixed_index >= 0 evaluates the exact table row of lambda = 10^(fixed_index-5) instead of `lam`.
 */
int pg_probe_precompute(pg_handle* h, const double* x_rot_host, double lam, int fixed_index, int full,
                        double* out9);
/* F(1, nu) survival function evaluated on the device for k values (scipy.stats.f.sf, lmm/lmm.py:482) */
int pg_probe_f_sf(pg_handle* h, const double* F_host, double nu, int64_t k, double* p_host);
/* rotated genotypes of the last SNP block pg_scan processed: copies min(count, block) * n doubles
 * (SNP-major) and reports the global index of the block's first SNP in *row0 (nullable) */
int pg_probe_rotated(pg_handle* h, double* xr_host, int64_t count, int64_t* row0);
/*
 * Host-only probe: how many launches the fused rotation of one block of mb SNPs takes at n samples on a device with the
 * given L2 persisting set-aside / access-policy window limits (bytes) and SM count, and (*group_tiles, nullable) the eigen
 * tiles per launch -- 0 and one launch when the block is too short (fewer than 8 waves of cluster tiles per launch) or a
 * group of digit planes does not fit the set-aside (large n).  See INTEGRATION.md, "L2 set-aside".
 */
int pg_probe_rotation_launches(int n, int64_t mb, int64_t setaside_bytes, int64_t max_window_bytes, int sm_count,
                               int32_t* group_tiles);
/* The SNP-block boundaries pg_scan uses for m SNPs when one block holds at most `blk` SNPs (pure host arithmetic, needs
 * no device): host_input != 0 -> the geometric ramp of short first blocks that lets packing / upload of block b+1 hide
 * under the compute of block b; on-device input and a caller-fixed block size (pg_set_options) -> plain blocks.
 * Writes min(count + 1, cap) boundaries (first = 0, last = m) and returns the block count, or PG_ERR_ARG. */
int pg_probe_block_plan(int64_t m, int64_t blk, int host_input, int fixed_block, int64_t* starts, int cap);

#ifdef __cplusplus
}
#endif
#endif /* PYGEMMA_B200_H */
