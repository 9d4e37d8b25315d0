"""GPU parity tests AT THE SIZES THE BENCH NUMBERS ARE QUOTED ON (BASELINE.json configs[2], [3], [4]), through the C ABI.

Every committed small-size test keeps the persistent loop of rotate_i8_tc2_kernel below one wave of clusters; here each
rotation launch has thousands of (512 SNPs x 32 eigenvectors) tiles per launch, >= 13 eigen-tile groups and ragged tails,
so accumulator-phase flips, stage-ring wrap across tiles and the remote tmem_empty arrivals are compared with
  * the cuBLAS int8 split (bit for bit, every SNP),
  * the FP64 GEMM rotation (<= 1e-13 of the column norm, three different column ranges),
  * the CPU oracle (reference algorithm, oracle/reml_oracle.c) on U^T x formed independently in FP64 (1e-6 relative).
torch is used here only to make the large inputs on the device (QR, kinship panel) and for the FP64 check products."""
import os

import numpy as np
import pytest

from conftest import COLS, GOLDEN

pytestmark = pytest.mark.gpu

TOL = 1e-6


def rel(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _capi():
    from pygemma_b200 import _capi

    return _capi


def _check(o, ref, idx=None, tol=TOL, tag=""):
    for c in COLS:
        a, b = np.asarray(o[c]), np.asarray(ref[c])
        if idx is not None:
            a = a[idx]
        assert np.array_equal(np.isnan(a), np.isnan(b)), (tag, c)
        e = rel(a[~np.isnan(b)], b[~np.isnan(b)])
        assert e.size == 0 or e.max() < tol, (tag, c, float(e.max()), int(e.argmax()))


def _genotypes(torch, n, m, seed, dev="cuda:0"):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    X = torch.empty((n, m), dtype=torch.int8, device=dev)
    step = max(256, (1 << 26) // n // 256 * 256)
    for a in range(0, m, step):
        b = min(m, a + step)
        mf = torch.rand(b - a, generator=g, device=dev) * 0.45 + 0.05
        X[:, a:b] = ((torch.rand(n, b - a, generator=g, device=dev) < mf).to(torch.int8)
                     + (torch.rand(n, b - a, generator=g, device=dev) < mf).to(torch.int8))
    return X


def _design(torch, n, c0, d_t, Q, seed, dev="cuda:0"):
    """W = [1, gaussians]; y = U sqrt(d) z (polygenic part with covariance ~ K) + noise + covariate effects."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    W = torch.cat([torch.ones(n, 1, device=dev, dtype=torch.float64),
                   torch.randn(n, c0 - 1, generator=g, device=dev, dtype=torch.float64)], dim=1)
    u = Q @ (d_t.sqrt() * torch.randn(n, generator=g, device=dev, dtype=torch.float64))
    u = u / u.std()
    y = 0.7 * u + 0.7 * torch.randn(n, generator=g, device=dev, dtype=torch.float64) + 0.05 * W[:, 1:].sum(dim=1)
    return W, y


@pytest.fixture(scope="module")
def eig10k():
    """n = 10 000: a random orthogonal U (QR of a gaussian, on the device) with a chi-square-like spectrum -- the
    pg_set_eigen door, no syevd needed -- plus two designs (c0 = 10 and c0 = 40)."""
    import torch

    n = 10000
    dev = "cuda:0"
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    A = torch.randn(n, n, generator=g, device=dev, dtype=torch.float64)
    Q, _ = torch.linalg.qr(A)
    del A
    d_t = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) ** 2 * 4 + 0.02).sort().values
    W10, y10 = _design(torch, n, 10, d_t, Q, 5)
    W40, y40 = _design(torch, n, 40, d_t, Q, 6)
    out = {"n": n, "U": Q.cpu().numpy(), "d": d_t.cpu().numpy(), "W10": W10.cpu().numpy(), "y10": y10.cpu().numpy(),
           "W40": W40.cpu().numpy(), "y40": y40.cpu().numpy()}
    del Q
    torch.cuda.empty_cache()
    return out


def test_c3_n10000_int8_every_tile_of_the_fused_rotation(eig10k):
    """BASELINE configs[2] shape: n = 10 000, c0 = 10, int8 dosages, m = 41 237 SNPs (ragged), default blocking."""
    import torch

    from oracle import oracle

    capi = _capi()
    n, U, d, W, y = eig10k["n"], eig10k["U"], eig10k["d"], eig10k["W10"], eig10k["y10"]
    m = 41237
    X = _genotypes(torch, n, m, 77).cpu().numpy()
    with capi.Handle(n, 10) as h:
        h.set_eigen(U, d)
        h.set_design(W, y)
        res = {}
        for eng in (capi.PG_ROT_I8TC, capi.PG_ROT_I8SPLIT):
            h.set_options(rotation=eng)   # block_snps = 0: the library's own blocking (ramped first blocks, ragged tail)
            o = h.scan(X)
            assert o["timing"]["rot_engine"] == eng and o["timing"]["n_blocks"] >= 4
            assert (o["status"] == 0).all()
            res[eng] = o
        a, b = res[capi.PG_ROT_I8TC], res[capi.PG_ROT_I8SPLIT]
        for c in COLS + ["n_eval2", "n_eval3"]:   # every SNP, every bit
            assert np.array_equal(a[c], b[c], equal_nan=True), c
        # rotated vectors against the FP64 GEMM on three column ranges (head, middle, ragged tail)
        for lo, hi in ((0, 5120), (20000, 25120), (m - 5003, m)):
            rows = {}
            for eng in (capi.PG_ROT_I8TC, capi.PG_ROT_FP64):
                h.set_options(rotation=eng)
                o = h.scan(X[:, lo:hi])
                assert o["timing"]["rot_engine"] == eng
                xr, row0 = h.probe_rotated(512)
                rows[eng] = (xr, row0, o)
            (x8, r8, o8), (x64, r64, o64) = rows[capi.PG_ROT_I8TC], rows[capi.PG_ROT_FP64]
            assert r8 == r64
            cols = X[:, lo + r8: lo + r8 + x8.shape[0]].astype(np.float64)
            scale = np.sqrt((cols ** 2).sum(0))[:, None]
            assert (np.abs(x8 - x64) / scale).max() < 1e-13, (lo, hi)
            for c in COLS:   # lambda is only determined to the optimiser's stopping rule where the likelihood is flat
                assert rel(o8[c], o64[c]).max() < (1e-6 if c == "lambda" else 1e-8), (lo, hi, c)
                assert np.array_equal(o8[c], a[c][lo:hi], equal_nan=True), (lo, hi, c)   # position-independent bits
    # 40 SNPs spread over all blocks and tiles against the oracle, rotation done independently in FP64 on the host
    idx = np.unique(np.linspace(0, m - 1, 40).astype(np.int64))
    xr = U.T @ X[:, idx].astype(np.float64)
    ref = oracle.scan_rotated(d, U.T @ y, U.T @ W, np.ascontiguousarray(xr.T))
    _check(a, ref, idx=idx, tag="c3 n=10000")


def test_c3_n10000_through_kinship_syevd_and_factorize():
    """The same size through K -> cuSOLVER syevd (lmm.factorize), the public lmm.pygemma call with the factor in the K
    position, a second call with another covariate count (handle re-created by pg_copy_eigen), vs the oracle."""
    import torch

    from oracle import oracle
    from pygemma_b200 import lmm

    n, m, c0 = 10000, 8192 + 77, 10
    dev = "cuda:0"
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    mk = 2 * n
    maf = torch.rand(mk, generator=g, device=dev) * 0.45 + 0.05
    G = ((torch.rand(n, mk, generator=g, device=dev) < maf).double() + (torch.rand(n, mk, generator=g, device=dev) < maf).double())
    sd = G.std(dim=0)
    sd[sd == 0] = 1.0
    G = (G - G.mean(dim=0)) / sd
    K = G @ G.T / mk
    K.diagonal().add_(1e-3)
    u = G @ torch.randn(mk, generator=g, device=dev, dtype=torch.float64) / mk ** 0.5
    del G
    u = u / u.std()
    W = torch.cat([torch.ones(n, 1, device=dev, dtype=torch.float64),
                   torch.randn(n, c0 - 1, generator=g, device=dev, dtype=torch.float64)], dim=1)
    y = 0.7 * u + 0.7 * torch.randn(n, generator=g, device=dev, dtype=torch.float64) + 0.05 * W[:, 1:].sum(dim=1)
    Xd = _genotypes(torch, n, m, 3)
    X, Wh, yh = Xd.cpu().numpy(), W.cpu().numpy(), y.cpu().numpy()
    K_host = K.cpu().numpy()
    del K
    idx = np.unique(np.linspace(0, m - 1, 32).astype(np.int64))
    with lmm.factorize(K_host, c0=c0) as F:
        assert F.eig_ms > 0
        df = lmm.pygemma(yh, X, Wh, F, snps=np.arange(m))
        assert list(df.columns) == COLS + ["SNPs"] and len(df) == m
        h = F.handle(c0)
        U_t = torch.empty(n * n, dtype=torch.float64, device=dev)
        d_t = torch.empty(n, dtype=torch.float64, device=dev)
        h.get_eigen_device(U_t.data_ptr(), d_t.data_ptr())
        Ut = U_t.view(n, n)   # column-major U read row-major = U^T
        assert (torch.diff(d_t) >= 0).all() and d_t.min() >= 0
        xr = (Ut @ Xd[:, torch.from_numpy(idx).to(dev)].double()).T.contiguous().cpu().numpy()
        ref = oracle.scan_rotated(d_t.cpu().numpy(), (Ut @ y).cpu().numpy(), (Ut @ W).cpu().numpy(), xr)
        _check({c: df[c].to_numpy() for c in COLS}, ref, idx=idx, tag="syevd n=10000")
        # fewer covariates: the factor re-creates its handle and copies U, d device to device
        df4 = lmm.pygemma(yh, X[:, :2048], Wh[:, :4], F)
        assert F.handle(4) is not h and F.handle(4).c0 == 4
        i4 = idx[idx < 2048][:8]
        xr4 = (Ut @ Xd[:, torch.from_numpy(i4).to(dev)].double()).T.contiguous().cpu().numpy()
        ref4 = oracle.scan_rotated(d_t.cpu().numpy(), (Ut @ y).cpu().numpy(), (Ut @ W[:, :4]).cpu().numpy(), xr4)
        _check({c: df4[c].to_numpy() for c in COLS}, ref4, idx=i4, tag="copy_eigen c0=4")
        del U_t, Ut
    # the one-shot call (eigh inside, like the reference) gives the same frame
    df1 = lmm.pygemma(yh, X[:, :1024], Wh, K_host)
    for c in COLS:
        assert rel(df1[c].to_numpy(), df[c].to_numpy()[:1024]).max() < 1e-7, c


def test_c5_n10000_c40_grid_vs_oracle(eig10k):
    """BASELINE configs[4]: grid-search lambda with c0 = 40 covariates at n = 10 000."""
    import torch

    from oracle import oracle

    capi = _capi()
    n, U, d, W, y = eig10k["n"], eig10k["U"], eig10k["d"], eig10k["W40"], eig10k["y40"]
    m = 3000 + 13
    X = _genotypes(torch, n, m, 41).cpu().numpy()
    with capi.Handle(n, 40) as h:
        h.set_eigen(U, d)
        h.set_design(W, y)
        og = h.scan(X, grid=True)
        ob = h.scan(X[:, :512])   # Brent + Newton at c0 = 40 as well
        # the same scans with the moments fused into the rotation (pg_set_moment_fusion; the automatic mode takes it at this
        # covariate count when the spectrum has <= 140 nodes): thousands of eigen + G tiles per launch, every SNP compared
        unfused = og["timing"]["rot_engine"] == capi.PG_ROT_I8TC
        h.set_moment_fusion(1)
        ogf = h.scan(X, grid=True)
        obf = h.scan(X[:, :512])
        assert ogf["timing"]["rot_engine"] == capi.PG_ROT_I8TC_MOMENTS and h.fusion_info()["g_columns"] >= 41 * 100
        if unfused:
            assert np.array_equal(ogf["lambda"], og["lambda"])
            for c in COLS:
                assert rel(ogf[c], og[c]).max() < 1e-8, ("fused grid", c)
                assert rel(obf[c], ob[c]).max() < (1e-6 if c == "lambda" else 1e-8), ("fused brent", c)
    assert (og["status"] == 0).all()
    assert set(np.unique(og["lambda"])) <= set(10.0 ** np.arange(-5, 6))
    idx = np.unique(np.linspace(0, m - 1, 24).astype(np.int64))
    yr, wr = U.T @ y, U.T @ W
    xr = np.ascontiguousarray((U.T @ X[:, idx].astype(np.float64)).T)
    refg = oracle.scan_rotated(d, yr, wr, xr, grid=True)
    assert np.array_equal(og["lambda"][idx], refg["lambda"])
    _check(og, refg, idx=idx, tag="c5 grid")
    _check(ogf, refg, idx=idx, tag="c5 grid, moments fused")
    ib = idx[idx < 512]
    refb = oracle.scan_rotated(d, yr, wr, np.ascontiguousarray((U.T @ X[:, ib].astype(np.float64)).T))
    _check(ob, refb, idx=ib, tag="c5 brent")
    _check(obf, refb, idx=ib, tag="c5 brent, moments fused")


def test_c4_n50000_shape_fused_vs_cublas_and_oracle():
    """BASELINE configs[3] shape: n = 50 000 (U = 20 GB fp64, 17.5 GB of digit planes, 1 563 eigen tiles), device-resident
    inputs through pg_set_eigen_device / pg_scan_device: fused tcgen05 rotation bit-identical to the cuBLAS split over
    ~2 x 74 x 512 SNPs (ragged), 8 SNPs against the oracle."""
    import torch

    from oracle import oracle

    capi = _capi()
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs a 180 GB B200")
    n, c0 = 50000, 10
    m = 2 * 74 * 512 + 333
    dev = "cuda:0"
    g = torch.Generator(device=dev)
    g.manual_seed(50)
    A = torch.randn(n, n, generator=g, device=dev, dtype=torch.float64)
    Q, _ = torch.linalg.qr(A)
    del A
    torch.cuda.empty_cache()
    d_t = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 3 + 0.05).sort().values
    W, y = _design(torch, n, c0, d_t, Q, 8)
    Ut = Q.t().contiguous()   # row i = eigenvector i: as a flat buffer this is column-major U
    del Q
    torch.cuda.empty_cache()
    X = _genotypes(torch, n, m, 9)
    out = {e: torch.empty((6, m), dtype=torch.float64, device=dev) for e in (capi.PG_ROT_I8TC, capi.PG_ROT_I8SPLIT)}
    st = {e: torch.zeros((3, m), dtype=torch.int32, device=dev) for e in out}
    with capi.Handle(n, c0) as h:
        h.set_eigen_device(Ut.data_ptr(), False, d_t.data_ptr())
        h.set_design(W.cpu().numpy(), y.cpu().numpy())
        for eng in out:
            h.set_options(rotation=eng)
            tm = h.scan_device(X.data_ptr(), capi.PG_X_I8, m, capi.PG_X_SAMPLE_MAJOR, m, False,
                               [out[eng][i].data_ptr() for i in range(6)], st[eng][0].data_ptr(), st[eng][1].data_ptr(),
                               st[eng][2].data_ptr())
            assert tm["rot_engine"] == eng
    a, b = out[capi.PG_ROT_I8TC], out[capi.PG_ROT_I8SPLIT]
    assert torch.equal(st[capi.PG_ROT_I8TC], st[capi.PG_ROT_I8SPLIT]) and int((st[capi.PG_ROT_I8TC][0] != 0).sum()) == 0
    assert torch.equal(a.view(torch.int64), b.view(torch.int64))   # bits, NaN patterns included
    idx = np.unique(np.linspace(0, m - 1, 8).astype(np.int64))
    xr = (Ut @ X[:, torch.from_numpy(idx).to(dev)].double()).T.contiguous().cpu().numpy()
    ref = oracle.scan_rotated(d_t.cpu().numpy(), (Ut @ y).cpu().numpy(), (Ut @ W).cpu().numpy(), xr)
    got = a.cpu().numpy()
    _check({c: got[i] for i, c in enumerate(COLS)}, ref, idx=idx, tag="c4 n=50000")


def test_multi_phenotype_traits_match_the_oracle():
    """pg_set_design_multi / lmm.pygemma_multi: every trait of one multi-trait pass against the oracle run on that trait
    alone (the bit-identity with single-trait CUDA scans is test_multi_phenotype_scan_equals_one_scan_per_trait)."""
    from oracle import oracle
    from pygemma_b200 import lmm
    from pygemma_b200.synth import make_problem

    n, m, c0, q = 900, 400, 4, 3
    p = make_problem(n, m, c0, seed=21, m_k=2000)
    rng = np.random.default_rng(5)
    Y = np.stack([p["Y"].reshape(-1), rng.standard_normal(n),
                  p["Y"].reshape(-1) * 0.3 + p["X"][:, 7] * 0.2 + rng.standard_normal(n)], axis=1)
    for grid in (False, True):
        frames = lmm.pygemma_multi(Y, p["X"], p["W"], p["K"], grid=grid)
        assert len(frames) == q
        for ph in range(q):
            ref = oracle.pygemma(Y[:, ph], p["X"], p["W"], p["K"], grid=grid)
            got = {c: frames[ph][c].to_numpy() for c in COLS}
            if ph == 1 and not grid:
                # trait 1 is pure noise: the restricted likelihood is flat, lambda-hat sits wherever the reference's Newton
                # iteration happens to stop (relative step < 1e-5, pyx:1411) and a last-bit change of the tables can cost or
                # save one iteration; beta, se, tau, F, p do not feel it (fuzz notes)
                assert rel(got["lambda"], ref["lambda"]).max() < 2e-5
                got = dict(got, **{"lambda": ref["lambda"]})
            _check(got, ref, tag=("multi", grid, ph))


def test_wide_spectrum_needs_more_nodes_than_one_shared_memory_slab():
    """Eigenvalues spread log-uniformly over 22 decades: the compression plan keeps > 1 210 nodes, more than the fixed-lambda
    contraction can hold in shared memory at once (it walks H in slabs); results must still match the oracle."""
    from oracle import oracle

    capi = _capi()
    rng = np.random.default_rng(17)
    n, m, c0 = 3000, 160, 5
    d = np.sort(10.0 ** rng.uniform(-14, 8, n))
    W = np.c_[np.ones(n), rng.standard_normal((n, c0 - 1))]
    X = rng.standard_normal((n, m))
    y = rng.standard_normal(n) * np.sqrt(np.minimum(0.5 * d, 50.0) + 1.0) + 0.1 * X[:, 0]
    ref = oracle.scan_rotated(d, y, W, np.ascontiguousarray(X.T))
    refg = oracle.scan_rotated(d, y, W, np.ascontiguousarray(X.T), grid=True)
    with capi.Handle(n, c0) as h:
        h.set_eigen(None, d)
        h.set_design(W, y, already_rotated=True)
        o = h.scan(X)
        og = h.scan(X, grid=True)
    assert o["timing"]["n_nodes"] > 1210, o["timing"]["n_nodes"]
    _check(o, ref, tag="wide spectrum")
    _check(og, refg, tag="wide spectrum grid")


def test_literal_float32_reference_is_tracked():
    """The reference as its users run it (float32 storage, float32 eigh and sgemm: 'ref32' goldens, made by
    tests/golden/make_golden.py from the unmodified sources) is noisy at 1e-5..1e-3; the float64 product
    must stay inside that noise: loose bounds per column and the same ranking of the p-values."""
    import glob

    from scipy import stats

    from pygemma_b200 import lmm

    bounds = {"beta": 5e-3, "se_beta": 1e-4, "tau": 1e-4, "lambda": 1e-3, "F_wald": 1e-2, "p_wald": 2e-3}
    for path in sorted(glob.glob(os.path.join(GOLDEN, "e2e_*.npz"))):
        g = np.load(path)
        df = lmm.pygemma(g["Y"], g["X"], g["W"], g["K"], grid=bool(g["grid"]))
        assert all(df[c].dtype == np.float64 for c in COLS)   # documented deviation: the reference returns float32 columns
        for c in COLS:
            e = rel(df[c].to_numpy(), g[f"r32_{c}"])
            assert e.max() < bounds[c], (path, c, float(e.max()))
        assert stats.spearmanr(df["p_wald"].to_numpy(), g["r32_p_wald"]).statistic > 0.9999


def _oracle_de(d, Yr, Wr, Xr):
    """calculate_de (reference lmm/lmm.py:498-532) on rotated float64 inputs via the oracle: per column g the phenotype is
    X[:, g] and the tested regressor Y -- calc_lambda_restricted(d, X[:, g], [W, Y]) then
    calc_beta_vg_ve_restricted_overload(d, W, Y, lambda, X[:, g])."""
    from oracle import oracle

    out = {c: [] for c in COLS}
    yrow = np.ascontiguousarray(np.asarray(Yr, float).reshape(1, -1))
    for g in range(Xr.shape[1]):
        r = oracle.scan_rotated(d, Xr[:, g], Wr, yrow)
        for c in COLS:
            out[c].append(r[c][0])
    return {c: np.array(v) for c, v in out.items()}


def test_de_mode_is_the_role_swapped_scan():
    """de=True: each column of X is an outcome (real-valued, FP64 rotation), Y the predictor."""
    from oracle import oracle
    from pygemma_b200 import lmm
    from pygemma_b200.synth import make_problem

    n, m, c0 = 700, 96, 4
    p = make_problem(n, m, c0, seed=44, m_k=1500)
    rng = np.random.default_rng(3)
    pred = p["X"][:, 5].astype(np.float64)                      # the predictor: one SNP's dosages
    E = rng.standard_normal((n, m)) + 0.15 * pred[:, None] * (np.arange(m) % 3 == 0) + p["Y"] * 0.3   # "expression" columns
    for c0_use in (c0, 0):
        W = p["W"][:, :c0_use]
        df = lmm.pygemma(pred, E, W, p["K"], de=True, snps=np.arange(m))
        assert list(df.columns) == COLS + ["SNPs"] and len(df) == m
        d, U, yr, xr, wr = oracle.eigen_rotate(p["K"], pred, E, W if c0_use else np.zeros((n, 0)))
        ref = _oracle_de(d, yr.reshape(-1), wr.reshape(n, c0_use), xr)
        _check({c: df[c].to_numpy() for c in COLS}, ref, tag=("de", c0_use))
    # the golden vectors made by the reference's own cpdefs called with swapped roles (tests/golden/make_golden.py)
    path = os.path.join(GOLDEN, "de_small.npz")
    if os.path.exists(path):
        g = np.load(path)
        df = lmm.pygemma(g["yr"], g["xr"], g["wr"], g["d"], eigen=False, de=True)
        _check({c: df[c].to_numpy() for c in COLS}, {c: g[f"r64_{c}"] for c in COLS}, tag="de golden")


@pytest.mark.parametrize("name", ["lrt_interior", "lrt_low_h2", "lrt_c8"])
def test_lrt_outputs_match_reference_functions(name):
    """lrt=True: null model and per-SNP ML fits on the device (null_model_kernel, MlSolver in reml_solve_kernel) against
    the reference's live lmm.calc_lambda / likelihood_lambda (tests/golden/lrt_*.npz) and the dense oracle; the Wald
    columns must not change a bit when the LRT columns are requested."""
    from oracle import oracle
    from pygemma_b200 import lmm

    capi = _capi()
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    d, yr, wr, xr = g["d"], g["yr"], g["wr"], g["xr"]
    n, m = xr.shape
    with capi.Handle(n, wr.shape[1]) as h:
        h.set_eigen(None, d)
        h.set_design(wr, yr, already_rotated=True)
        h.set_options(block_snps=16)   # several blocks
        o = h.scan(np.ascontiguousarray(xr), lrt=True)
        o0 = h.scan(np.ascontiguousarray(xr))
        nm = h.null_model(0)
    for c in COLS:
        assert np.array_equal(o[c], o0[c], equal_nan=True), c
    assert (o["status"] == 0).all()
    assert rel(nm["lambda_null"], g["r64_lambda_null"]) < TOL and rel(nm["l_null"], g["r64_l_null"]) < 1e-9
    assert rel(o["lambda_ml"], g["r64_lambda_alt"]).max() < TOL
    assert rel(o["loglik_ml"], g["r64_l_alt"]).max() < 1e-9
    assert np.abs(o["D_lrt"] - g["r64_D_lrt"]).max() < 1e-7
    ok = g["r64_p_lrt"] > 1e-9
    assert rel(o["p_lrt"][ok], g["r64_p_lrt"][ok]).max() < TOL
    ref = oracle.lrt_rotated(d, yr, wr, np.ascontiguousarray(xr.T))
    assert rel(nm["tau_null"], ref["tau_null"]) < 1e-9
    assert rel(o["p_lrt"], ref["p_lrt"]).max() < TOL   # accurate tail on both sides
    # the public call: three more columns, named as the reference's commented-out keys
    df = lmm.pygemma(yr, xr, wr, d, eigen=False, lrt=True, snps=np.arange(m))
    assert list(df.columns) == COLS + ["D_lrt", "p_lrt", "likelihood", "SNPs"]
    assert np.array_equal(df["D_lrt"].to_numpy(), o["D_lrt"]) and np.array_equal(df["likelihood"].to_numpy(), o["loglik_ml"])
    assert rel(lmm.last_null_model[0]["l_null"], g["r64_l_null"]) < 1e-9
    assert rel(lmm.null_model(yr, wr, d, eigen=False)["lambda_null"], g["r64_lambda_null"]) < TOL


def test_lrt_through_kinship_and_multi_trait():
    """lrt=True end to end (syevd, int8 rotation) vs the dense oracle on LAPACK-rotated inputs, single- and multi-trait."""
    from oracle import oracle
    from pygemma_b200 import lmm
    from pygemma_b200.synth import make_problem

    n, m, c0 = 500, 120, 3
    p = make_problem(n, m, c0, seed=71, m_k=1200)
    rng = np.random.default_rng(2)
    Y = np.stack([p["Y"].reshape(-1), 0.4 * p["Y"].reshape(-1) + rng.standard_normal(n)], axis=1)
    frames = lmm.pygemma_multi(Y, p["X"], p["W"], p["K"], lrt=True)
    nulls = list(lmm.last_null_model)   # one per trait; a later single-trait call replaces the list
    d, U, _, xr, wr = oracle.eigen_rotate(p["K"], Y[:, 0], p["X"], p["W"])
    for ph in range(2):
        ref = oracle.lrt_rotated(d, U.T @ Y[:, ph], wr, np.ascontiguousarray(xr.T))
        df = frames[ph]
        assert np.abs(df["D_lrt"].to_numpy() - ref["D_lrt"]).max() < 1e-6
        assert rel(df["likelihood"].to_numpy(), ref["loglik_ml"]).max() < 1e-8
        assert rel(df["p_lrt"].to_numpy(), ref["p_lrt"]).max() < 1e-5
        assert rel(nulls[ph]["l_null"], ref["l_null"]) < 1e-9
        one = lmm.pygemma(Y[:, ph], p["X"], p["W"], p["K"], lrt=True)
        for c in COLS + ["D_lrt", "p_lrt", "likelihood"]:
            assert np.array_equal(one[c].to_numpy(), df[c].to_numpy(), equal_nan=True), (ph, c)


def test_traw_ingest_with_std_filter_matches_oracle(tmp_path):
    """PLINK .traw -> packed 2-bit block -> device decode / mean imputation / int8 rotation, with the callers' std > 0
    filter (experiments/wtccc/run_pygemma.py:407-410): against the oracle on the NumPy-imputed kept columns."""
    from oracle import oracle
    from pygemma_b200 import traw
    from pygemma_b200.synth import make_problem

    n, m, c0 = 401, 90, 3
    p = make_problem(n, m, c0, seed=13, m_k=900)
    rng = np.random.default_rng(6)
    G = p["X"].astype(np.float64)
    G[rng.random(G.shape) < 0.02] = np.nan
    G[:, 7] = 0.0       # monomorphic -> filtered
    G[:, 8] = np.nan    # all missing -> filtered
    G[:, 30] = p["X"][:, 30]   # complete column
    path = str(tmp_path / "toy.traw.gz")
    traw.write_traw(path, G)
    for policy in ("omit", "propagate"):
        for std in (False, True):
            df, keep = traw.pygemma_traw(p["Y"], path, p["W"], p["K"], filter_std=True, nan_policy=policy, standardize=std)
            with np.errstate(invalid="ignore"):
                want = (np.nan_to_num(np.nanstd(np.where(np.isnan(G).all(0), 0.0, G), axis=0)) > 0) if policy == "omit" \
                    else (np.nan_to_num(G.std(axis=0)) > 0)
            assert np.array_equal(keep, want) and not keep[7] and not keep[8] and keep[30]
            assert len(df) == int(keep.sum()) and list(df["SNPs"]) == [f"rs{i}" for i in np.where(keep)[0]]
            Xk = G[:, keep]
            Xi = np.where(np.isnan(Xk), np.nanmean(Xk, axis=0), Xk)   # SimpleImputer(strategy='mean')
            if std:
                Xi = (Xi - Xi.mean(axis=0)) / Xi.std(axis=0)
            ref = oracle.pygemma(p["Y"], Xi, p["W"], p["K"])
            _check({c: df[c].to_numpy() for c in COLS}, ref, tag=("traw", policy, std))
    df_all, keep_all = traw.pygemma_traw(p["Y"], path, p["W"], p["K"])
    assert keep_all.all() and len(df_all) == m


def test_four_cta_cluster_rotation_is_bit_identical():
    """PG_TC_CLUSTER=4 (rotate_i8_tc4.cuh: two CTA pairs per cluster, genotype tiles TMA-multicast to both pairs; kept as a
    measured alternative, not the default) against the default CTA-pair kernel: exact integer arithmetic, so every output
    bit must agree -- odd eigen-tile count (phantom tile of the second pair), ragged SNP tail, several waves of units."""
    import subprocess
    import sys
    import tempfile

    code = ("import numpy as np, sys; sys.path.insert(0, '.');"
            "from pygemma_b200 import _capi; from pygemma_b200.synth import make_problem;"
            "n, m = 2080, 9000 + 37;"   # 65 eigen tiles (odd), 18 SNP tiles -> 594 units over <= 37 clusters
            "p = make_problem(n, 64, 3, seed=12, m_k=2 * n);"
            "X = np.random.default_rng(4).integers(0, 3, size=(n, m), dtype=np.int8);"
            "h = _capi.Handle(n, 3); h.set_kinship(p['K']); h.set_design(p['W'], p['Y']);"
            "h.set_options(rotation=_capi.PG_ROT_I8TC, block_snps=0);"
            "o = h.scan(X); xr, _ = h.probe_rotated(64);"
            "np.save(sys.argv[1], np.concatenate([np.stack([o[c] for c in ['beta','se_beta','tau','lambda','F_wald','p_wald']]).ravel(), xr.ravel()]))")
    outs = []
    for cl in ("2", "4"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            env = dict(os.environ, PG_TC_CLUSTER=cl)
            subprocess.check_call([sys.executable, "-c", code, f.name], env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
            outs.append(np.load(f.name))
    assert np.array_equal(outs[0], outs[1], equal_nan=True)


def test_tables_by_gemm_agree_with_row_reductions():
    """The SNP-independent lambda tables come from one DGEMM per design (powers of 1/(lambda d + 1) kept per eigen-system);
    PG_TABLES_GEMM=0 selects the per-row reduction kernel of round 1.  Both must give the same scan to summation rounding --
    single- and multi-trait, a second design on the same handle, c0 = 0."""
    import subprocess
    import sys
    import tempfile

    code = ("import numpy as np, sys; sys.path.insert(0, '.');"
            "from pygemma_b200 import _capi; from pygemma_b200.synth import make_problem;"
            "p = make_problem(1500, 400, 7, seed=5, m_k=3000); rng = np.random.default_rng(1);"
            "Y2 = np.stack([p['Y'].reshape(-1), rng.standard_normal(1500)], axis=1);"
            "h = _capi.Handle(1500, 7); h.set_kinship(p['K']); outs = [];"
            "h.set_design(p['W'], p['Y']); o = h.scan(p['X']); outs.append(np.stack([o[c] for c in ['beta','se_beta','tau','lambda','F_wald','p_wald']]));"
            "h.set_design(p['W'], Y2); o = h.scan(p['X'], grid=True); outs.append(np.stack([o[c][1] for c in ['beta','se_beta','tau','lambda','F_wald','p_wald']]));"
            "h.close(); h = _capi.Handle(1500, 0); h.set_kinship(p['K']); h.set_design(np.zeros((1500, 0)), p['Y']); o = h.scan(p['X']);"
            "outs.append(np.stack([o[c] for c in ['beta','se_beta','tau','lambda','F_wald','p_wald']]));"
            "np.save(sys.argv[1], np.stack(outs))")
    res = []
    for g in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            env = dict(os.environ, PG_TABLES_GEMM=g)
            subprocess.check_call([sys.executable, "-c", code, f.name], env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
            res.append(np.load(f.name))
    assert np.array_equal(np.isnan(res[0]), np.isnan(res[1]))
    e = rel(res[0], res[1])
    assert np.nanmax(e[:, [0, 1, 2, 4, 5]]) < 1e-9 and np.nanmax(e[:, 3]) < 1e-7, (float(np.nanmax(e)),)
