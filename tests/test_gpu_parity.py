"""GPU parity tests: the CUDA path, called through the C ABI (include/pygemma_b200.h), against
  * the golden vectors produced by the reference's own compiled code (ref64, tests/golden/), and
  * the CPU oracle (oracle/reml_oracle.c) on seeded inputs.
Tolerance (BASELINE.json north_star): lambda, beta, se, p within 1e-6 relative, FP64 throughout;
row order and row count identical (bit-exact: row i <-> genotype column i)."""
import glob
import os

import numpy as np
import pytest

from conftest import COLS, GOLDEN

pytestmark = pytest.mark.gpu

TOL = 1e-6  # north star: 1e-6 relative on lambda, beta, se, p (tau and F are held to the same bar)


def rel(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _capi():
    from pygemma_b200 import _capi

    return _capi


def _check(o, ref, cols=COLS, tol=TOL, idx=None, tag=""):
    for c in cols:
        a, b = np.asarray(o[c]), np.asarray(ref[c])
        if idx is not None:
            a, b = a[idx], b[idx]
        assert np.array_equal(np.isnan(a), np.isnan(b)), (tag, c)
        e = rel(a[~np.isnan(b)], b[~np.isnan(b)])
        assert e.size == 0 or e.max() < tol, (tag, c, float(e.max()), int(e.argmax()))


SCANS = sorted(glob.glob(os.path.join(GOLDEN, "scan_*.npz")))


@pytest.mark.parametrize("path", SCANS, ids=[os.path.basename(p)[5:-4] for p in SCANS])
def test_golden_scan_rotated_inputs(path):
    """eigen=False entry: rotated inputs straight into the REML kernel, vs the reference's lmm.calculate."""
    capi = _capi()
    g = np.load(path)
    d, yr, wr, xr = g["d"], g["yr"], g["wr"], g["xr"]
    n, m = xr.shape
    with capi.Handle(n, wr.shape[1]) as h:
        h.set_eigen(None, d)
        h.set_design(wr, yr, already_rotated=True)
        o = h.scan(np.ascontiguousarray(xr), grid=bool(g["grid"]))
        o_snp = h.scan(np.ascontiguousarray(xr.T), grid=bool(g["grid"]), layout=capi.PG_X_SNP_MAJOR)
    assert o["beta"].shape[0] == m and (o["status"] == 0).all()
    ref = {c: g[f"r64_{c}"] for c in COLS}
    idx = np.array([0, 2, 4, 5, 6, 7]) if "degenerate" in path else None
    _check(o, ref, idx=idx, tag=path)
    for c in COLS:  # layout must not change a single bit
        assert np.array_equal(o[c], o_snp[c], equal_nan=True), c


def test_golden_end_to_end_with_syevd():
    """Whole path incl. cuSOLVER syevd and the rotation, vs the reference's lmm.pygemma (ref64)."""
    capi = _capi()
    for path in sorted(glob.glob(os.path.join(GOLDEN, "e2e_*.npz"))):
        g = np.load(path)
        n, m = g["X"].shape
        ref = {c: g[f"r64_{c}"] for c in COLS}
        for X in (g["X"], g["X"].astype(np.float32), g["X"].astype(np.float64)):
            with capi.Handle(n, g["W"].shape[1]) as h:
                dvals, _ = h.set_kinship(g["K"])
                assert (dvals >= 0).all() and np.all(np.diff(dvals) >= 0)
                h.set_design(g["W"], g["Y"])
                o = h.scan(np.ascontiguousarray(X), grid=bool(g["grid"]))
            _check(o, ref, tag=(path, str(X.dtype)))


def test_pygemma_dataframe_contract():
    """lmm.pygemma: same columns, order, row count, SNP labels as the reference (lmm/lmm.py:403-411)."""
    import pandas as pd

    from pygemma_b200 import lmm

    g = np.load(os.path.join(GOLDEN, "e2e_small.npz"))
    df = lmm.pygemma(g["Y"], g["X"], g["W"], g["K"], snps=g["snps"], verbose=0)
    assert list(df.columns) == list(g["columns"])
    assert isinstance(df.index, pd.RangeIndex) and len(df) == g["X"].shape[1]
    assert list(df["SNPs"]) == list(g["snps"])
    _check({c: df[c].values for c in COLS}, {c: g[f"r64_{c}"] for c in COLS}, tag="df")
    df2 = lmm.pygemma(Y=g["Y"].reshape(-1), X=g["X"].astype(np.int64), W=g["W"], K=g["K"])  # keywords, 1-D Y, int64 X
    assert list(df2.columns) == COLS
    for c in COLS:
        assert np.array_equal(df2[c].values, df[c].values), c
    # eigen=False: caller supplies eigenvalues and rotated inputs (lmm/lmm.py:164-167)
    s = np.load(os.path.join(GOLDEN, "scan_interior.npz"))
    df3 = lmm.pygemma(s["yr"], s["xr"], s["wr"], s["d"], eigen=False)
    _check({c: df3[c].values for c in COLS}, {c: s[f"r64_{c}"] for c in COLS}, tag="eigen=False")
    # a pandas Series with a shifted index aligns like in the reference
    lab = pd.Series(g["snps"], index=np.arange(1, len(g["snps"]) + 1))
    df4 = lmm.pygemma(g["Y"], g["X"], g["W"], g["K"], snps=lab)
    assert pd.isna(df4["SNPs"].iloc[0]) and df4["SNPs"].iloc[1] == g["snps"][0]


def test_precompute_kat_on_device():
    """The reference test-suite's seeded matrices and lambda probes (tests/test_pygemma.py:195-253)."""
    capi = _capi()
    g = np.load(os.path.join(GOLDEN, "kat_precompute.npz"))
    d, wr, xr, yr = (g[k].astype(np.float64) for k in ("d", "Wr", "xr", "Yr"))
    n, c0 = wr.shape
    cf = c0 + 1
    with capi.Handle(n, c0) as h:
        h.set_eigen(None, d)
        h.set_design(wr, yr, already_rotated=True)
        for li, lam in enumerate(g["lams"]):
            s = h.probe_precompute(xr, lam, full=True)
            want = [g[f"r64_full_{k}_{li}"][cf] for k in ("yt_Pi_y", "yt_Pi_Pi_y", "yt_Pi_Pi_Pi_y", "tr_Pi", "tr_Pi_Pi")]
            assert rel(s[:5], want).max() < 1e-7, (lam, rel(s[:5], want))
            sc = g[f"r64_scal_{li}"]
            assert rel(s[5], sc[0]) < 1e-10 and rel(s[6], sc[1]) < 1e-9
            assert rel(s[7], sc[4]) < 1e-8 and rel(s[8], sc[5]) < 1e-6
            s2 = h.probe_precompute(xr, lam, full=False)
            assert rel(s2[[0, 1, 3]], s[[0, 1, 3]]).max() < 1e-12
        # fixed-table row == interpolated evaluation at the same lambda
        a = h.probe_precompute(xr, 1e3, fixed_index=8, full=True)
        b = h.probe_precompute(xr, 1e3, fixed_index=-1, full=True)
        assert rel(a, b).max() < 1e-9
        o = h.scan(xr.reshape(-1, 1))
        og = h.scan(xr.reshape(-1, 1), grid=True)
    assert o["lambda"][0] == g["r64_calc_lambda"][0] and og["lambda"][0] == g["r64_calc_lambda"][1]


def test_f_tail_on_device():
    from scipy import stats

    capi = _capi()
    rng = np.random.default_rng(2)
    with capi.Handle(64, 1) as h:
        for nu in (5, 443, 9989, 49989):
            F = np.concatenate([10.0 ** rng.uniform(-12, 3.3, 500), [0.0, 1.0, 1400.0]])
            ref = stats.f.sf(F, 1, nu)
            got = h.probe_f_sf(F, nu)
            m = ref > 1e-300
            assert rel(got[m], ref[m]).max() < 1e-9, nu


@pytest.mark.parametrize("n,m,c0,grid,h2", [(700, 300, 5, False, 0.5), (1940, 256, 6, False, 0.3),
                                              (449, 400, 6, True, 0.5), (1000, 200, 16, False, 0.6),
                                              (900, 128, 40, True, 0.5), (333, 257, 0, False, 0.4)])
def test_seeded_vs_oracle(n, m, c0, grid, h2):
    """Seeded synthetic problems (mouse / GD449 shapes among them) through K -> syevd -> rotation -> REML,
    against the CPU oracle fed with the device's own eigendecomposition outputs."""
    from oracle import oracle
    from pygemma_b200.synth import make_problem

    capi = _capi()
    p = make_problem(n, m, c0, seed=n + m + c0, h2=h2, m_k=max(200, n // 3))
    with capi.Handle(n, c0) as h:
        h.set_options(block_snps=96)  # several SNP blocks, ragged tail
        h.set_kinship(p["K"])
        h.set_design(p["W"], p["Y"])
        o = h.scan(p["X"], grid=grid)
        assert o["timing"]["n_blocks"] == -(-m // 96)
    ref = oracle.pygemma(p["Y"], p["X"], p["W"], p["K"], grid=grid)
    _check(o, ref, tag=(n, m, c0, grid))
    assert (o["status"] == 0).all()


def test_every_covariate_count_launches_and_matches():
    """c0 = 0..62 (the ABI's maximum) one by one: every x-row width walks through the solver's template and launch-shape switches (warp or
    half-warp tiles, one or two x rows per lane, slab staging, the shared-memory opt-in).  The round-2 fuzz sweep found
    c0 = 13 failing to LAUNCH (static + dynamic shared memory just over the 48 KB default limit) while 12 and 14 ran."""
    from oracle import oracle
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m = 160, 40
    for c0 in range(0, 63):
        p = make_problem(n, m, c0, seed=900 + c0, h2=0.4, m_k=300)
        with capi.Handle(n, c0) as h:
            h.set_kinship(p["K"])
            h.set_design(p["W"], p["Y"])
            for grid in (False, True):
                o = h.scan(p["X"], grid=grid, lrt=not grid)
                ref = oracle.pygemma(p["Y"], p["X"], p["W"], p["K"], grid=grid)
                _check(o, ref, tag=(c0, grid))
                assert (o["status"] & 1 == 0).all(), (c0, grid)


def test_rotation_matches_lapack_rotation():
    """U^T X from the device against float64 LAPACK eigh + matmul, up to eigenvector signs (via |.|)
    and through rotation-invariant sums."""
    from pygemma_b200.synth import make_problem

    capi = _capi()
    p = make_problem(300, 64, 2, seed=5, m_k=900)  # m_k > n: no degenerate eigenspace
    with capi.Handle(300, 2) as h:
        d, _ = h.set_kinship(p["K"])
        h.set_design(p["W"], p["Y"])
        h.scan(p["X"])
        xr, row0 = h.probe_rotated(64)
    assert row0 == 0
    X = p["X"].astype(np.float64)
    w, U = np.linalg.eigh(p["K"])
    assert rel(d, np.maximum(w, 0)).max() < 1e-9
    ref = (U.T @ X).T
    assert np.abs(np.abs(xr) - np.abs(ref)).max() < 1e-9
    assert rel((xr ** 2).sum(1), (X ** 2).sum(0)).max() < 1e-12  # orthogonality: norms preserved


def test_empty_and_single_snp():
    capi = _capi()
    g = np.load(os.path.join(GOLDEN, "scan_interior.npz"))
    n = g["xr"].shape[0]
    with capi.Handle(n, g["wr"].shape[1]) as h:
        h.set_eigen(None, g["d"])
        h.set_design(g["wr"], g["yr"], already_rotated=True)
        o0 = h.scan(np.empty((n, 0)))
        assert o0["beta"].shape == (0,)
        o1 = h.scan(np.ascontiguousarray(g["xr"][:, 3:4]))
        assert rel(o1["lambda"][0], g["r64_lambda"][3]) < TOL
        # a strided view (every other column) must be honoured through ld
        o2 = h.scan(g["xr"][:, ::2])
        assert rel(o2["beta"], g["r64_beta"][::2]).max() < TOL


def test_full_size_properties():
    """BASELINE.json config 3 shape at full n (n = 10 000, c0 = 10), bounded m: size-independent properties.
    * scale invariance: beta(x*s) = beta(x)/s, identical lambda and p
    * covariate-shift invariance: x + 3*w_1 leaves beta, se, p unchanged (w_1 is projected out)
    * duplicate columns give bit-identical rows; row order follows column order."""
    from oracle import oracle
    from pygemma_b200.synth import make_spectral_problem

    capi = _capi()
    n, m, c0 = 10000, 192, 10
    p = make_spectral_problem(n, m, c0, seed=77, xdtype=np.float64)
    X = p["X"]
    X[:, 5] = X[:, 2]
    Xs = X * 4.0
    Xc = X + 3.0 * p["W"][:, 1:2]
    with capi.Handle(n, c0) as h:
        h.set_eigen(None, p["d"])
        h.set_design(p["W"], p["Y"], already_rotated=True)
        o = h.scan(X)
        os_ = h.scan(Xs)
        oc = h.scan(Xc)
        og = h.scan(X, grid=True)
    assert (o["status"] == 0).all()
    for c in COLS:
        assert o[c][5] == o[c][2], c
    assert rel(os_["beta"] * 4.0, o["beta"]).max() < 1e-9 and rel(os_["p_wald"], o["p_wald"]).max() < 1e-8
    assert rel(os_["lambda"], o["lambda"]).max() < 1e-8
    assert rel(oc["beta"], o["beta"]).max() < 1e-7 and rel(oc["p_wald"], o["p_wald"]).max() < 1e-7
    assert set(np.unique(og["lambda"])) <= set(10.0 ** np.arange(-5, 6))
    # first 24 SNPs against the oracle at full n
    ref = oracle.scan_rotated(p["d"], p["Y"], p["W"], np.ascontiguousarray(X[:, :24].T))
    _check(o, ref, idx=np.arange(24), tag="n=10000")


def test_i8split_rotation_equals_fp64_rotation():
    """The exact integer-split tensor-core rotation against the FP64 GEMM rotation on the same int8 dosages:
    rotated vectors agree to fp64 rounding of the FP64 path, scan outputs to 1e-9."""
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m = 777, 203  # odd sizes: exercises every padding path
    p = make_problem(n, m, 3, seed=9, m_k=2000)
    res = {}
    with capi.Handle(n, 3) as h:
        h.set_kinship(p["K"])
        h.set_design(p["W"], p["Y"])
        for eng in (capi.PG_ROT_FP64, capi.PG_ROT_I8SPLIT):
            h.set_options(rotation=eng, block_snps=128)
            o = h.scan(p["X"])
            assert o["timing"]["rot_engine"] == eng
            xr, row0 = h.probe_rotated(64)
            assert row0 == 128
            res[eng] = (o, xr)
        h.set_options(rotation=capi.PG_ROT_I8SPLIT)
        with pytest.raises(capi.PgError):  # float genotypes that are not level-codeable cannot take an int8 engine
            h.scan(p["X"].astype(np.float32) + np.random.default_rng(0).random(p["X"].shape, dtype=np.float32))
    (o64, x64), (o8, x8) = res[capi.PG_ROT_FP64], res[capi.PG_ROT_I8SPLIT]
    scale = np.sqrt((p["X"].astype(np.float64) ** 2).sum(0))[128:192, None]
    assert (np.abs(x64 - x8) / scale).max() < 1e-13
    for c in COLS:
        assert rel(o8[c], o64[c]).max() < 1e-9, c


def test_reml_engines_agree():
    """The three REML engines -- compressed moments (default), CTA-lock-step streaming, warp-per-SNP -- on the same
    rotated inputs: identical optimiser paths (evaluation counts) and results to 1e-9."""
    from pygemma_b200.synth import make_spectral_problem

    capi = _capi()
    for (n, m, c0, grid) in ((1300, 160, 14, False), (2500, 300, 3, False), (640, 100, 25, True)):
        p = make_spectral_problem(n, m, c0, seed=4 + c0, xdtype=np.float64)
        outs = {}
        with capi.Handle(n, c0) as h:
            h.set_eigen(None, p["d"])
            h.set_design(p["W"], p["Y"], already_rotated=True)
            for eng in (capi.PG_REML_COMPRESSED, capi.PG_REML_STREAM, capi.PG_REML_WARP):
                h.set_reml_engine(eng)
                outs[eng] = h.scan(p["X"], grid=grid)
                assert outs[eng]["timing"]["reml_engine"] == eng
        a = outs[capi.PG_REML_COMPRESSED]
        assert a["timing"]["n_nodes"] < n
        for eng in (capi.PG_REML_STREAM, capi.PG_REML_WARP):
            b = outs[eng]
            for c in COLS:
                assert rel(a[c], b[c]).max() < 1e-9, (n, eng, c, float(rel(a[c], b[c]).max()))
            assert np.array_equal(a["n_eval2"], b["n_eval2"]) and np.array_equal(a["n_eval3"], b["n_eval3"])


def test_unsorted_and_degenerate_spectra():
    """Eigenvalues supplied in arbitrary order, with a large null space, exact ties and isolated outliers: the handle
    sorts internally (compression needs ascending d); results must equal the oracle's on the caller's order."""
    from oracle import oracle

    capi = _capi()
    rng = np.random.default_rng(11)
    n, m, c0 = 900, 96, 4
    d = np.concatenate([np.zeros(300), np.full(200, 0.37), rng.chisquare(3, 395), [1e-12, 1e-30, 40.0, 900.0, 3e4]])
    rng.shuffle(d)
    W = np.c_[np.ones(n), rng.standard_normal((n, c0 - 1))]
    X = rng.standard_normal((n, m))
    y = rng.standard_normal(n) * np.sqrt(0.7 * d + 1.0) + 0.2 * X[:, 0]
    ref = oracle.scan_rotated(d, y, W, np.ascontiguousarray(X.T))
    with capi.Handle(n, c0) as h:
        h.set_eigen(None, d)
        h.set_design(W, y, already_rotated=True)
        for eng in (capi.PG_REML_COMPRESSED, capi.PG_REML_STREAM):
            h.set_reml_engine(eng)
            o = h.scan(X)
            _check(o, ref, tag=("unsorted", eng))
            o2 = h.scan(np.ascontiguousarray(X.T), layout=capi.PG_X_SNP_MAJOR)
            for c in COLS:
                assert np.array_equal(o[c], o2[c], equal_nan=True), c
        assert o["timing"]["n_nodes"] <= n


def test_transposed_gemm_operand_is_bit_identical_to_staged():
    """The int8 rotation feeds the sample-major block to the GEMM as a transposed operand (default) or stages it
    SNP-major first (PG_GEMM_TT=0): integer partial products, hence every output bit, must be identical."""
    import subprocess
    import sys
    import tempfile

    code = ("import numpy as np, sys; sys.path.insert(0, '.');"
            "from pygemma_b200 import _capi; from pygemma_b200.synth import make_problem;"
            "p = make_problem(640, 352, 3, seed=12, m_k=1500);"
            "h = _capi.Handle(640, 3); h.set_kinship(p['K']); h.set_design(p['W'], p['Y']);"
            "h.set_options(rotation=_capi.PG_ROT_I8SPLIT, block_snps=128);"
            "o = h.scan(p['X']); np.save(sys.argv[1], np.stack([o[c] for c in ['beta','se_beta','tau','lambda','F_wald','p_wald']]))")
    outs = []
    for tt in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            env = dict(os.environ, PG_GEMM_TT=tt)
            subprocess.check_call([sys.executable, "-c", code, f.name], env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
            outs.append(np.load(f.name))
    assert np.array_equal(outs[0], outs[1], equal_nan=True)


def test_fused_tcgen05_rotation_is_bit_identical_to_cublas_split():
    """PG_ROT_I8TC (TMA + tcgen05.mma kind::i8, seven planes in TMEM, recombination in the epilogue) against
    PG_ROT_I8SPLIT (cuBLAS int8 GEMM + combine kernel): exact integer partial products and the same single rounding,
    so rotated vectors and every scan output must agree bit for bit -- ragged SNP / eigenvector / sample counts included."""
    from pygemma_b200.synth import make_problem

    capi = _capi()
    for (n, m, blk) in ((777, 300, 0), (1024, 640, 256), (333, 129, 0)):
        p = make_problem(n, m, 3, seed=n, m_k=2 * n)
        res = {}
        with capi.Handle(n, 3) as h:
            h.set_kinship(p["K"])
            h.set_design(p["W"], p["Y"])
            for eng in (capi.PG_ROT_I8SPLIT, capi.PG_ROT_I8TC):
                h.set_options(rotation=eng, block_snps=blk)
                o = h.scan(p["X"])
                assert o["timing"]["rot_engine"] == eng
                xr, _ = h.probe_rotated(min(m, 128))
                res[eng] = (o, xr)
        (oa, xa), (ob, xb) = res[capi.PG_ROT_I8SPLIT], res[capi.PG_ROT_I8TC]
        assert np.array_equal(xa, xb), (n, m)
        for c in COLS:
            assert np.array_equal(oa[c], ob[c], equal_nan=True), (n, m, c)


def test_float_dosages_take_the_level_coded_int8_path():
    """Float genotypes as the reference's callers pass them -- raw 0.0/1.0/2.0 or standardised per SNP
    (experiments/wtccc/run_pygemma.py:432) -- take at most three values per column: they must go through the exact int8
    tensor-core rotation on their level codes (rot_engine == I8SPLIT; one pass when the levels are equally spaced to
    double rounding, code + indicator passes otherwise) and match the oracle run on the very same float values.
    Mean-imputed dosages (a fourth level) stay on that path; a block holding a five-valued column falls back to the
    FP64 GEMM."""
    from oracle import oracle
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m, c0 = 600, 200, 4
    p = make_problem(n, m, c0, seed=21, m_k=1500)
    g = p["X"].astype(np.float64)
    g[:, 7] = 1.0            # constant column (one level)
    g[:, 9] = (g[:, 9] > 0)  # two levels
    sd = g.std(axis=0)
    sd[sd == 0] = 1.0
    odd = g.copy()
    odd[:, 20] = np.where(g[:, 20] == 2, 10.0, g[:, 20])       # levels {0, 1, 10}: unequal spacing
    odd[:, 21] = np.where(g[:, 21] == 0, -3.5, g[:, 21] * 0.25)  # levels {-3.5, 0.25, 0.5}
    variants = {
        "odd_levels_f64": odd,
        "raw_f64": g,
        "raw_f32": g.astype(np.float32),
        "std_f64": (g - g.mean(axis=0)) / sd,
        "std_f32": ((g - g.mean(axis=0)) / sd).astype(np.float32),
    }
    with capi.Handle(n, c0) as h:
        h.set_kinship(p["K"])
        h.set_design(p["W"], p["Y"])
        for name, X in variants.items():
            ref = oracle.pygemma(p["Y"], X.astype(np.float64), p["W"], p["K"])
            for layout_snp in (False, True):
                Xin = np.ascontiguousarray(X.T) if layout_snp else np.ascontiguousarray(X)
                o = h.scan(Xin, layout=capi.PG_X_SNP_MAJOR if layout_snp else capi.PG_X_SAMPLE_MAJOR)
                # float32-standardised columns are not equally spaced to double rounding: they take the two-component
                # (code + indicator) form of the same int8 path
                assert o["timing"]["rot_engine"] in (capi.PG_ROT_I8SPLIT, capi.PG_ROT_I8TC), name
                ok = np.ones(m, dtype=bool)
                ok[7] = False  # constant column: degenerate (x is collinear with the intercept), finite garbage in both
                _check(o, ref, idx=np.where(ok)[0], tag=(name, layout_snp))
        # forcing the FP64 GEMM gives the same numbers to rounding
        h.set_options(rotation=capi.PG_ROT_FP64)
        o64 = h.scan(np.ascontiguousarray(variants["std_f64"]))
        assert o64["timing"]["rot_engine"] == capi.PG_ROT_FP64
        h.set_options(rotation=capi.PG_ROT_AUTO)
        o8 = h.scan(np.ascontiguousarray(variants["std_f64"]))
        # the library-GEMM form of the level-coded path gives the same bits as the fused kernel
        h.set_options(rotation=capi.PG_ROT_I8SPLIT)
        for name in ("std_f64", "std_f32", "odd_levels_f64"):
            ol = h.scan(np.ascontiguousarray(variants[name]))
            h.set_options(rotation=capi.PG_ROT_AUTO)
            oa = h.scan(np.ascontiguousarray(variants[name]))
            h.set_options(rotation=capi.PG_ROT_I8SPLIT)
            assert ol["timing"]["rot_engine"] == capi.PG_ROT_I8SPLIT and oa["timing"]["rot_engine"] == capi.PG_ROT_I8TC
            for c in COLS:
                assert np.array_equal(ol[c], oa[c], equal_nan=True), (name, c)
        h.set_options(rotation=capi.PG_ROT_AUTO)
        ok = np.arange(m) != 7
        for c in COLS:
            assert rel(o8[c][ok], o64[c][ok]).max() < 1e-8, c
        # mean-imputed dosages (SimpleImputer(strategy='mean'), experiments/animal_gwas/run_gwas.py:93-94): three equally
        # spaced levels + the column mean -> still the int8 path (code + indicator), raw and standardised
        gm = g.copy()
        miss = np.random.default_rng(3).random(gm.shape) < 0.03
        gm[miss] = np.nan
        gm = np.where(np.isnan(gm), np.nanmean(gm, axis=0), gm)
        sdm = gm.std(axis=0)
        sdm[sdm == 0] = 1.0
        for name, Xi in (("imputed_raw", gm), ("imputed_std", (gm - gm.mean(axis=0)) / sdm)):
            oi = h.scan(np.ascontiguousarray(Xi))
            assert oi["timing"]["rot_engine"] == capi.PG_ROT_I8TC, name
            refi = oracle.pygemma(p["Y"], Xi, p["W"], p["K"])
            _check(oi, refi, idx=np.where(ok)[0], tag=name)
        # a column with five distinct values makes the block dense (FP64 GEMM)
        Xd = variants["std_f64"].copy()
        Xd[3, 11] = 0.123456
        Xd[4, 11] = -0.654321
        od = h.scan(Xd)
        assert od["timing"]["rot_engine"] == capi.PG_ROT_FP64
        refd = oracle.pygemma(p["Y"], Xd, p["W"], p["K"])
        _check(od, refd, idx=np.where(ok)[0], tag="dense")


def test_grm_on_device_matches_reference_definition():
    """pg_grm against the reference's calculate_genetic_relatedness_matrix (experiments/animal_gwas/run_gwas.py:45-55)
    restated in NumPy, for int8 / float32 / float64 markers, both layouts, a constant column, several blocks; and the
    fused GRM -> syevd -> scan pipeline against the same steps with K passed through the host."""
    from pygemma_b200 import grm
    from pygemma_b200.synth import make_problem

    capi = _capi()
    rng = np.random.default_rng(4)
    n, p = 500, 3000
    G = rng.binomial(2, rng.uniform(0.05, 0.5, p), size=(n, p)).astype(np.int8)
    G[:, 17] = 1  # constant column: sd -> 1
    def ref(X):
        X = X.astype(np.float64)
        sd = X.std(axis=0)
        sd[sd == 0] = 1
        Z = (X - X.mean(axis=0)) / sd
        return Z @ Z.T / X.shape[1]
    Kref = ref(G)
    for X in (G, G.astype(np.float32), G.astype(np.float64)):
        K = grm.calculate_genetic_relatedness_matrix(X)
        assert K.shape == (n, n) and np.allclose(K, K.T, rtol=0, atol=0)
        assert np.abs(K - Kref).max() < 1e-12 * np.abs(Kref).max(), str(X.dtype)
    with capi.Handle(n, 1) as h:
        Ks = h.grm(np.ascontiguousarray(G.T), layout=capi.PG_X_SNP_MAJOR)["K"]
    assert np.abs(Ks - Kref).max() < 1e-12 * np.abs(Kref).max()
    # GRM kept on the device and eigendecomposed in place == K through the host
    pr = make_problem(n, 64, 3, seed=2, m_k=400)
    with capi.Handle(n, 3) as h:
        r = h.grm(G, return_K=False, set_kinship=True)
        assert (np.diff(r["d"]) >= 0).all() and r["d"].min() >= 0
        h.set_design(pr["W"], pr["Y"])
        o1 = h.scan(pr["X"])
    with capi.Handle(n, 3) as h:
        h.set_kinship(Kref)
        h.set_design(pr["W"], pr["Y"])
        o2 = h.scan(pr["X"])
    for c in COLS:
        assert rel(o1[c], o2[c]).max() < 1e-7, (c, float(rel(o1[c], o2[c]).max()))


def test_pageable_host_input_bounce_path_is_identical():
    """Large pageable genotype arrays are packed into pinned bounce buffers by host threads (pg_scan) instead of a
    pageable 2-D copy: same bits as the direct copy (PG_NO_BOUNCE=1) and as a strided view of a wider array."""
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m = 2048, 36000  # 74 MB of int8: above the 64 MB threshold of the bounce path
    p = make_problem(n, 64, 3, seed=8, m_k=3000)
    rng = np.random.default_rng(1)
    X = rng.integers(0, 3, size=(n, m), dtype=np.int8)
    wide = rng.integers(0, 3, size=(n, m + 500), dtype=np.int8)
    wide[:, 100:100 + m] = X
    outs = []
    with capi.Handle(n, 3) as h:
        h.set_kinship(p["K"])
        h.set_design(p["W"], p["Y"])
        h.set_options(block_snps=8192)
        for env, arr in (("", X), ("1", X), ("", wide[:, 100:100 + m])):
            if env:
                os.environ["PG_NO_BOUNCE"] = env
            else:
                os.environ.pop("PG_NO_BOUNCE", None)
            outs.append(h.scan(arr))
        os.environ.pop("PG_NO_BOUNCE", None)
    for o in outs[1:]:
        for c in COLS:
            assert np.array_equal(outs[0][c], o[c], equal_nan=True), c


def test_packed_bed_scan_matches_oracle_on_imputed_genotypes(tmp_path):
    """PLINK .bed ingest: packed 2-bit genotypes are decoded, mean-imputed and (optionally) standardised on the device and
    rotated on the exact int8 path (codes + missing indicator).  Against the oracle fed with the NumPy decode + np.nanmean
    imputation (+ StandardScaler arithmetic) the reference's callers perform; both allele-count conventions; a column
    without missing calls, one that is all missing, ragged n; and the file-level entry point."""
    from oracle import oracle
    from pygemma_b200 import bed
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m, c0 = 602, 150, 3   # n % 4 != 0: padding bits in the last byte
    p = make_problem(n, m, c0, seed=31, m_k=1500)
    rng = np.random.default_rng(5)
    G = p["X"].astype(np.float64)
    G[rng.random(G.shape) < 0.04] = np.nan
    G[:, 3] = p["X"][:, 3]          # no missing call
    G[:, 5] = np.nan                # all missing -> constant column
    packed = bed.encode_packed(G)
    def impute(D, standardize):
        mu = np.where(np.isnan(D).all(axis=0), 0.0, np.nanmean(np.where(np.isnan(D).all(axis=0), 0.0, D), axis=0))
        X = np.where(np.isnan(D), mu, D)
        if standardize:
            sd = X.std(axis=0)
            sd[sd == 0] = 1.0
            X = (X - X.mean(axis=0)) / sd
        return X
    ok = np.arange(m) != 5
    with capi.Handle(n, c0) as h:
        h.set_kinship(p["K"])
        h.set_design(p["W"], p["Y"])
        h.set_options(block_snps=64)
        for a1 in (False, True):
            D = bed.decode_packed(packed, n, count_A1=a1)
            for std in (False, True):
                o = h.scan_bed(packed, count_A1=a1, standardize=std)
                assert o["timing"]["rot_engine"] == capi.PG_ROT_I8TC and o["beta"].shape == (m,)
                ref = oracle.pygemma(p["Y"], impute(D, std), p["W"], p["K"])
                _check(o, ref, idx=np.where(ok)[0], tag=("bed", a1, std))
        h.set_options(rotation=capi.PG_ROT_I8SPLIT, block_snps=64)
        o2 = h.scan_bed(packed)
        h.set_options(rotation=capi.PG_ROT_AUTO, block_snps=64)
        o1 = h.scan_bed(packed)
        for c in COLS:
            assert np.array_equal(o1[c], o2[c], equal_nan=True), c
    prefix = str(tmp_path / "toy")
    bed.write_bed(prefix, G)
    df = bed.pygemma_bed(p["Y"], prefix, p["W"], p["K"])
    assert list(df.columns) == COLS + ["SNPs"] and list(df["SNPs"][:3]) == ["rs0", "rs1", "rs2"]
    for c in COLS:
        assert np.array_equal(df[c].values, o1[c], equal_nan=True), c


def test_multi_phenotype_scan_equals_one_scan_per_trait():
    """pg_set_design_multi: genotype blocks are rotated once and the REML stage runs per trait; every trait's rows
    are bit-identical to a single-trait pg_set_design + pg_scan, for every REML engine, several blocks, grid mode,
    and through lmm.pygemma_multi; pg_set_design afterwards returns the handle to one trait."""
    from pygemma_b200 import lmm
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m, c0, q = 900, 2500, 4, 3
    p = make_problem(n, m, c0, seed=21, m_k=2000)
    rng = np.random.default_rng(5)
    Y = np.stack([p["Y"].reshape(-1), rng.standard_normal(n),
                  p["Y"].reshape(-1) * 0.3 + p["X"][:, 7] * 0.2 + rng.standard_normal(n)], axis=1)
    keys = COLS + ["status", "n_eval2", "n_eval3"]
    with capi.Handle(n, c0) as h:
        h.set_kinship(p["K"])
        h.set_options(block_snps=1024)  # several blocks, ragged tail
        for engine in (capi.PG_REML_AUTO, capi.PG_REML_STREAM, capi.PG_REML_WARP):
            h.set_reml_engine(engine)
            for grid in (False, True):
                h.set_design(p["W"], Y)
                om = h.scan(p["X"], grid=grid)
                assert om["beta"].shape == (q, m)
                for ph in range(q):
                    h.set_design(p["W"], Y[:, ph])
                    o1 = h.scan(p["X"], grid=grid)
                    assert o1["beta"].shape == (m,)
                    for c in keys:
                        assert np.array_equal(om[c][ph], o1[c], equal_nan=True), (engine, grid, ph, c)
    frames = lmm.pygemma_multi(Y, p["X"], p["W"], p["K"], snps=p["snps"] if "snps" in p else None)
    assert len(frames) == q
    for ph in range(q):
        one = lmm.pygemma(Y[:, ph], p["X"], p["W"], p["K"])
        for c in COLS:
            assert np.array_equal(frames[ph][c].to_numpy(), one[c].to_numpy(), equal_nan=True), (ph, c)


def test_more_traits_than_one_pass_holds():
    """pg_set_design_multi takes at most PG_MAX_TRAITS traits (PG_ERR_ARG beyond); lmm.pygemma_multi scans in groups and
    still returns one frame per trait, identical to single-trait calls."""
    from pygemma_b200 import lmm
    from pygemma_b200.synth import make_problem

    capi = _capi()
    n, m, c0 = 300, 150, 3
    q = capi.PG_MAX_TRAITS + 3
    p = make_problem(n, m, c0, seed=33, m_k=600)
    rng = np.random.default_rng(9)
    Y = p["Y"].reshape(-1, 1) * rng.uniform(0.2, 1.0, size=(1, q)) + rng.standard_normal((n, q))
    with capi.Handle(n, c0) as h:
        h.set_kinship(p["K"])
        with pytest.raises(capi.PgError):
            h.set_design(p["W"], Y)
        h.set_design(p["W"], Y[:, :capi.PG_MAX_TRAITS])
        assert h.scan(p["X"])["beta"].shape == (capi.PG_MAX_TRAITS, m)
    frames = lmm.pygemma_multi(Y, p["X"], p["W"], p["K"])
    assert len(frames) == q
    for ph in (0, capi.PG_MAX_TRAITS - 1, capi.PG_MAX_TRAITS, q - 1):
        one = lmm.pygemma(Y[:, ph], p["X"], p["W"], p["K"])
        for c in COLS:
            assert np.array_equal(frames[ph][c].to_numpy(), one[c].to_numpy(), equal_nan=True), (ph, c)


def test_moment_fusion_matches_the_unfused_path_and_the_oracle():
    """pg_set_moment_fusion(1): the fused tcgen05 rotation produces the eigenvalue-space moments itself -- linear moments
    as extra exact int8 tiles against G = U V, x^2 moments in its epilogue -- and never writes the rotated genotypes.
    Same results as the unfused pipeline to rounding (1e-9: only the summation order differs) and the oracle's tolerance,
    for int8 dosages (direct MN-major operand and the staged copy for n % 16 != 0), a spectrum with a null space and
    isolated eigenvalues (COPY rows), several blocks with a ragged tail, level-coded float genotypes (affine fix-up in
    the fused epilogue; a block that needs the eps pass is compressed as before), several traits, grid mode and the LRT
    columns; fused scans are deterministic and independent of the block a SNP sits in."""
    from oracle import oracle

    capi = _capi()
    rng = np.random.default_rng(17)
    for n, m, c0 in ((1024, 1700, 5), (1000, 777, 3)):
        U, _ = np.linalg.qr(rng.standard_normal((n, n)))
        d = np.sort(10.0 ** (rng.random(n) * 1.5 - 1.0))
        d[: n // 100] = 0.0                              # null space: one COMPRESS node at d = 0
        d[-3:] *= np.array([3.0, 10.0, 40.0])           # isolated eigenvalues: COPY rows
        W = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(c0 - 1)])
        u = U @ (np.sqrt(d) * rng.standard_normal(n))
        y = 0.7 * u / u.std() + 0.7 * rng.standard_normal(n) + 0.05 * W[:, 1:].sum(1)
        maf = rng.random(m) * 0.45 + 0.05
        X = ((rng.random((n, m)) < maf).astype(np.int8) + (rng.random((n, m)) < maf).astype(np.int8))
        idx = np.unique(np.linspace(0, m - 1, 40).astype(np.int64))
        ref = oracle.scan_rotated(d, U.T @ y, U.T @ W, np.ascontiguousarray((U.T @ X[:, idx].astype(np.float64)).T))
        with capi.Handle(n, c0) as h:
            h.set_eigen(U, d)
            h.set_design(W, y)
            h.set_options(block_snps=512)   # several blocks, ragged tail
            assert not h.fusion_info()["fused"]
            o0 = h.scan(X)
            assert o0["timing"]["rot_engine"] == capi.PG_ROT_I8TC
            h.set_moment_fusion(1)
            info = h.fusion_info()
            assert info["fused"] and info["g_columns"] > 0 and info["pieces"] > 0
            o1 = h.scan(X)
            assert o1["timing"]["rot_engine"] == capi.PG_ROT_I8TC_MOMENTS and o1["timing"]["compress_ms"] < 0.05
            assert np.array_equal(o0["status"], o1["status"])
            for c in COLS:
                assert rel(o1[c], o0[c]).max() < 1e-9, (n, c, float(rel(o1[c], o0[c]).max()))
            _check({c: o1[c][idx] for c in COLS}, ref, tag=("fused", n))
            # deterministic, and a SNP's bits do not depend on its block or its position in a tile
            o1b = h.scan(X)
            h.set_options(block_snps=0)
            o1c = h.scan(X[:, 300:1200 if m > 1200 else m])
            for c in COLS:
                assert np.array_equal(o1[c], o1b[c], equal_nan=True), c
                assert np.array_equal(o1[c][300:300 + o1c[c].shape[0]], o1c[c], equal_nan=True), c
            with pytest.raises(capi.PgError):
                h.probe_rotated(4)              # a fused scan leaves no rotated genotypes behind
            # grid mode and the likelihood-ratio columns read the same moments
            og0 = h.scan(X, grid=True)
            h.set_moment_fusion(0)
            og1 = h.scan(X, grid=True)
            assert np.array_equal(og0["lambda"], og1["lambda"])
            ol0 = h.scan(X[:, :256], lrt=True)
            h.set_moment_fusion(1)
            ol1 = h.scan(X[:, :256], lrt=True)
            for c in ("D_lrt", "loglik_ml"):
                assert np.abs(ol1[c] - ol0[c]).max() < 1e-8, c
            # level-coded float genotypes: standardised in float64 -> affine, fused; float32 -> eps pass, not fused
            g = X[:, :300].astype(np.float64)
            sd = g.std(axis=0)
            sd[sd == 0] = 1.0
            Xs = (g - g.mean(axis=0)) / sd
            of = h.scan(np.ascontiguousarray(Xs))
            assert of["timing"]["rot_engine"] == capi.PG_ROT_I8TC_MOMENTS
            o32 = h.scan(np.ascontiguousarray(Xs.astype(np.float32)))
            assert o32["timing"]["rot_engine"] == capi.PG_ROT_I8TC
            h.set_moment_fusion(0)
            of0 = h.scan(np.ascontiguousarray(Xs))
            ok = sd[:300] > 0
            ok &= g.std(axis=0) > 0
            for c in COLS:
                assert rel(of[c][ok], of0[c][ok]).max() < 1e-8, ("affine", c)
            # several traits share the tiles of G built from [W0, y_0 .. y_{q-1}]
            Y = np.stack([y, rng.standard_normal(n), 0.3 * y + rng.standard_normal(n)], axis=1)
            h.set_design(W, Y)
            om0 = h.scan(X[:, :600])
            h.set_moment_fusion(1)
            om1 = h.scan(X[:, :600])
            assert om1["timing"]["rot_engine"] == capi.PG_ROT_I8TC_MOMENTS and om1["beta"].shape == (3, min(600, m))
            for c in COLS:
                assert rel(om1[c], om0[c]).max() < 1e-9, ("multi", c)


def test_cta_parallel_table_elimination_is_bit_identical_to_the_serial_kernel(tmp_path):
    """The table-2 rows (covariate levels eliminated once per table lambda, pyx:1011-1031 restricted to [W0, y]) are built
    by one CTA per lambda; PG_ELIM_SERIAL=1 selects the one-thread-per-lambda kernel (read once per process, hence the two
    subprocesses).  Same operations per entry in the same order: every output of a scan (Brent + Newton and grid, c0 = 1,
    12 and 40) agrees bit for bit."""
    import subprocess
    import sys

    from pygemma_b200.synth import make_problem

    capi = _capi()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from pygemma_b200 import _capi\n"
        "from pygemma_b200.synth import make_problem\n"
        "out = {}\n"
        "for c0 in (1, 12, 40):\n"
        "    p = make_problem(700, 96, c0, seed=40 + c0, m_k=900)\n"
        "    with _capi.Handle(700, c0) as h:\n"
        "        h.set_kinship(p['K']); h.set_design(p['W'], p['Y'])\n"
        "        for grid in (False, True):\n"
        "            o = h.scan(p['X'], grid=grid)\n"
        "            for c in ('beta', 'se_beta', 'tau', 'lambda', 'F_wald', 'p_wald'):\n"
        "                out[f'{c0}_{int(grid)}_{c}'] = o[c]\n"
        "np.savez(sys.argv[1], **out)\n"
    )
    files = {}
    for serial in ("0", "1"):
        files[serial] = str(tmp_path / f"elim_{serial}.npz")
        env = dict(os.environ, PG_ELIM_SERIAL=serial)
        r = subprocess.run([sys.executable, "-c", code, files[serial]], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
    a, b = np.load(files["0"]), np.load(files["1"])
    assert sorted(a.files) == sorted(b.files) and len(a.files) == 36
    for k in a.files:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert capi is not None
