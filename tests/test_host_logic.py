"""CPU tests of the product's host/device scalar code (compiled for the host by tests/host_shim.cpp)
and of the Python host layer.  The optimiser state machine, the lambda tables + Chebyshev interpolation,
the Pab recursion and the F-tail are the same source the CUDA kernels compile."""
import glob
import os

import numpy as np
import pytest

import hostshim
from conftest import COLS, GOLDEN
from oracle import oracle
from pygemma_b200 import multi


TOL = 1e-6  # north star: lambda, beta, se, p within 1e-6 relative


def rel(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


SCANS = sorted(glob.glob(os.path.join(GOLDEN, "scan_*.npz")))


@pytest.mark.parametrize("path", SCANS, ids=[os.path.basename(p)[5:-4] for p in SCANS])
def test_solver_with_tables_matches_ref64(path):
    g = np.load(path)
    xt = np.ascontiguousarray(g["xr"].T)
    o = hostshim.scan(g["d"], g["yr"], g["wr"], xt, grid=bool(g["grid"]))
    assert (o["status"] == 0).all()
    ok = np.array([0, 2, 4, 5, 6, 7]) if "degenerate" in path else np.arange(xt.shape[0])
    for c in COLS:
        tol = 1e-6 if c == "lambda" else 1e-8
        assert rel(o[c][ok], g[f"r64_{c}"][ok]).max() < tol, (c, rel(o[c][ok], g[f"r64_{c}"][ok]).max())


def test_interpolated_tables_equal_exact_tables():
    g = np.load(os.path.join(GOLDEN, "scan_interior.npz"))
    xt = np.ascontiguousarray(g["xr"].T)
    a = hostshim.scan(g["d"], g["yr"], g["wr"], xt, exact_w0y=False)
    b = hostshim.scan(g["d"], g["yr"], g["wr"], xt, exact_w0y=True)
    for c in COLS:
        assert rel(a[c], b[c]).max() < 1e-10, c
    assert np.array_equal(a["n_eval2"], b["n_eval2"]) and np.array_equal(a["n_eval3"], b["n_eval3"])


def test_interpolation_error_bound():
    rng = np.random.default_rng(3)
    n, c0 = 600, 4
    L = hostshim.lib()
    for spectrum in ("mp", "wide"):
        d = rng.chisquare(4, n) / 4 if spectrum == "mp" else 10.0 ** rng.uniform(-8, 4, n)
        wy = np.asfortranarray(np.c_[np.ones(n), rng.standard_normal((n, c0))])
        for lam in 10.0 ** rng.uniform(-5, 5, 12):
            for p in (1, 2, 3):
                e = L.pgh_interp_error(n, c0, hostshim._p(np.ascontiguousarray(d)), hostshim._p(wy), float(lam), p)
                assert e < 2e-13, (spectrum, lam, p, e)


def test_f_sf_matches_scipy():
    from scipy import stats

    L = hostshim.lib()
    rng = np.random.default_rng(1)
    for nu in (3, 17, 443, 1929, 9989, 49989):
        F = np.concatenate([10.0 ** rng.uniform(-12, 3.3, 300), [0.0, 1e-300, 1.0, 50.0, 1400.0]])
        ref = stats.f.sf(F, 1, nu)
        got = np.array([L.pgh_f_sf(float(f), float(nu)) for f in F])
        m = ref > 1e-300
        assert rel(got[m], ref[m]).max() < 1e-9, (nu, rel(got[m], ref[m]).max())
    assert np.isnan(L.pgh_f_sf(float("nan"), 10.0))
    assert L.pgh_f_sf(float("inf"), 10.0) == 0.0


def test_brent_state_machine_is_scipy_brentq():
    import ctypes

    from scipy import optimize

    L = hostshim.lib()
    rng = np.random.default_rng(9)
    checked = 0
    for _ in range(6000):
        p = np.ascontiguousarray(rng.standard_normal(5) * rng.choice([0.0, 1.0], size=5, p=[0.3, 0.7]))
        a = 10.0 ** rng.integers(-5, 5)
        b = a * 10.0
        f = lambda x: oracle.poly_eval(p, x)
        fa, fb = f(a), f(b)
        if fa == 0 or fb == 0 or np.signbit(fa) == np.signbit(fb):
            continue
        xs_ref = []

        def g(x):
            xs_ref.append(x)
            return f(x)

        r_sp = optimize.brentq(g, a, b, rtol=0.1, maxiter=100, disp=False)
        xs = np.empty(256)
        nc = ctypes.c_int(0)
        r = L.pgh_brent(hostshim._p(p), a, b, 2e-12, 0.1, 100, hostshim._p(xs), 256, ctypes.byref(nc))
        # the shim evaluates the same polynomial with its own exp(): compare roots to rounding, counts exactly
        assert nc.value == len(xs_ref)
        assert abs(r - r_sp) <= 1e-12 * abs(r_sp)
        checked += 1
    assert checked > 300


def test_shard_range_is_sampleiter():
    """multi.shard_range reproduces SampleIter's chunking (reference lmm/lmm.py:427-434): contiguous chunks of
    ceil(m / nproc) columns in rank order -- rounded up to 128 columns once chunks are that long, so that device-resident
    shards start on aligned columns (results do not depend on the chunk boundaries)."""
    for m in (1, 7, 100, 101, 12226, 100000):
        for world in (1, 2, 3, 8):
            per = multi.shard_len(m, world)
            ceil = int(np.ceil(m / world))
            assert per == ceil or (world > 1 and ceil >= 128 and per % 128 == 0 and 0 <= per - ceil < 128)
            cover = []
            for r in range(world):
                a, b = multi.shard_range(m, r, world)
                assert a == min(r * per, m) and b == min((r + 1) * per, m)
                cover += list(range(a, b))
            assert cover == list(range(m))


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    m, world = 11, 3
    full = {k: rng.standard_normal(m) for k in multi.RESULT_KEYS[:6]}
    full.update({k: rng.integers(0, 30, m).astype(np.int32) for k in multi.RESULT_KEYS[6:]})
    per = -(-m // world)
    blocks = []
    for r in range(world):
        a, b = multi.shard_range(m, r, world)
        blocks.append(multi.pack_results({k: v[a:b] for k, v in full.items()}, per))
    out = multi.unpack_results(blocks, m, world)
    for k in multi.RESULT_KEYS:
        assert np.array_equal(out[k], full[k]), k


def test_pygemma_argument_contract():
    from pygemma_b200 import lmm

    n, m = 20, 5
    Y, X, W, K = np.zeros(n), np.zeros((n, m), dtype=np.int8), np.ones((n, 1)), np.eye(n)
    with pytest.raises(ValueError):
        lmm.pygemma(Y[:-1], X, W, K)
    with pytest.raises(ValueError):
        lmm.pygemma(Y, X, W, np.eye(n + 1))
    Xf = X.astype(np.float64)
    Xf[0, 0] = np.nan
    with pytest.raises(ValueError, match="NaNs present in data"):
        lmm.pygemma(Y, Xf, W, K, disable_checks=False)
    assert lmm._as_genotypes(np.zeros((2, 2), dtype=np.int64)).dtype == np.int8
    assert lmm._as_genotypes(np.full((2, 2), 300)).dtype == np.float64
    assert lmm._as_genotypes(np.zeros((2, 2), dtype=np.float16)).dtype == np.float32


def test_pygemma_multi_groups_traits_and_keeps_order(monkeypatch):
    """lmm.pygemma_multi on a recording stand-in for the device handle: traits go through in groups of PG_MAX_TRAITS,
    one genotype pass per group, frames come back in trait order with the reference's columns, failed rows as NaN."""
    from pygemma_b200 import _capi, lmm

    calls = []

    class FakeHandle:
        def __init__(self, n, c0, device=0):
            self.n, self.c0, self.q = n, c0, 1

        def set_kinship(self, K):
            return 0.0

        def set_eigen(self, U, d):
            pass

        def set_scan_mode(self, mode=0):
            calls.append(("mode", mode))

        def set_design(self, W, y, already_rotated=False):
            y = np.asarray(y)
            self.q = y.shape[1] if y.ndim == 2 else 1
            self.first = float(y.reshape(self.n, -1)[0, 0])   # tags the group by its first trait's first value
            calls.append(("design", self.q))
            return 1.0

        def scan(self, X, grid=False, lrt=False):
            m = X.shape[1]
            calls.append(("scan", self.q))
            shape = m if self.q == 1 else (self.q, m)
            base = self.first + np.arange(self.q).reshape(-1, 1) + np.zeros((self.q, m))
            out = {k: (base + i).reshape(shape) for i, k in enumerate(lmm.COLUMNS)}
            st = np.zeros(shape, dtype=np.int32)
            st.reshape(self.q, m)[:, 1] = 1   # second SNP fails for every trait
            out.update(status=st, n_eval2=np.full(shape, 14, dtype=np.int32), n_eval3=np.full(shape, 3, dtype=np.int32))
            out["timing"] = {"rotate_ms": 0.0, "reml_ms": 0.0}
            return out

        def close(self):
            pass

    monkeypatch.setattr(_capi, "Handle", FakeHandle)
    monkeypatch.setattr(multi, "setup_eigen", lambda ctx, h, K: 0.0)
    n, m = 12, 4
    q = _capi.PG_MAX_TRAITS + 5
    Y = np.tile(np.arange(q, dtype=np.float64), (n, 1))      # trait t is the constant t
    X, W, K = np.zeros((n, m), dtype=np.int8), np.ones((n, 1)), np.eye(n)
    frames = lmm.pygemma_multi(Y, X, W, K, snps=[f"rs{i}" for i in range(m)])
    assert calls == [("mode", 0), ("design", _capi.PG_MAX_TRAITS), ("scan", _capi.PG_MAX_TRAITS), ("design", 5), ("scan", 5)]
    assert len(frames) == q
    for t, df in enumerate(frames):
        assert list(df.columns) == lmm.COLUMNS + ["SNPs"] and len(df) == m
        assert df["beta"].iloc[0] == float(t) and df["se_beta"].iloc[0] == float(t) + 1
        assert df.iloc[1][lmm.COLUMNS].isna().all() and not df.iloc[0][lmm.COLUMNS].isna().any()
    calls.clear()
    one = lmm.pygemma(Y[:, 3], X, W, K)
    assert calls == [("mode", 0), ("design", 1), ("scan", 1)] and one["beta"].iloc[0] == 3.0
    calls.clear()
    lmm.pygemma(Y[:, 3], X, W, K, de=True, grid=True)   # de: role-swapped scan, grid not forwarded (lmm/lmm.py:504)
    assert calls[0] == ("mode", _capi.PG_SCAN_DE)


# ---- eigenvalue-space compression (pygemma_b200/csrc/compress_plan.h) ----------------------------------

def _spectra(rng, n):
    mk = 2 * n
    g = rng.standard_normal((n, mk))
    mp = np.linalg.eigvalsh(g @ g.T / mk + 1e-3 * np.eye(n))
    lowrank = np.concatenate([np.zeros(n // 3), rng.chisquare(3, n - n // 3)])
    wide = 10.0 ** rng.uniform(-9, 4, n)
    clustered = np.concatenate([np.full(n // 2, 0.37), 0.37 * (1 + 1e-9 * rng.random(n - n // 2 - 5)), [1e-30, 5.0, 6.0, 7e3, 1e-12]])
    return {"mp": mp, "lowrank": lowrank, "wide": wide, "clustered": clustered}


def test_compression_error_bound():
    """sum_l a_l h_l^p from the compressed nodes agrees with the direct sum to rounding, for every lambda."""
    rng = np.random.default_rng(5)
    n = 1500
    L = hostshim.lib()
    for name, d in _spectra(rng, n).items():
        d = np.sort(np.maximum(d, 0.0))
        kc = L.pgh_plan_nodes(n, hostshim._p(d))
        assert kc <= n
        if name in ("mp", "lowrank", "clustered"):
            assert kc < n // 4, (name, kc)
        for a in (rng.standard_normal(n), rng.standard_normal(n) ** 2):
            a = np.ascontiguousarray(a)
            for p in (1, 2, 3):
                e = L.pgh_compress_error(n, hostshim._p(d), hostshim._p(a), p)
                assert e < (5e-14 if p == 3 else 2e-15), (name, p, e)


def test_fused_moment_plan_reproduces_the_compressed_moments():
    """compress_plan.h: build_fused_plan -- the bookkeeping of the x^2 moments the fused rotation forms in its epilogue (one
    partial sum per 16-eigenvector chunk and COMPRESS segment, reduced per segment in eigen order; COPY rows to their node).
    Replayed on the host it must give the compressed moments of their definition, on every kind of spectrum, and keep the
    structure the kernel relies on (checked inside the shim: -1 on violation)."""
    import ctypes

    rng = np.random.default_rng(9)
    L = hostshim.lib()
    for n in (449, 1500, 2048):
        for name, d in _spectra(rng, n).items():
            d = np.sort(np.maximum(d, 0.0))
            a = np.ascontiguousarray(rng.standard_normal(n) ** 2)
            for tile, piece in ((32, 16), (32, 32)):
                npieces, cnodes = ctypes.c_int32(0), ctypes.c_int32(0)
                e = L.pgh_fused_plan_check(n, hostshim._p(d), hostshim._p(a), tile, piece, ctypes.byref(npieces),
                                           ctypes.byref(cnodes))
                assert 0.0 <= e < 1e-13, (n, name, tile, piece, e)
                kc = L.pgh_plan_nodes(n, hostshim._p(d))
                assert cnodes.value <= kc
                # at most one piece per chunk and segment: chunks + segment boundaries bound the count
                assert npieces.value <= (n + piece - 1) // piece + kc, (n, name, npieces.value)


@pytest.mark.parametrize("path", SCANS, ids=[os.path.basename(p)[5:-4] for p in SCANS])
def test_compressed_scan_matches_ref64(path):
    g = np.load(path)
    xt = np.ascontiguousarray(g["xr"].T)
    o = hostshim.scan_compressed(g["d"], g["yr"], g["wr"], xt, grid=bool(g["grid"]))
    assert (o["status"] == 0).all()
    ok = np.array([0, 2, 4, 5, 6, 7]) if "degenerate" in path else np.arange(xt.shape[0])
    for c in COLS:
        tol = 1e-6 if c == "lambda" else 1e-8
        assert rel(o[c][ok], g[f"r64_{c}"][ok]).max() < tol, (c, rel(o[c][ok], g[f"r64_{c}"][ok]).max())


@pytest.mark.parametrize("path", SCANS, ids=[os.path.basename(p)[5:-4] for p in SCANS])
def test_xrow_form_equals_full_recursion(path):
    """Covariate levels from the table-2 rows + x-row recursion vs the full (c0+2)^2 recursion per evaluation:
    same optimiser path, results equal far below the parity tolerance."""
    g = np.load(path)
    xt = np.ascontiguousarray(g["xr"].T)
    a = hostshim.scan_compressed(g["d"], g["yr"], g["wr"], xt, grid=bool(g["grid"]), xrow_form=False)
    b = hostshim.scan_compressed(g["d"], g["yr"], g["wr"], xt, grid=bool(g["grid"]), xrow_form=True)
    ok = np.array([0, 2, 4, 5, 6, 7]) if "degenerate" in path else np.arange(xt.shape[0])
    for c in COLS:  # lambda itself is ill-conditioned where the likelihood is flat (boundary_lo: 7e-9)
        assert rel(a[c][ok], b[c][ok]).max() < (1e-7 if c == "lambda" else 1e-9), (c, rel(a[c][ok], b[c][ok]).max())
    assert np.array_equal(a["n_eval2"][ok], b["n_eval2"][ok]) and np.array_equal(a["n_eval3"][ok], b["n_eval3"][ok])


def test_compressed_scan_equals_direct_scan():
    """Same optimiser path (evaluation counts) and results to 1e-9 with and without the compression."""
    g = np.load(os.path.join(GOLDEN, "scan_interior.npz"))
    xt = np.ascontiguousarray(g["xr"].T)
    a = hostshim.scan(g["d"], g["yr"], g["wr"], xt)
    b = hostshim.scan_compressed(g["d"], g["yr"], g["wr"], xt)
    assert b["nodes"] < g["d"].shape[0]
    for c in COLS:
        assert rel(a[c], b[c]).max() < 1e-9, (c, rel(a[c], b[c]).max())
    assert np.array_equal(a["n_eval2"], b["n_eval2"]) and np.array_equal(a["n_eval3"], b["n_eval3"])


def test_bed_pack_unpack_and_files(tmp_path):
    """pygemma_b200.bed: NumPy decode / encode of packed PLINK genotypes and the file reader (header, .bim / .fam counts)."""
    from pygemma_b200 import bed

    rng = np.random.default_rng(0)
    for n in (5, 8, 13, 64):
        G = rng.integers(0, 3, size=(n, 7)).astype(np.float64)
        G[rng.random(G.shape) < 0.2] = np.nan
        for a1 in (False, True):
            packed = bed.encode_packed(G, count_A1=a1)
            assert packed.shape == (7, (n + 3) // 4) and packed.dtype == np.uint8
            back = bed.decode_packed(packed, n, count_A1=a1)
            assert np.array_equal(back, G, equal_nan=True)
    # PLINK's own convention on a hand-made byte: samples 0..3 = 00 (hom A1), 01 (missing), 10 (het), 11 (hom A2)
    byte = np.array([[0b11100100]], dtype=np.uint8)
    assert np.array_equal(bed.decode_packed(byte, 4, count_A1=False)[:, 0], [0.0, np.nan, 1.0, 2.0], equal_nan=True)
    assert np.array_equal(bed.decode_packed(byte, 4, count_A1=True)[:, 0], [2.0, np.nan, 1.0, 0.0], equal_nan=True)
    prefix = str(tmp_path / "toy")
    G = rng.integers(0, 3, size=(11, 5)).astype(np.float64)
    G[3, 2] = np.nan
    bed.write_bed(prefix, G)
    packed, n, ids = bed.read_packed(prefix)
    assert n == 11 and packed.shape == (5, 3) and list(ids) == [f"rs{i}" for i in range(5)]
    assert np.array_equal(bed.read_bed(prefix), G, equal_nan=True)
    open(prefix + ".bed", "r+b").write(b"\x00")
    with pytest.raises(ValueError):
        bed.read_packed(prefix)


def test_de_mode_solver_matches_reference_cpdefs():
    """The role-swapped final level (pg_eval.cuh: xrow_final_level_swapped, PG_SCAN_DE) through the product's scalar
    headers on the CPU, against the reference's cpdefs called in calculate_de's roles (tests/golden/de_small.npz)."""
    g = np.load(os.path.join(GOLDEN, "de_small.npz"))
    o = hostshim.scan_compressed_ex(g["d"], g["yr"], g["wr"], np.ascontiguousarray(g["xr"].T), swap=True)
    assert (o["status"] == 0).all()
    for c in COLS:
        assert rel(o[c], g[f"r64_{c}"]).max() < TOL, c


@pytest.mark.parametrize("name", ["lrt_interior", "lrt_low_h2", "lrt_c8"])
def test_ml_solver_matches_reference_functions(name):
    """pg::MlSolver (pg_math.cuh) + table-2 finals on the CPU against the reference's lmm.calc_lambda / likelihood_lambda
    (tests/golden/lrt_*.npz): null model and every alternative model."""
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    o = hostshim.scan_compressed_ex(g["d"], g["yr"], g["wr"], np.ascontiguousarray(g["xr"].T), lrt=True)
    assert (o["status"] == 0).all()
    assert rel(o["lambda_null"], g["r64_lambda_null"]) < TOL and rel(o["l_null"], g["r64_l_null"]) < 1e-9
    assert rel(o["lambda_ml"], g["r64_lambda_alt"]).max() < TOL
    assert rel(o["loglik_ml"], g["r64_l_alt"]).max() < 1e-9
    assert np.abs(2 * (o["loglik_ml"] - o["l_null"]) - g["r64_D_lrt"]).max() < 1e-7


def test_traw_reader_and_std_filter(tmp_path):
    """.traw text ingest (host) and the callers' `X.std(axis=0) > 0` filter (experiments/wtccc/run_pygemma.py:407-410)
    computed from packed genotype counts: round trip, NA handling, gz, ragged n, both NaN policies against NumPy."""
    from pygemma_b200 import bed, traw

    rng = np.random.default_rng(3)
    n, m = 37, 60   # n % 4 != 0
    G = rng.binomial(2, rng.uniform(0.05, 0.5, m), size=(n, m)).astype(np.float64)
    G[rng.random(G.shape) < 0.03] = np.nan
    G[:, 4] = 1.0                       # monomorphic
    G[:, 9] = np.nan                    # all missing
    G[:, 11] = 2.0; G[3, 11] = np.nan   # monomorphic among the observed calls
    G[:, 20] = rng.integers(0, 3, n)    # complete, polymorphic
    for name in ("toy.traw", "toy.traw.gz"):
        path = str(tmp_path / name)
        traw.write_traw(path, G, snp_ids=[f"snp{i}" for i in range(m)])
        packed, n2, info, samples = traw.read_traw(path, chunk_snps=17)
        assert n2 == n and packed.shape == (m, (n + 3) // 4) and len(samples) == n
        assert list(info.columns) == traw.INFO_COLUMNS and list(info["SNP"][:2]) == ["snp0", "snp1"]
        D = bed.decode_packed(packed, n)
        assert np.array_equal(D, G, equal_nan=True)
    c = traw.genotype_counts(packed, n)
    assert (c.sum(axis=1) == n).all() and c[9, 1] == n and c[4, 2] == n
    with np.errstate(invalid="ignore"):
        ref_prop = np.nan_to_num(G.std(axis=0), nan=0.0) > 0           # NaN > 0 is False
        ref_omit = np.nan_to_num(np.nanstd(np.where(np.isnan(G).all(0), 0.0, G), axis=0), nan=0.0) > 0
    assert np.array_equal(traw.std_filter(packed, n), ref_prop)
    assert np.array_equal(traw.std_filter(packed, n, nan_policy="omit"), ref_omit)
    assert not traw.std_filter(packed, n, "omit")[[4, 9, 11]].any() and traw.std_filter(packed, n)[20]
    bad = str(tmp_path / "bad.traw")
    open(bad, "w").write("CHR\tSNP\t(C)M\tPOS\tCOUNTED\tALT\ta_a\tb_b\n1\trs1\t0\t1\tA\tG\t0.5\t1\n")
    with pytest.raises(ValueError, match="0, 1, 2 or NA"):
        traw.read_traw(bad)
