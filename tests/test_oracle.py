"""The CPU oracle (oracle/reml_oracle.c) against the reference's own outputs.

Golden vectors in tests/golden/ were produced by the reference's compiled
Cython code (ref64 = float64-promoted build; tests/golden/make_golden.py).
"""
import glob
import os

import numpy as np
import pytest

from conftest import COLS, GOLDEN
from oracle import oracle


def rel(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


SCANS = sorted(glob.glob(os.path.join(GOLDEN, "scan_*.npz")))


@pytest.mark.parametrize("path", SCANS, ids=[os.path.basename(p)[5:-4] for p in SCANS])
def test_scan_matches_ref64(path):
    g = np.load(path)
    o = oracle.scan_rotated(g["d"], g["yr"], g["wr"], np.ascontiguousarray(g["xr"].T), grid=bool(g["grid"]))
    degenerate = "degenerate" in path
    for c in COLS:
        ref = g[f"r64_{c}"]
        assert np.array_equal(np.isnan(o[c]), np.isnan(ref)), c
        if degenerate:
            # collinear / constant genotype columns: the reference returns finite garbage governed by
            # its 1e-35 clamps; only well-posed columns are value-checked
            ok = np.array([0, 2, 4, 5, 6, 7])
            assert rel(o[c][ok], ref[ok]).max() < 1e-9, c
        else:
            tol = 1e-6 if c == "lambda" else 1e-8  # flat likelihoods: lambda itself is only determined to ~1e-7
            assert rel(o[c], ref).max() < tol, (c, rel(o[c], ref).max())
    # the literal fp32 reference agrees to its own noise floor (reported, loose)
    if not degenerate and "rough" not in path:
        assert np.median(rel(o["p_wald"], g["r32_p_wald"])) < 1e-3


def test_e2e_matches_ref64():
    for path in sorted(glob.glob(os.path.join(GOLDEN, "e2e_*.npz"))):
        g = np.load(path)
        o = oracle.pygemma(g["Y"], g["X"], g["W"], g["K"], grid=bool(g["grid"]))
        for c in COLS:
            assert rel(o[c], g[f"r64_{c}"]).max() < 1e-8, (path, c)


def test_precompute_kat():
    """The reference test-suite's own seeded matrices and lambda probes (tests/test_pygemma.py:195-253)."""
    g = np.load(os.path.join(GOLDEN, "kat_precompute.npz"))
    d, wr, xr, yr = (g[k].astype(np.float64) for k in ("d", "Wr", "xr", "Yr"))
    n, c0 = wr.shape
    cf = c0 + 1
    for li, lam in enumerate(g["lams"]):
        lev, scal = oracle.precompute_probe(lam, d, wr, xr, yr, full=True)
        for j, k in enumerate(["yt_Pi_y", "yt_Pi_Pi_y", "yt_Pi_Pi_Pi_y", "tr_Pi", "tr_Pi_Pi"]):
            ref = g[f"r64_full_{k}_{li}"]
            assert rel(lev[:, j], ref).max() < 1e-7, (lam, k, rel(lev[:, j], ref).max())
        s = g[f"r64_scal_{li}"]
        assert rel(scal[5], s[0]) < 1e-10 and rel(scal[6], s[1]) < 1e-9
        assert rel(scal[7], s[4]) < 1e-8 and rel(scal[8], s[5]) < 1e-6
        L = oracle.lib()
        d1 = L.pgo_d1(lam, n, cf, scal[0], scal[1], scal[3])
        d2 = L.pgo_d2(lam, n, cf, scal[0], scal[1], scal[2], scal[3], scal[4])
        ll = L.pgo_loglik(n, cf, scal[0], scal[5], scal[6])
        assert abs(d1 - s[6]) <= 1e-6 * max(abs(s[6]), abs(0.5 * (n - cf) / lam) * 1e-3), (lam, d1, s[6])
        assert rel(d2, s[7]) < 1e-5, (lam, d2, s[7])
        assert rel(ll, s[8]) < 1e-10
        lev2, scal2 = oracle.precompute_probe(lam, d, wr, xr, yr, full=False)
        assert rel(lev2[:, 0], g[f"r64_half_yt_Pi_y_{li}"]).max() < 1e-7
        assert rel(lev2[:, 3], g[f"r64_half_tr_Pi_{li}"]).max() < 1e-10
        w = g[f"r64_wald_{li}"]
        beta = scal2[8] / scal2[7]
        se = np.sqrt(scal2[0]) / (np.sqrt(max(scal2[7], 1e-35)) * np.sqrt(n - c0 - 1))
        tau = (n - c0 - 1) / scal2[0]
        assert rel(beta, w[0]) < 1e-6 and rel(se, w[1]) < 1e-8 and rel(tau, w[2]) < 1e-8
    o = oracle.scan_rotated(d, yr, wr, xr[None, :])
    og = oracle.scan_rotated(d, yr, wr, xr[None, :], grid=True)
    assert o["lambda"][0] == g["r64_calc_lambda"][0] and og["lambda"][0] == g["r64_calc_lambda"][1]


def test_brentq_is_scipy_brentq():
    """pgo_brentq restates scipy.optimize.brentq: identical call sequences and roots."""
    from scipy import optimize

    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(10000):
        p = rng.standard_normal(5) * rng.choice([0.0, 1.0], size=5, p=[0.3, 0.7])
        a = 10.0 ** rng.integers(-5, 5)
        b = a * 10.0
        f = lambda x: oracle.poly_eval(p, x)
        fa, fb = f(a), f(b)
        if fa == 0 or fb == 0 or np.signbit(fa) == np.signbit(fb):
            continue
        rtol = rng.choice([0.1, 1e-3, 4 * np.finfo(float).eps])
        xs = []

        def g(x):
            xs.append(x)
            return f(x)

        r_sp = optimize.brentq(g, a, b, rtol=rtol, maxiter=100, disp=False)
        r, nc, seq = oracle.brentq_probe(p, a, b, rtol=rtol)
        assert nc == len(xs)
        assert np.array_equal(seq, np.array(xs))
        assert r == r_sp
        checked += 1
    assert checked > 500


def test_oracle_eval_counts_match_survey():
    """~26 precompute_mat calls per SNP in the default mode, 13 in grid mode."""
    g = np.load(os.path.join(GOLDEN, "scan_interior.npz"))
    xt = np.ascontiguousarray(g["xr"].T)
    o = oracle.scan_rotated(g["d"], g["yr"], g["wr"], xt)
    tot = o["n_eval2"] + o["n_eval3"]
    assert 20 <= tot.mean() <= 32
    og = oracle.scan_rotated(g["d"], g["yr"], g["wr"], xt, grid=True)
    assert (og["n_eval2"] == 13).all() and (og["n_eval3"] == 0).all()


def test_oracle_stays_inside_the_float32_reference_noise():
    """ref32 (the literal float32 reference) vs the float64 oracle on the whole-call goldens: the documented noise of the
    reference's float32 storage bounds the deviation; p-values rank identically."""
    import glob
    import os

    from scipy import stats

    from conftest import COLS, GOLDEN
    from oracle import oracle

    bounds = {"beta": 5e-3, "se_beta": 1e-4, "tau": 1e-4, "lambda": 1e-3, "F_wald": 1e-2, "p_wald": 2e-3}
    for path in sorted(glob.glob(os.path.join(GOLDEN, "e2e_*.npz"))):
        g = np.load(path)
        o = oracle.pygemma(g["Y"], g["X"], g["W"], g["K"], grid=bool(g["grid"]))
        for c in COLS:
            e = np.abs(o[c] - g[f"r32_{c}"]) / np.maximum(np.abs(g[f"r32_{c}"]), 1e-300)
            assert e.max() < bounds[c], (path, c, float(e.max()))
        assert stats.spearmanr(o["p_wald"], g["r32_p_wald"]).statistic > 0.9999


def test_oracle_de_mode_matches_reference_cpdefs():
    """de=True (lmm/lmm.py:498-532): the oracle with swapped roles against the reference's cpdefs called in calculate_de's
    roles (tests/golden/de_small.npz, made by make_golden.py de)."""
    import os

    from conftest import COLS, GOLDEN
    from oracle import oracle

    g = np.load(os.path.join(GOLDEN, "de_small.npz"))
    m = g["xr"].shape[1]
    yrow = np.ascontiguousarray(g["yr"].reshape(1, -1))
    for j in range(m):
        r = oracle.scan_rotated(g["d"], g["xr"][:, j], g["wr"], yrow)
        for c in COLS:
            assert rel(r[c][0], g[f"r64_{c}"][j]) < (1e-7 if c == "lambda" else 1e-9), (j, c)


@pytest.mark.parametrize("name", ["lrt_interior", "lrt_low_h2", "lrt_c8"])
def test_oracle_lrt_matches_reference_functions(name):
    """LRT scaffolding (lmm/lmm.py:176-190,:278-300): the oracle's dense restatement against the reference's live
    lmm.calc_lambda / likelihood_lambda (tests/golden/lrt_*.npz, made by make_golden.py lrt)."""
    import os

    from conftest import GOLDEN
    from oracle import oracle

    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    o = oracle.lrt_rotated(g["d"], g["yr"], g["wr"], np.ascontiguousarray(g["xr"].T))
    assert rel(o["lambda_null"], g["r64_lambda_null"]) < 1e-7 and rel(o["l_null"], g["r64_l_null"]) < 1e-10
    assert rel(o["lambda_ml"], g["r64_lambda_alt"]).max() < 1e-6
    assert rel(o["loglik_ml"], g["r64_l_alt"]).max() < 1e-10
    assert np.abs(o["D_lrt"] - g["r64_D_lrt"]).max() < 1e-8
    ok = g["r64_p_lrt"] > 1e-9   # the reference's 1 - cdf loses the tail
    assert rel(o["p_lrt"][ok], g["r64_p_lrt"][ok]).max() < 1e-6
