#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own compiled code.

Run in the BUILD container only (needs /root/reference -> oracle/_ref via
oracle/build_ref.py).  The .npz files are committed; the GPU box and the CPU
test-suite only read them.

  ref64  = reference sources with float32 promoted to float64 (the 1e-6 gate)
  ref32  = the literal reference (reported alongside, looser)

Fixtures
  kat_precompute.npz   the reference test-suite's own seeded generator
                       (tests/test_pygemma.py:195-212, n=1000, covars=10, seed=42)
                       replayed, with precompute_mat / derivative / newton /
                       calc_lambda / Wald outputs at its five lambda probes
                       {1e-3, 5, 400, 1e3, 1e5} (tests/test_pygemma.py:253).
  scan_<name>.npz      rotated-space scans through lmm.calculate (lmm/lmm.py:461)
  e2e_<name>.npz       whole lmm.pygemma calls (lmm/lmm.py:87), eigh included
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore")

COLS = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]


def _reference_generate_test_matrices(n=1000, covars=10, seed=42):
    """Replay of the reference's tests/test_pygemma.py:195-212 random stream (legacy np.random.seed)."""
    np.random.seed(seed)
    K = np.random.uniform(size=(n, n))
    K = np.abs(np.tril(K) + np.tril(K, -1).T)
    K = np.dot(K, K.T)
    eigenVals, U = np.linalg.eig(K)
    eigenVals = np.maximum(0, eigenVals)
    W = np.random.rand(n, covars)
    W = np.c_[W, np.ones(n)]
    x = np.random.choice([0, 1, 2], size=(n, 1), replace=True)
    Y = np.random.rand(n, 1).reshape(-1, 1)
    _ = np.random.rand(covars + 2, 1)
    f32 = np.float32
    return x.astype(f32), Y.astype(f32), W.astype(f32), np.real(eigenVals).astype(f32), np.real(U).astype(f32)


def _scan(lmm, d, yr, wr, xr, grid, dt):
    with np.errstate(all="ignore"):
        rows = lmm.calculate((d.astype(dt), yr.astype(dt).reshape(-1, 1), np.ascontiguousarray(wr.astype(dt)),
                              np.ascontiguousarray(xr.astype(dt)), grid))
    return {c: np.array([float(r[c]) for r in rows]) for c in COLS}


def make_kat(lmm32, lmm64):
    x, Y, W, d, U = _reference_generate_test_matrices()
    xr = (U.T @ x).reshape(-1)
    Yr = U.T @ Y
    Wr = U.T @ W
    out = {"d": d, "Wr": Wr, "xr": xr, "Yr": Yr.reshape(-1), "lams": np.array([1e-3, 5.0, 400, 1e3, 1e5])}
    keys = ["yt_Pi_y", "yt_Pi_Pi_y", "yt_Pi_Pi_Pi_y", "tr_Pi", "tr_Pi_Pi"]
    for tag, lmm, dt in (("r64", lmm64, np.float64), ("r32", lmm32, np.float32)):
        dd, ww, yy = d.astype(dt), np.c_[Wr, xr].astype(dt), Yr.astype(dt)
        n, c = ww.shape
        for li, lam in enumerate(out["lams"]):
            with np.errstate(all="ignore"):
                pf = lmm.precompute_mat(lam, dd, ww, yy, True)
                pn = lmm.precompute_mat(lam, dd, ww, yy, False)
            for k in keys:
                out[f"{tag}_full_{k}_{li}"] = np.asarray(pf[k], dtype=np.float64)
            for k in keys[:2] + ["tr_Pi"]:
                out[f"{tag}_half_{k}_{li}"] = np.asarray(pn[k], dtype=np.float64)
            out[f"{tag}_scal_{li}"] = np.array([
                pf["logdet_H"], pf["logdet_Wt_H_inv_W"], pn["logdet_H"], pn["logdet_Wt_H_inv_W"],
                float(pn["wjt_Pi_wk"][c - 1, c - 1, c - 1]), float(pn["wjt_Pi_wk"][c, c - 1, c - 1]),
                float(lmm.likelihood_derivative1_restricted_lambda_overload(lam, n, c, pf["yt_Pi_y"][c], pf["yt_Pi_Pi_y"][c], pf["tr_Pi"][c])),
                float(lmm.likelihood_derivative2_restricted_lambda_overload(lam, n, c, pf["yt_Pi_y"][c], pf["yt_Pi_Pi_y"][c], pf["yt_Pi_Pi_Pi_y"][c], pf["tr_Pi"][c], pf["tr_Pi_Pi"][c])),
                float(lmm.likelihood_restricted_lambda_overload(lam, n, c, pf["yt_Pi_y"][c], pf["logdet_H"], pf["logdet_Wt_W"], pf["logdet_Wt_H_inv_W"])),
                float(lmm.wrapper_likelihood_derivative1_restricted_lambda(lam, dd, yy, ww)),
                float(lmm.newton(lam, dd, yy, ww, True)),
            ])
            b, _, se, tau = lmm.calc_beta_vg_ve_restricted_overload(dd, Wr.astype(dt), xr.astype(dt).reshape(-1, 1), dt(lam), yy)
            out[f"{tag}_wald_{li}"] = np.array([float(b), float(se), float(tau)])
        out[f"{tag}_calc_lambda"] = np.array([float(lmm.calc_lambda_restricted(dd, yy, ww)),
                                              float(lmm.calc_lambda_restricted(dd, yy, ww, True, True))])
    np.savez_compressed(os.path.join(HERE, "kat_precompute.npz"), **out)
    print("kat_precompute", out["r64_calc_lambda"], out["r32_calc_lambda"])


def _rot_problem(n, m, c0, seed, h2, m_k):
    from pygemma_b200.synth import make_problem
    from oracle import oracle

    p = make_problem(n, m, c0, seed=seed, h2=h2, m_k=m_k)
    d, U, yr, xr, wr = oracle.eigen_rotate(p["K"], p["Y"], p["X"], p["W"])
    return d, yr.reshape(-1), wr, xr


def make_scans(lmm32, lmm64):
    rng = np.random.default_rng(99)
    cases = {}
    # (name, n, m, c0, seed, h2, m_k, grid)
    spec = [
        ("interior", 400, 48, 4, 11, 0.5, 120, False),
        ("boundary_lo", 300, 48, 3, 12, 0.0, 100, False),
        ("high_h2", 300, 48, 2, 13, 1.0, 100, False),
        ("intercept_only", 256, 40, 1, 14, 0.4, 80, False),
        ("grid", 320, 48, 6, 15, 0.5, 100, True),
        ("grid_lo", 200, 32, 3, 16, 0.0, 60, True),
        ("c12", 500, 32, 12, 17, 0.6, 150, False),
    ]
    for name, n, m, c0, seed, h2, m_k, grid in spec:
        d, yr, wr, xr = _rot_problem(n, m, c0, seed, h2, m_k)
        cases[name] = (d, yr, wr, xr, grid)
    # rough likelihood surfaces: random spectrum + random y -> several sign changes / odd brackets
    for t in range(3):
        n, m, c0 = 48 + 16 * t, 96, 1 + t
        d = np.sort(rng.lognormal(0.0, 3.0, size=n))
        wr = rng.standard_normal((n, c0))
        yr = rng.standard_normal(n) * np.sqrt(rng.choice([1e-3, 1.0, 1e3], size=n) * d + 1.0)
        xr = rng.standard_normal((n, m))
        cases[f"rough{t}"] = (d, yr, wr, xr, False)
    # degenerate genotype columns: constant, duplicate of a covariate, all-zero
    d, yr, wr, xr = _rot_problem(200, 8, 3, 21, 0.5, 60)
    xr = xr.copy()
    xr[:, 1] = wr[:, 0] * 2.0
    xr[:, 2] = 0.0
    xr[:, 3] = wr[:, 1] - 0.5 * wr[:, 2]
    cases["degenerate"] = (d, yr, wr, xr, False)

    for name, (d, yr, wr, xr, grid) in cases.items():
        r64 = _scan(lmm64, d, yr, wr, xr, grid, np.float64)
        r32 = _scan(lmm32, d, yr, wr, xr, grid, np.float32)
        out = {"d": d, "yr": yr, "wr": wr, "xr": xr, "grid": np.array(grid)}
        out.update({f"r64_{c}": r64[c] for c in COLS})
        out.update({f"r32_{c}": r32[c] for c in COLS})
        np.savez_compressed(os.path.join(HERE, f"scan_{name}.npz"), **out)
        lam = r64["lambda"]
        print(f"scan_{name}: n={d.shape[0]} m={xr.shape[1]} c0={wr.shape[1]} lam[min,med,max]="
              f"{np.nanmin(lam):.3g},{np.nanmedian(lam):.3g},{np.nanmax(lam):.3g} "
              f"lo={np.mean(lam == 1e-5):.2f} hi={np.mean(lam == 1e5):.2f} nan={np.isnan(lam).sum()}")


def make_e2e(lmm32, lmm64):
    from pygemma_b200.synth import make_problem

    for name, n, m, c0, seed, h2, grid in [("small", 128, 40, 3, 31, 0.5, False), ("grid", 96, 32, 2, 32, 0.3, True),
                                            ("wide", 160, 64, 5, 33, 0.7, False)]:
        p = make_problem(n, m, c0, seed=seed, h2=h2, m_k=2 * n)
        snps = np.array([f"rs{i}" for i in range(m)])
        with np.errstate(all="ignore"):
            r64 = lmm64.pygemma(p["Y"], p["X"].astype(np.float64), p["W"], p["K"], snps=snps, grid=grid)
            r32 = lmm32.pygemma(p["Y"], p["X"].astype(np.float32), p["W"], p["K"], snps=snps, grid=grid)
        assert list(r64.columns) == COLS + ["SNPs"], r64.columns
        out = {"Y": p["Y"], "X": p["X"], "W": p["W"], "K": p["K"], "grid": np.array(grid), "snps": snps,
               "columns": np.array(list(r64.columns))}
        out.update({f"r64_{c}": r64[c].values.astype(np.float64) for c in COLS})
        out.update({f"r32_{c}": r32[c].values.astype(np.float64) for c in COLS})
        np.savez_compressed(os.path.join(HERE, f"e2e_{name}.npz"), **out)
        print(f"e2e_{name}: lambda med {np.median(r64['lambda']):.4g}")


def make_de(lmm32, lmm64):
    """de=True: calculate_de (lmm/lmm.py:498-532) cannot be run as written -- its driver passes SampleIter's 5-tuple
    (lmm/lmm.py:499 vs :434) and it hands calc_lambda_restricted a 1-D X[:, g] where the cpdef demands (n, 1) (:504,
    pyx:61).  The goldens are made with the two cpdefs it calls, in its roles, with that one reshape:
        lambda = calc_lambda_restricted(d, X[:, g].reshape(-1, 1), np.c_[W, Y])
        beta, _, se, tau = calc_beta_vg_ve_restricted_overload(d, W, Y, lambda, X[:, g].reshape(-1, 1))
    and F_wald / p_wald as at lmm/lmm.py:507-519."""
    from scipy import stats

    from pygemma_b200.synth import make_problem
    from oracle import oracle

    n, m, c0 = 300, 40, 3
    p = make_problem(n, m, c0, seed=51, h2=0.5, m_k=100)
    rng = np.random.default_rng(7)
    pred = p["X"][:, 4].astype(np.float64)
    E = rng.standard_normal((n, m)) + 0.2 * pred[:, None] * (np.arange(m) % 2 == 0) + 0.4 * p["Y"]
    d, U, yr, xr, wr = oracle.eigen_rotate(p["K"], pred, E, p["W"])
    out = {"d": d, "yr": yr.reshape(-1), "wr": wr, "xr": xr}
    for tag, lmm, dt in (("r64", lmm64, np.float64), ("r32", lmm32, np.float32)):
        dd, Y, W, X = d.astype(dt), yr.astype(dt).reshape(-1, 1), np.ascontiguousarray(wr.astype(dt)), np.ascontiguousarray(xr.astype(dt))
        rows = {c: [] for c in COLS}
        with np.errstate(all="ignore"):
            for g in range(m):
                xg = np.ascontiguousarray(X[:, g].reshape(-1, 1))
                lam = lmm.calc_lambda_restricted(dd, xg, np.c_[W, Y])
                beta, _, se, tau = lmm.calc_beta_vg_ve_restricted_overload(dd, W, Y, lam, xg)
                F = np.float64(beta / se) ** 2.0
                for c, v in zip(COLS, (beta, se, tau, lam, F, stats.f.sf(x=F, dfn=1, dfd=n - c0 - 1))):
                    rows[c].append(float(v))
        for c in COLS:
            out[f"{tag}_{c}"] = np.array(rows[c])
    np.savez_compressed(os.path.join(HERE, "de_small.npz"), **out)
    print("de_small: lambda med", np.median(out["r64_lambda"]), "beta[:4]", out["r64_beta"][:4])


def make_lrt(lmm32, lmm64):
    """Likelihood-ratio outputs (commented out in the reference driver, lmm/lmm.py:137-141,:176-190,:278-300): the live
    functions those lines call -- lmm.calc_lambda (ML lambda: lmm/lmm.py:22-84), likelihood_lambda (pyx:1542) -- run on
    rotated inputs for the null model [W] and for every alternative [W, x_g]:
        D_lrt = 2 (l_alt - l_null),  p_lrt = 1 - chi2.cdf(D_lrt, 1)          (lmm/lmm.py:282,:300)
    (likelihood(lam, tau = n / yPy, beta_hat, ...) of lmm/lmm.py:189,:281 equals likelihood_lambda(lam, ...) identically.)"""
    from scipy import stats

    cases = [("lrt_interior", 300, 40, 3, 61, 0.5, 100), ("lrt_low_h2", 260, 32, 2, 62, 0.05, 80), ("lrt_c8", 400, 24, 8, 63, 0.7, 150)]
    for name, n, m, c0, seed, h2, m_k in cases:
        d, yr, wr, xr = _rot_problem(n, m, c0, seed, h2, m_k)
        out = {"d": d, "yr": yr, "wr": wr, "xr": xr}
        for tag, lmm, dt in (("r64", lmm64, np.float64), ("r32", lmm32, np.float32)):
            dd, Y, W, X = d.astype(dt), yr.astype(dt).reshape(-1, 1), np.ascontiguousarray(wr.astype(dt)), xr.astype(dt)
            with np.errstate(all="ignore"):
                lam0 = lmm.calc_lambda(dd, Y, W)
                l0 = float(lmm.likelihood_lambda(lam0, dd, Y, W))
                lam_a, l_a = [], []
                for g in range(m):
                    Wx = np.ascontiguousarray(np.c_[W, X[:, g]])
                    la = lmm.calc_lambda(dd, Y, Wx)
                    lam_a.append(float(la))
                    l_a.append(float(lmm.likelihood_lambda(la, dd, Y, Wx)))
            lam_a, l_a = np.array(lam_a), np.array(l_a)
            D = 2.0 * (l_a - l0)
            out.update({f"{tag}_lambda_null": np.array(float(lam0)), f"{tag}_l_null": np.array(l0), f"{tag}_lambda_alt": lam_a,
                        f"{tag}_l_alt": l_a, f"{tag}_D_lrt": D, f"{tag}_p_lrt": 1.0 - stats.chi2.cdf(D, 1)})
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "lambda_null", out["r64_lambda_null"], "l_null", out["r64_l_null"], "D[:4]", out["r64_D_lrt"][:4],
              "lam_alt[min,max]", out["r64_lambda_alt"].min(), out["r64_lambda_alt"].max())


if __name__ == "__main__":
    from oracle import build_ref

    assert build_ref.build(), "reference build unavailable"
    from pygemma import lmm as lmm32
    from pygemma64 import lmm as lmm64

    only = sys.argv[1:]   # e.g. `make_golden.py de` regenerates one family
    for name, fn in (("kat", make_kat), ("scans", make_scans), ("e2e", make_e2e), ("de", make_de), ("lrt", make_lrt)):
        if not only or name in only:
            fn(lmm32, lmm64)
