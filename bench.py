#!/usr/bin/env python
"""bench.py -- SNPs/sec of the per-SNP LMM association scan (BASELINE.json metric: n = 10 000 samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5]

A "step" is one pass of the hot path (rotation U^T X + per-SNP REML/Wald scan) over one batch of synthetic SNPs.
Default workload = BASELINE.json configs[2] ("c3": n = 10 000 x 100 000 SNPs, c0 = 10, Brent+Newton, int8 dosages);
the other configs of BASELINE.json are selectable (c1 mouse shape, c2 GD449 shape, c4 n = 50 000 streamed from host,
c5 grid + c0 = 40).  The one-time eigendecomposition (cuSOLVER syevd) is setup: timed separately, not in the metric.

  value       whole-job SNPs/s with the genotypes already resident in HBM (pg_scan_device), CUDA events, max over ranks
  e2e         the same metric through the public call `pygemma_b200.lmm.pygemma(Y, X, W, lmm.factorize(K))` with a
              PAGEABLE NumPy genotype matrix, timed to the returned DataFrame (uploads, downloads, gather inside)
  e2e_pinned  the C-ABI host-buffer call pg_scan fed from pinned memory (what round 1 reported as e2e)
  roofline    dominant kernel: achieved algorithmic op/s over the measured peak
  parity_spot >= 32 SNPs of the timed problem (per rank) against the CPU oracle fed with the device's own U, d
  cpu_baseline / --impl reference: the reference's own Cython path (oracle/_ref, built from /root/reference by
              oracle/build_ref.py) on this box's host cores on a bounded sample, in both modes of SURVEY 8d:
              (i) Pool(cores) x 1 BLAS thread, (ii) one process x BLAS threads = cores; rotation with all cores

N > 1 (torchrun, one process per GPU): ONE fixed problem (strong scaling): rank 0 decomposes, U and d are broadcast once
over NCCL (setup, timed separately), every rank scans its contiguous SNP shard and the per-SNP results are all-gathered
inside the timed step; rank 0 compares the gathered rows bit for bit with its own single-GPU scan of the whole problem.
The weak-scaling figure (every rank scans the full batch, no collective) is reported alongside under "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SNPs/sec at n=10k samples"
COLS = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]

# BASELINE.json configs (index 0..4 <-> c1..c5); m is the whole problem
CONFIGS = {
    "c1": dict(n=1940, m=12226, c0=11, grid=False, note="mouse_hs1940 shape (BASELINE.json configs[0]), synthetic genotypes"),
    "c2": dict(n=449, m=100000, c0=6, grid=False, note="1000G GD449 shape (BASELINE.json configs[1]), synthetic genotypes"),
    "c3": dict(n=10000, m=100000, c0=10, grid=False, note="synthetic UKB shape (BASELINE.json configs[2])"),
    "c4": dict(n=50000, m=1000000, c0=10, grid=False, note="n=50 000 x 1 000 000 SNPs, SNP-sharded, genotypes streamed "
                                                            "from host memory (BASELINE.json configs[3])"),
    "c5": dict(n=10000, m=100000, c0=40, grid=True, note="grid-search lambda + c0=40 (BASELINE.json configs[4])"),
}


# ------------------------------------------------------------------------------------------------
# reference / CPU-baseline worker (separate interpreter: no CUDA context, BLAS threads set by the parent)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(path: str) -> None:
    """Times the reference's scan (lmm.calculate through multiprocessing.Pool, reference lmm/lmm.py:378-401) and / or
    its fp32 rotation (lmm/lmm.py:244) on the sample stored in `path`, with the BLAS thread count of this process's
    environment."""
    import multiprocessing
    import warnings

    warnings.filterwarnings("ignore")
    z = np.load(path)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    kind = "reference"
    try:
        from pygemma import lmm as ref  # literal reference build (fp32 storage)
    except Exception:  # pragma: no cover - the reference did not build in this container
        ref = None
        kind = "port"
    d, y, w, x = z["d"], z["y"], z["w"], z["x"]  # rotated, x is (n, m_s)
    steps, warmup, nproc, legs = int(z["steps"]), int(z["warmup"]), int(z["nproc"]), str(z["legs"])
    n, m_s = x.shape
    times, rot_times = [], []
    if "scan" in legs:
        if ref is not None:
            f32 = np.float32
            d32, y32, w32, x32 = d.astype(f32), y.astype(f32).reshape(-1, 1), np.ascontiguousarray(w.astype(f32)), \
                np.ascontiguousarray(x.astype(f32))
            with multiprocessing.Pool(nproc) as pool:
                for it in range(warmup + steps):
                    t0 = time.perf_counter()
                    rows = []
                    with np.errstate(all="ignore"):
                        for r in pool.imap(ref.calculate, ref.SampleIter(x32, y32, w32, d32, bool(z["grid"]), nproc)):
                            rows = rows + r
                    dt = time.perf_counter() - t0
                    assert len(rows) == m_s
                    if it >= warmup:
                        times.append(dt)
        else:
            from oracle import oracle

            for it in range(warmup + steps):
                t0 = time.perf_counter()
                oracle.scan_rotated(d, y, w, np.ascontiguousarray(x.T), grid=bool(z["grid"]), fast=True)
                dt = time.perf_counter() - t0
                if it >= warmup:
                    times.append(dt)
    m_rot = 0
    if "rot" in legs:
        # U.T @ X as the reference does it (float32 sgemm, lmm/lmm.py:244) with this process's BLAS threads
        rng = np.random.default_rng(1)
        m_rot = int(z["m_rot"])
        u = rng.standard_normal((n, n), dtype=np.float32)
        xs = rng.integers(0, 3, size=(n, m_rot)).astype(np.float32)
        for it in range(3):
            t0 = time.perf_counter()
            _ = u.T @ xs
            rot_times.append(time.perf_counter() - t0)
    print(json.dumps({"kind": kind, "nproc": nproc, "m_sample": int(m_s), "scan_s": float(np.mean(times)) if times else None,
                      "rot_s_per_snp": (float(np.min(rot_times)) / m_rot) if rot_times else None, "m_rot": m_rot}))


def _run_worker(path, blas_threads):
    env = dict(os.environ, OPENBLAS_NUM_THREADS=str(blas_threads), OMP_NUM_THREADS=str(blas_threads),
               MKL_NUM_THREADS=str(blas_threads), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu_worker", path], env=env, capture_output=True,
                       text=True, timeout=3000)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError(f"cpu reference worker failed: {r.stdout[-2000:]} {r.stderr[-2000:]}")
    return json.loads(lines[-1])


def run_cpu_reference(n, c0, grid, steps, warmup, seed=5):
    """Bounded rotated-space sample of the workload, timed with the reference in fresh interpreters:
    mode (i)  multiprocessing.Pool(cores) with one BLAS thread per process (the reference's multi-process mode),
    mode (ii) one process with BLAS threads = cores,
    rotation  float32 U.T @ X with BLAS threads = cores (the reference's launcher sets OPENBLAS_NUM_THREADS = cores,
              experiments/benchmarks/benchmark.sh:25)."""
    from pygemma_b200.synth import make_spectral_problem

    cores = os.cpu_count() or 1
    m_i = max(64, 24 * cores)                 # mode (i): 24 SNPs per process
    m_ii = max(16, min(48, 3 * cores))        # mode (ii): one process, a few dozen SNPs
    p = make_spectral_problem(n, m_i, c0, seed=seed, xdtype=np.float64)
    rot_leg = n <= 20000                      # a 50 000^2 float32 U is 10 GB of host noise: scan only there
    with tempfile.TemporaryDirectory() as td:
        def dump(name, m_s, nproc, legs):
            path = os.path.join(td, name)
            np.savez(path, d=p["d"], y=p["Y"].reshape(-1), w=p["W"], x=p["X"][:, :m_s].astype(np.float32), grid=np.array(grid),
                     steps=steps, warmup=warmup, nproc=nproc, legs=np.array(legs), m_rot=2048)
            return path
        r_i = _run_worker(dump("i.npz", m_i, cores, "scan"), 1)
        r_ii = _run_worker(dump("ii.npz", m_ii, 1, "scan+rot" if rot_leg else "scan"), cores)
    rot = r_ii["rot_s_per_snp"] or 0.0
    sps_i = 1.0 / (r_i["scan_s"] / r_i["m_sample"] + rot)
    sps_ii = 1.0 / (r_ii["scan_s"] / r_ii["m_sample"] + rot)
    best = max(sps_i, sps_ii)
    mode = "i" if sps_i >= sps_ii else "ii"
    m_best = r_i["m_sample"] if mode == "i" else r_ii["m_sample"]
    return {
        "kind": r_i["kind"], "cores": cores, "snps_per_s": best, "mode": mode, "m_sample": m_best,
        "ms_per_step": 1e3 * m_best / best,
        "modes": {"i_pool_cores_x_1_blas_thread": {"snps_per_s": sps_i, "scan_only_snps_per_s": r_i["m_sample"] / r_i["scan_s"],
                                                     "m_sample": r_i["m_sample"]},
                  "ii_1_process_x_blas_cores": {"snps_per_s": sps_ii, "scan_only_snps_per_s": r_ii["m_sample"] / r_ii["scan_s"],
                                                  "m_sample": r_ii["m_sample"]}},
        "rotation_s_per_snp": rot,
        "rotation_gflops": (2.0 * n * n / rot / 1e9) if rot else None,
        "sample": (f"{m_best} synthetic SNPs at n={n}, c0={c0}, {'grid' if grid else 'Brent+Newton'} mode; reference "
                   f"lmm.calculate, best of mode (i) multiprocessing.Pool({cores}) x 1 BLAS thread and mode (ii) 1 process x "
                   f"{cores} BLAS threads (mode {mode} reported); "
                   + (f"plus the float32 U.T@X rotation timed on 2048 SNPs with {cores} BLAS threads; "
                      if rot_leg else "rotation leg skipped at this n (scan only: overstates the reference); ")
                   + f"mean of {steps} runs after {warmup} warm-ups"),
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "source_hbm": "fallback"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        peaks["hbm_gbs"] = float(j.get("hbm_gbs", peaks["hbm_gbs"]))
        peaks["source_hbm"] = "measured (MEASURED_PEAKS.json)"
    q = os.path.join(ROOT, "profiles", "peaks_fp64_int8.json")
    if os.path.exists(q):
        j = json.load(open(q))
        peaks.update({k: float(v) for k, v in j.items() if isinstance(v, (int, float))})
        peaks["source_fp64"] = "measured on this pool by tools/peak_fp64 (profiles/peaks_fp64_int8.json)"
    else:
        peaks.update({"fp64_fma_tflops": 34.0, "fp64_dmma_tflops": 37.0, "dgemm_tflops": 36.0, "int8_gemm_tops": 3600.0})
        peaks["source_fp64"] = "fallback"
    # the rotation hands B over transposed: use the better of the two measured cuBLAS int8 rates as the denominator
    peaks["int8_gemm_tops"] = max(peaks.get("int8_gemm_tops", 0.0), peaks.get("int8_gemm_tt_tops", 0.0))
    return peaks


NCU_FILES = {  # committed `ncu --set full` summaries (tools/ncu_summary.py), newest first; SNPs per captured launch
    "rotate_i8_tc2": [("ncu_r02_tc2_persist_25088snps.json", 25088), ("ncu_r02_tc2_16384snps.json", 16384),
                      ("ncu_r01_tc2_final_16384snps.json", 16384)],
    # moments fused into the rotation (c0 = 10, 206 nodes: +23 % tiles; the c5 shape has +54 %): an order of magnitude only
    "rotate_i8_tc2_kernel<1, 1>": [("ncu_r02_tc2_fused_16384snps.json", 16384)],
    "compress_dmma": [("ncu_r02_reml_8192snps.json", 8192), ("ncu_r01_reml_final_8192snps.json", 8192)],
    "reml_solve": [("ncu_r02_reml_8192snps.json", 8192), ("ncu_r01_reml_final_8192snps.json", 8192)],
    "cutlass": [("ncu_r01_hot_kernels_8192snps.json", 3584)],
    "combine_i8": [("ncu_r01_hot_kernels_8192snps.json", 3584)],
}


def ncu_traffic(kernel_substr: str):
    """DRAM bytes per SNP of a kernel from the committed `ncu --set full` captures under profiles/ (n = 10 000)."""
    for name, snps in NCU_FILES.get(kernel_substr, []):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        for e in json.load(open(p)):
            if kernel_substr in e["kernel"] and "dram_traffic_bytes" in e:
                return e["dram_traffic_bytes"] / snps, name
    return None, None


def make_gpu_problem(torch, dev, n, m, c0, seed, x_seed, with_k=True):
    """Synthetic UKB-shape inputs generated on the device: binomial dosages (int8, sample-major), kinship from an
    independent standardised panel (+1e-3 I), intercept + gaussian covariates, polygenic phenotype.  Every rank that
    calls this with the same seeds gets the same arrays (same device type, same generator stream)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    K = None
    if with_k:
        m_k = 2 * n
        maf = torch.rand(m_k, generator=g, device=dev) * 0.45 + 0.05
        G = ((torch.rand(n, m_k, generator=g, device=dev) < maf).to(torch.float64)
             + (torch.rand(n, m_k, generator=g, device=dev) < maf).to(torch.float64))
        sd = G.std(dim=0)
        sd[sd == 0] = 1.0
        G = (G - G.mean(dim=0)) / sd
        K = G @ G.T / m_k
        K.diagonal().add_(1e-3)
        b = torch.randn(m_k, generator=g, device=dev, dtype=torch.float64)
        u = G @ b / (m_k ** 0.5)
        del G
    else:
        u = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    u = u / u.std()
    W = torch.cat([torch.ones(n, 1, device=dev, dtype=torch.float64),
                   torch.randn(n, max(c0 - 1, 0), generator=g, device=dev, dtype=torch.float64)], dim=1)[:, :max(c0, 1)]
    y = (0.5 ** 0.5) * u + (0.5 ** 0.5) * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    if c0 > 1:
        y = y + 0.05 * W[:, 1:].sum(dim=1)
    X = make_gpu_genotypes(torch, dev, n, m, x_seed)
    return K, W, y, X


def make_gpu_genotypes(torch, dev, n, m, x_seed):
    gx = torch.Generator(device=dev)
    gx.manual_seed(x_seed)
    X = torch.empty((n, m), dtype=torch.int8, device=dev)
    step = max(256, (1 << 26) // n // 256 * 256)
    for a in range(0, m, step):
        bnd = min(m, a + step)
        mf = torch.rand(bnd - a, generator=gx, device=dev) * 0.45 + 0.05
        X[:, a:bnd] = ((torch.rand(n, bnd - a, generator=gx, device=dev) < mf).to(torch.int8)
                       + (torch.rand(n, bnd - a, generator=gx, device=dev) < mf).to(torch.int8))
    return X


def parity_spot(torch, h, dev, n, W_host, y_host, X_cols_dev, got, grid):
    """>= 32 SNPs of the timed problem against the CPU oracle (oracle/reml_oracle.c), fed with the device's own
    eigen-system: U, d are read back from the handle, U^T y / U^T W / U^T x formed in FP64 (torch matmul, independent of
    the int8 rotation kernel), and the oracle runs the reference's per-SNP algorithm on them.  Returns the worst relative
    deviation over the six result columns."""
    from oracle import oracle

    U = torch.empty(n * n, dtype=torch.float64, device=dev)
    d = torch.empty(n, dtype=torch.float64, device=dev)
    h.get_eigen_device(U.data_ptr(), d.data_ptr())
    Ut = U.view(n, n)  # column-major U: row i of this view is eigenvector i, i.e. the view is U^T
    yr = (Ut @ torch.from_numpy(y_host).to(dev)).cpu().numpy()
    wr = (Ut @ torch.from_numpy(W_host).to(dev)).cpu().numpy()
    xr = (Ut @ X_cols_dev.to(torch.float64)).T.contiguous().cpu().numpy()
    d_host = d.cpu().numpy()
    del U, Ut
    ref = oracle.scan_rotated(d_host, yr, wr, xr, grid=grid)
    worst, per_col = 0.0, {}
    for c in COLS:
        a, b = np.asarray(got[c], float), np.asarray(ref[c], float)
        if not np.array_equal(np.isnan(a), np.isnan(b)):
            per_col[c] = float("inf")
            worst = float("inf")
            continue
        ok = ~np.isnan(b)
        e = float((np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)).max()) if ok.any() else 0.0
        per_col[c] = e
        worst = max(worst, e)
    return worst, per_col


def run_ours(args):
    import pandas as pd
    import torch
    import torch.distributed as dist

    from pygemma_b200 import _capi, lmm, multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pygemma_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    cfg = dict(CONFIGS[args.config])
    n, c0, grid = cfg["n"], cfg["c0"], cfg["grid"]
    if args.n:
        n = args.n
    if args.c0 is not None:
        c0 = args.c0
    if args.grid:
        grid = True
    m = args.snps or cfg["m"]
    c4 = args.config == "c4"
    if c4 and not args.snps:
        m = cfg["m"] if world > 1 else 125000   # one GPU: one rank's share of the 8-GPU problem
    ctx = multi.context(local)
    a_sh, b_sh = ctx.shard(m)
    per = multi.shard_len(m, world)
    m_loc = b_sh - a_sh
    stream = torch.cuda.current_stream(dev)
    seed = 20240 + 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- setup (excluded from the metric): eigen-system on rank 0, NCCL broadcast of U and d
    t0 = time.perf_counter()
    if c4:
        # forming K and running syevd at n = 50 000 is not benchmark material: a random orthogonal U (QR on the device)
        # and a chi-square-like spectrum enter through the pg_set_eigen_device door (the eigen=False-style entry)
        _, W, y, _ = make_gpu_problem(torch, dev, n, 0, c0, seed, seed, with_k=False)
        U_t = torch.empty(n * n, dtype=torch.float64, device=dev)
        d_t = torch.empty(n, dtype=torch.float64, device=dev)
        if rank == 0:
            g = torch.Generator(device=dev)
            g.manual_seed(seed + 7)
            A = torch.randn(n, n, generator=g, device=dev, dtype=torch.float64)
            Q, _ = torch.linalg.qr(A)
            del A
            U_t.copy_(Q.t().contiguous().view(-1))   # column-major U
            del Q
            d_t.copy_((torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 3 + 0.05).sort().values)
            torch.cuda.empty_cache()
        h = _capi.Handle(n, c0, local)
        multi.setup_eigen_from(ctx, h, U_t, d_t)
        del U_t, d_t
        torch.cuda.empty_cache()
        factor, eig_ms = None, 0.0
    else:
        K, W, y, X = make_gpu_problem(torch, dev, n, m, c0, seed, seed * 1000 + 17)
        K_host = K.cpu().numpy() if rank == 0 else None
        del K
        factor = lmm.factorize(K_host, device=local, c0=c0)
        del K_host
        h, eig_ms = factor.handle(c0), factor.eig_ms
    setup_s = time.perf_counter() - t0
    bcast_ms = 1e3 * multi.last_collective_s["broadcast"] if world > 1 else 0.0
    h.set_stream(stream.cuda_stream)
    rot_opt = {"auto": _capi.PG_ROT_AUTO, "fp64": _capi.PG_ROT_FP64, "i8split": _capi.PG_ROT_I8SPLIT,
               "i8tc": _capi.PG_ROT_I8TC}[args.rotation]
    h.set_options(rotation=rot_opt)
    W_host, y_host = W.cpu().numpy(), y.cpu().numpy()
    design_ms = h.set_design(W_host, y_host)

    def timed(fn, steps, warmup, sample_clocks=False, tag="pg_timed"):
        for _ in range(warmup):
            fn()
        sampler = ClockSampler(local) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        tms = []
        torch.cuda.nvtx.range_push(tag)  # lets `ncu --nvtx --nvtx-include "<tag>/"` list exactly the timed launches
        for _ in range(steps):
            tms.append(fn())
        torch.cuda.nvtx.range_pop()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        dev_ms = e0.elapsed_time(e1)
        ms = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]), tms, clocks

    line_extra = {}
    if c4:
        # ---- C4: the rank's shard lives in (pageable) host memory and is streamed through pg_scan's pinned double buffers
        X_np = np.empty((n, m_loc), dtype=np.int8)
        chunk = 8192
        gx_seed = seed * 1000 + 17 + rank
        for a in range(0, m_loc, chunk):
            bnd = min(m_loc, a + chunk)
            X_np[:, a:bnd] = make_gpu_genotypes(torch, dev, n, bnd - a, gx_seed + a).cpu().numpy()
        torch.cuda.empty_cache()
        res_holder = {}

        def step_stream():
            t1 = time.perf_counter()
            r = h.scan(X_np, grid=grid)
            res_holder["res"] = r
            if world > 1:
                res_holder["all"] = multi.gather_results(ctx, r, m)
            tm = dict(r["timing"])
            tm["wall_ms"] = 1e3 * (time.perf_counter() - t1)
            return tm

        dev_ms, wall_ms, tms, clocks = timed(step_stream, args.steps, args.warmup, sample_clocks=True, tag="pg_timed_stream")
        ms_per_step = wall_ms / args.steps   # host-synchronous call: wall time is the step (device time is inside it)
        value = m / (ms_per_step * 1e-3)
        res_tm = tms[-1]
        e2e = {"value": value, "unit": "SNPs/s", "h2d_bytes_per_step": int(n) * int(m), "d2h_bytes_per_step": int(m) * 60,
               "ms_per_step": ms_per_step, "api": "pg_scan (C ABI) on pageable host genotypes of every rank's shard, pinned "
               "double-buffered uploads inside the library, + multi.gather_results", "copies_declared": True,
               "h2d_gbs_per_gpu": n * m_loc / (res_tm["wall_ms"] * 1e-3) / 1e9}
        res = res_holder["res"]
        got_cols = {c: res[c] for c in COLS}
        st_counts = np.stack([res["status"], res["n_eval2"], res["n_eval3"]])
        idx = np.unique(np.linspace(0, m_loc - 1, 8 if n > 20000 else 32).astype(np.int64))
        X_spot = torch.from_numpy(np.ascontiguousarray(X_np[:, idx])).to(dev)
        line_extra["value_note"] = ("C4 has no HBM-resident variant: 50 GB of genotypes are streamed from host memory, so value "
                                    "and e2e are the same host-streamed measurement")
        e2e_pinned = None
        weak = None
        check = None
        gather_ms = 1e3 * multi.last_collective_s["gather"] if world > 1 else 0.0
    else:
        # ---- resident: every rank scans its shard of the ONE device-resident problem, results all-gathered on the device
        out_dev = torch.full((6, per), float("nan"), dtype=torch.float64, device=dev)
        st_dev = torch.zeros((3, per), dtype=torch.int32, device=dev)
        out_ptrs = [out_dev[i].data_ptr() for i in range(6)]
        X_shard_ptr = X.data_ptr() + a_sh   # sample-major int8: column a_sh of row 0; ld = m
        gath = {}
        g_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        gather_ms_list = []

        def step_resident():
            tm = h.scan_device(X_shard_ptr, _capi.PG_X_I8, m, _capi.PG_X_SAMPLE_MAJOR, m_loc, grid, out_ptrs,
                               st_dev[0].data_ptr(), st_dev[1].data_ptr(), st_dev[2].data_ptr())
            if world > 1:
                g_ev[0].record(stream)
                gath["out"], gath["st"] = multi.gather_device(ctx, out_dev, st_dev)
                g_ev[1].record(stream)
                g_ev[1].synchronize()
                gather_ms_list.append(g_ev[0].elapsed_time(g_ev[1]))
            return tm

        dev_ms, wall_ms, tms, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True,
                                             tag="pg_timed_resident")
        res_tm = tms[-1]
        ms_per_step = dev_ms / args.steps
        value = m / (ms_per_step * 1e-3)
        gather_ms = float(np.mean(gather_ms_list[-args.steps:])) if gather_ms_list else 0.0
        got_dev = out_dev[:, :m_loc].cpu().numpy()
        got_cols = {c: got_dev[i] for i, c in enumerate(COLS)}
        st_counts = st_dev[:, :m_loc].cpu().numpy()

        # ---- multi-GPU check: the gathered table on rank 0 against rank 0's own single-GPU scan of the whole problem
        check = None
        if world > 1:
            allo = gath["out"].view(world, 6, per).cpu().numpy()
            alls = gath["st"].view(world, 3, per).cpu().numpy()
            if rank == 0:
                full_out = torch.empty((6, m), dtype=torch.float64, device=dev)
                full_st = torch.zeros((3, m), dtype=torch.int32, device=dev)
                h.scan_device(X.data_ptr(), _capi.PG_X_I8, m, _capi.PG_X_SAMPLE_MAJOR, m, grid,
                              [full_out[i].data_ptr() for i in range(6)], full_st[0].data_ptr(), full_st[1].data_ptr(),
                              full_st[2].data_ptr())
                fo, fs = full_out.cpu().numpy(), full_st.cpu().numpy()
                same = True
                for r in range(world):
                    ra, rb = multi.shard_range(m, r, world)
                    same = same and np.array_equal(allo[r][:, :rb - ra], fo[:, ra:rb], equal_nan=True) \
                        and np.array_equal(alls[r][:, :rb - ra], fs[:, ra:rb])
                check = {"bit_identical_to_1gpu": bool(same), "rows": int(m), "ranks": world,
                         "how": "rank 0 re-scans all SNPs alone and compares the all-gathered 6 result columns + status / "
                                "evaluation counts bit for bit"}
                del full_out, full_st

        # ---- weak scaling (second metric): every rank scans the FULL batch concurrently, no collective
        weak = None
        if world > 1:
            wo = torch.empty((6, m), dtype=torch.float64, device=dev)
            ws = torch.zeros((3, m), dtype=torch.int32, device=dev)
            wp = [wo[i].data_ptr() for i in range(6)]

            def step_weak():
                return h.scan_device(X.data_ptr(), _capi.PG_X_I8, m, _capi.PG_X_SAMPLE_MAJOR, m, grid, wp,
                                     ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr())
            w_dev, _, _, _ = timed(step_weak, max(2, args.steps // 2), 1, tag="pg_timed_weak")
            w_ms = w_dev / max(2, args.steps // 2)
            weak = {"value": world * m / (w_ms * 1e-3), "unit": "SNPs/s", "ms_per_step": w_ms, "snps_per_gpu": m,
                    "scaling": "weak", "note": "every rank scans its own copy of the full batch; no data-path collective"}
            del wo, ws

        # ---- e2e: the public call, pageable NumPy genotypes, to the returned DataFrame
        X_np = X.cpu().numpy()   # pageable host copy of the whole problem (every rank holds it, like the reference's callers)
        y_col = y_host.reshape(-1, 1)
        df_holder = {}

        def step_e2e():
            df_holder["df"] = lmm.pygemma(y_col, X_np, W_host, factor, grid=grid)
            return dict(lmm.last_timing)

        e_dev, e_wall, e_tms, _ = timed(step_e2e, args.steps, max(1, args.warmup - 1), tag="pg_timed_e2e")
        e2e_ms = e_wall / args.steps
        df = df_holder["df"]
        assert isinstance(df, pd.DataFrame) and len(df) == m and list(df.columns) == COLS
        lt = e_tms[-1]
        e2e = {"value": m / (e2e_ms * 1e-3), "unit": "SNPs/s",
               "h2d_bytes_per_step": int(n) * int(m) + world * 8 * int(n) * (c0 + 1),
               "d2h_bytes_per_step": int(m) * (6 * 8 + 3 * 4) * (world if world > 1 else 1), "ms_per_step": e2e_ms,
               "api": "pygemma_b200.lmm.pygemma(Y, X, W, lmm.factorize(K)) -> pandas.DataFrame; X is a pageable NumPy int8 "
                      "array; design upload, genotype upload (pinned bounce buffers inside pg_scan), result download, "
                      "all-gather and DataFrame assembly are inside the timed region; eigh(K) is not (lmm.factorize, setup)",
               "copies_declared": True,
               "last_call": {"design_ms": lt.get("design_ms"), "scan_wall_s": lt.get("scan_wall_s"), "gather_s": lt.get("gather_s"),
                             "total_s": lt.get("total_s"),
                             "scan": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in lt.get("scan", {}).items()}}}
        # the DataFrame of the public call must be the resident result, bit for bit (rank 0's shard checked on each rank)
        for i, c in enumerate(COLS):
            if not np.array_equal(df[c].to_numpy()[a_sh:b_sh], got_dev[i], equal_nan=True):
                e2e["mismatch_vs_resident"] = c

        # ---- e2e_pinned: the C-ABI host call fed from pinned memory (upper bound of the host path)
        X_pin = torch.empty((n, m_loc), dtype=torch.int8, pin_memory=True)
        X_pin.copy_(X[:, a_sh:b_sh])
        X_pin_np = X_pin.numpy()

        def step_pinned():
            return h.scan(X_pin_np, grid=grid)

        _, p_wall, p_tms, _ = timed(step_pinned, args.steps, 1, tag="pg_timed_pinned")
        e2e_pinned = {"value": m / (p_wall / args.steps * 1e-3), "unit": "SNPs/s", "ms_per_step": p_wall / args.steps,
                      "api": "pg_scan (C ABI) on pinned host genotypes, per-rank shard, no gather",
                      "timing_last_call": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in p_tms[-1]["timing"].items()}}
        idx = np.unique(np.linspace(0, m_loc - 1, 32).astype(np.int64))
        X_spot = X[:, a_sh:b_sh][:, torch.from_numpy(idx).to(dev)]

    # ---- parity spot check against the CPU oracle (every rank, its own shard)
    spot = {"snps_checked": 0, "max_rel": None}
    try:
        worst, per_col = parity_spot(torch, h, dev, n, W_host, y_host, X_spot, {c: got_cols[c][idx] for c in COLS}, grid)
        w_t = torch.tensor([worst if np.isfinite(worst) else 1e300], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w_t, op=dist.ReduceOp.MAX)
        spot = {"snps_checked": int(len(idx)) * world, "max_rel": float(w_t[0]), "tolerance": 1e-6,
                "pass": bool(float(w_t[0]) < 1e-6), "rank0_per_column": per_col,
                "how": "oracle/reml_oracle.c (reference algorithm, FP64) on U^T x formed in FP64 from the handle's own U, d; "
                       "SNPs evenly spaced over every rank's shard of the timed problem"}
    except Exception as e:  # the checker must never take the measurement down
        spot = {"snps_checked": 0, "max_rel": None, "error": str(e)[:300]}

    bad = int((st_counts[0] != 0).sum())
    ev2, ev3 = float(st_counts[1].mean()), float(st_counts[2].mean())
    peaks = load_peaks()
    # ---- roofline of the dominant kernel (per launch = per SNP block; times are CUDA-event sums of the last step)
    rot_ms, reml_ms, cmp_ms, conv_ms = res_tm["rotate_ms"], res_tm["reml_ms"], res_tm["compress_ms"], res_tm["convert_ms"]
    solve_ms = reml_ms - cmp_ms
    step_ms = max(res_tm["total_ms"], 1e-9)
    nodes = res_tm["n_nodes"]
    i8 = res_tm.get("rot_engine") in (_capi.PG_ROT_I8SPLIT, _capi.PG_ROT_I8TC, _capi.PG_ROT_I8TC_MOMENTS)
    fused = res_tm.get("rot_engine") in (_capi.PG_ROT_I8TC, _capi.PG_ROT_I8TC_MOMENTS)
    # PG_FUSE_MOMENTS=1 (opt-in, pg_set_moment_fusion): the rotation kernel also contracts against the columns of G = U V
    moments_fused = res_tm.get("rot_engine") == _capi.PG_ROT_I8TC_MOMENTS
    g_cols = h.fusion_info()["g_columns"] if moments_fused else 0
    n_planes = int(_capi.load().pg_rotation_planes())   # 7 unless the library was built with -DPG_SLICES=6
    rot_ops = (n_planes if i8 else 1) * 2.0 * n * (n + g_cols)   # int8 (or fp64) multiply-add ops per SNP
    cmp_flops = 2.0 * n * 10 * (c0 + 2)                       # compression: n x kCq x (c0+2) FP64 FMAs per SNP
    stages = {
        "rotation": {"ms": rot_ms, "achieved": rot_ops * m_loc / (max(rot_ms, 1e-9) * 1e-3) / 1e12,
                     "peak": peaks["int8_gemm_tops"] if i8 else peaks["fp64_dmma_tflops"],
                     "kernel": (f"rotation U^T X, exact int8-split ({n_planes} base-256 digit planes): rotate_i8_tc2_kernel (hand-written TMA + "
                                "tcgen05 cta_group::2 kind::i8, recombination fused)") if (i8 and fused) else
                               (f"rotation U^T X, exact int8-split ({n_planes} base-256 digit planes): cuBLAS int8 GEMM (cutlass3x sm100 "
                                "tcgen05 2-SM kernel) + combine_i8_kernel") if i8 else
                               "rotation U^T X (cuBLAS DGEMM, FP64 tensor pipe)",
                     "peak_source": ("int8 tensor rate (cuBLAS int8 GEMM 16384x8192x8192) " if i8 else "FP64 DMMA rate ")
                                    + peaks["source_fp64"],
                     "algorithmic_ops_per_snp": rot_ops},
        "compress": {"ms": cmp_ms, "achieved": cmp_flops * m_loc / (max(cmp_ms, 1e-9) * 1e-3) / 1e12,
                     "peak": peaks["fp64_dmma_tflops"], "kernel": "compress_dmma_kernel (FP64 tensor pipe, mma.sync m8n8k4)",
                     "peak_source": "FP64 DMMA rate " + peaks["source_fp64"], "algorithmic_ops_per_snp": cmp_flops},
    }
    dom = max(stages, key=lambda k: stages[k]["ms"])
    if solve_ms > stages[dom]["ms"]:
        # the optimiser on the compressed moments: FP64 FMA work = passes x nodes x (c0+2) x powers
        sol_flops = 2.0 * nodes * (c0 + 2) * (2 * ev2 + 3 * ev3)
        stages["solve"] = {"ms": solve_ms, "achieved": sol_flops * m_loc / (solve_ms * 1e-3) / 1e12, "peak": peaks["fp64_fma_tflops"],
                           "kernel": "reml_solve_kernel", "peak_source": "FP64 FMA rate " + peaks["source_fp64"],
                           "algorithmic_ops_per_snp": sol_flops}
        dom = "solve"
    st = stages[dom]
    roofline = {"kernel": st["kernel"], "bound": "tensor" if dom != "solve" else "fp64", "achieved": st["achieved"],
                "peak": st["peak"], "unit": "TFLOP/s", "frac": st["achieved"] / st["peak"], "traffic": None,
                "peak_source": st["peak_source"], "algorithmic_ops_per_snp": st["algorithmic_ops_per_snp"],
                "ops": "int8 multiply-add ops counted 2 per MAC" if (dom == "rotation" and i8) else "fp64 flops",
                "share_of_step": st["ms"] / step_ms}
    per_snp, src = None, None
    if dom == "rotation" and moments_fused:
        per_snp, src = ncu_traffic("rotate_i8_tc2_kernel<1, 1>")
    elif dom == "rotation" and i8 and fused:
        per_snp, src = ncu_traffic("rotate_i8_tc2")
    elif dom == "rotation" and i8:
        (a_, s1), (b_, _) = ncu_traffic("cutlass"), ncu_traffic("combine_i8")
        per_snp, src = ((a_ + b_) if (a_ and b_) else None), s1
    elif dom == "compress":
        per_snp, src = ncu_traffic("compress_dmma")
    elif dom == "solve":
        per_snp, src = ncu_traffic("reml_solve")
    roofline["traffic"] = per_snp * res_tm["block_snps"] if (per_snp and n == 10000) else None
    roofline["traffic_note"] = (f"dram__bytes_read+write of the stage's kernels per SNP block (ncu at n=10000, profiles/{src}); "
                                "algorithmic HBM bytes per SNP for the fused rotation: n int8 in + 8n fp64 out (+ the digit planes "
                                "once per eigen-tile group)")
    roofline["algorithmic_bytes_per_launch"] = (9.0 * n) * res_tm["block_snps"] if dom == "rotation" else None
    roofline["per_kernel_ms_last_step"] = {"convert": conv_ms, "rotate": rot_ms, "compress": cmp_ms, "solve": solve_ms}
    roofline["rotate_tflops_fp64_equiv"] = 2.0 * n * n * m_loc / (max(rot_ms, 1e-9) * 1e-3) / 1e12
    roofline["compress_frac_of_dmma_peak"] = stages["compress"]["achieved"] / stages["compress"]["peak"]
    roofline["nodes_per_snp"] = nodes

    wl = (f"{cfg['note']}: n={n} samples x {m} SNPs in total, c0={c0} covariates, "
          f"{'grid-search' if grid else 'Brent+Newton'} lambda, int8 dosages")
    line = {
        "metric": METRIC, "value": value, "unit": "SNPs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": wl, "config": args.config, "n": n, "snps_total": m, "snps_per_gpu": m_loc, "c0": c0, "grid": grid,
                   "parallelism": (f"snp-shard x{world}: U, d broadcast once (NCCL), contiguous SNP shards, per-SNP results "
                                   f"all-gathered (NCCL) inside the timed step") if world > 1 else "snp-shard x1",
                   "l2": "inputs_larger_than_l2 (genotype block + 8n B/SNP rotated fp64 per step exceed the 126 MB L2)",
                   "reml_engine": "compressed (eigenvalue-space moments)",
                   "rotation_engine": (("int8-split fused tcgen05 + moments (pg_set_moment_fusion)" if moments_fused else
                                        "int8-split fused tcgen05") if fused else "int8-split cuBLAS") if i8 else "fp64"},
        "e2e": e2e, "e2e_pinned": e2e_pinned,
        "gpu_launches": int(sum(t["convert_launches"] + t["reml_launches"] + t["rotate_launches"] for t in tms)),
        "clocks": clocks, "roofline": roofline, "parity_spot": spot,
        "setup": {"syevd_ms": eig_ms, "eigen_setup_s": setup_s, "design_tables_ms": design_ms},
        "evals_per_snp": {"two_power_passes": ev2, "three_power_passes": ev3}, "nan_rows": bad,
    }
    line.update(line_extra)
    if world > 1:
        line["collective_ms"] = {"broadcast_U_d_ms_setup": bcast_ms, "gather_ms_per_step": gather_ms,
                                 "collective": "ncclBroadcast of U (8 n^2 B) and d once per decomposition; ncclAllGather of "
                                               "6 float64 + 3 int32 per SNP at the end of every step",
                                 "share_of_step": gather_ms / max(ms_per_step, 1e-9)}
        line["multi_gpu_check"] = check
        line["weak"] = weak
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = run_cpu_reference(n, c0, grid, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": cb["snps_per_s"], "unit": "SNPs/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"], "modes": cb["modes"], "rotation_gflops": cb["rotation_gflops"]}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "SNPs/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e}"}
    if factor is not None:
        factor.close()
    else:
        h.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CONFIGS[args.config])
    n = args.n or cfg["n"]
    c0 = cfg["c0"] if args.c0 is None else args.c0
    grid = bool(args.grid or cfg["grid"])
    cb = run_cpu_reference(n, c0, grid, steps=args.steps, warmup=args.warmup)
    line = {
        "metric": METRIC, "value": cb["snps_per_s"], "unit": "SNPs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64 (reference mix)", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{cfg['note']}: n={n} samples, c0={c0} covariates, "
                               f"{'grid-search' if grid else 'Brent+Newton'} lambda; bounded sample of "
                               f"{cb['m_sample']} SNPs per step", "config": args.config,
                   "n": n, "c0": c0, "grid": grid},
        "cpu_baseline": {"value": cb["snps_per_s"], "unit": "SNPs/s", "cores": cb["cores"], "kind": cb["kind"],
                         "sample": cb["sample"], "modes": cb["modes"], "rotation_gflops": cb["rotation_gflops"]},
        "e2e": {"value": cb["snps_per_s"], "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json config (default c3 = configs[2])")
    ap.add_argument("--snps", type=int, default=0, help="override the config's total SNP count")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--c0", type=int, default=None)
    ap.add_argument("--grid", action="store_true")
    ap.add_argument("--rotation", default="auto", choices=["auto", "fp64", "i8split", "i8tc"],
                    help="rotation engine (auto = library default; i8tc = hand-written fused tcgen05 kernel)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--_cpu_worker", default=None)
    args = ap.parse_args()
    if args._cpu_worker:
        _cpu_worker(args._cpu_worker)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
