#!/usr/bin/env python
"""bench.py -- SNPs/sec of the per-SNP LMM association scan at n = 10 000 samples (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--snps M] [--grid]

A "step" is one pass of the hot path (rotation U^T X + per-SNP REML/Wald scan) over one batch of
M synthetic SNPs (default: BASELINE.json configs[2], n = 10 000, M = 100 000, c0 = 10 covariates).
The one-time eigendecomposition (cuSOLVER syevd) is setup: timed separately, excluded from the metric.

  value     whole-job SNPs/s with the genotypes already resident in HBM (pg_scan_device), CUDA events
  e2e       the same metric through the host-buffer C-ABI call pg_scan (what lmm.pygemma calls): pinned
            host genotypes in, host result arrays out, copies inside the timed region
  roofline  dominant kernel: achieved algorithmic FLOP/s over the measured peak
  cpu_baseline / --impl reference: the reference's own Cython path (oracle/_ref, built from
            /root/reference by oracle/build_ref.py) timed on this box's host cores on a bounded sample

N > 1: one process per GPU (torchrun); U and d are broadcast once over NCCL, every rank scans its own
M SNPs (weak scaling), no collective in the data path; time is the max over ranks.
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

N_SAMPLES = 10000
C0 = 10
METRIC = "SNPs/sec at n=10k samples"


# ------------------------------------------------------------------------------------------------
# reference / CPU-baseline worker (separate interpreter: no CUDA context, BLAS threads pinned to 1)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(path: str) -> None:
    """Times the reference's scan (lmm.calculate through multiprocessing.Pool, reference lmm/lmm.py:378-401)
    and its fp32 rotation (lmm/lmm.py:244) on the sample stored in `path`."""
    import multiprocessing
    import warnings

    warnings.filterwarnings("ignore")
    z = np.load(path)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    kind = "reference"
    try:
        from pygemma import lmm as ref  # literal reference build (fp32 storage)
    except Exception as e:  # pragma: no cover - the reference did not build in this container
        ref = None
        kind = "port"
        err = str(e)
    d, y, w, x = z["d"], z["y"], z["w"], z["x"]  # rotated, x is (n, m_s)
    steps, warmup, nproc = int(z["steps"]), int(z["warmup"]), int(z["nproc"])
    n, m_s = x.shape
    times = []
    rot_times = []
    u = z["u"] if "u" in z.files else None
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
        if u is not None:
            xs = np.ascontiguousarray(z["xraw"].astype(f32))
            for it in range(3):
                t0 = time.perf_counter()
                _ = u.T @ xs
                rot_times.append(time.perf_counter() - t0)
    else:
        from oracle import oracle

        for it in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.scan_rotated(d, y, w, np.ascontiguousarray(x.T), grid=bool(z["grid"]), fast=True)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    t_scan = float(np.mean(times))
    t_rot = float(np.min(rot_times)) * (m_s / z["xraw"].shape[1]) if rot_times else 0.0
    print(json.dumps({"kind": kind, "cores": nproc, "m_sample": int(m_s), "scan_s": t_scan, "rotate_s": t_rot,
                      "snps_per_s": m_s / (t_scan + t_rot), "snps_per_s_scan_only": m_s / t_scan,
                      "ms_per_step": 1e3 * (t_scan + t_rot)}))


def run_cpu_reference(n, c0, grid, steps, warmup, m_sample=None, seed=5):
    """Builds a rotated-space synthetic sample and times the reference on it in a fresh interpreter."""
    from pygemma_b200.synth import make_spectral_problem

    cores = os.cpu_count() or 1
    nproc = max(1, cores)
    if m_sample is None:
        m_sample = max(64, 24 * nproc)
    p = make_spectral_problem(n, m_sample, c0, seed=seed, xdtype=np.float64)
    rng = np.random.default_rng(seed)
    m_rot = min(m_sample, 256)
    u = rng.standard_normal((n, n), dtype=np.float32)  # timing of the fp32 sgemm only
    xraw = rng.integers(0, 3, size=(n, m_rot)).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.npz")
        np.savez(path, d=p["d"], y=p["Y"].reshape(-1), w=p["W"], x=p["X"].astype(np.float32), u=u, xraw=xraw, grid=np.array(grid),
                 steps=steps, warmup=warmup, nproc=nproc)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1",
                   CUDA_VISIBLE_DEVICES="")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu_worker", path], env=env,
                           capture_output=True, text=True, timeout=3000)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError(f"cpu reference worker failed: {r.stdout[-2000:]} {r.stderr[-2000:]}")
    out = json.loads(lines[-1])
    out["sample"] = (f"{out['m_sample']} synthetic SNPs at n={n}, c0={c0}, {'grid' if grid else 'Brent+Newton'} mode, "
                     f"reference lmm.calculate over multiprocessing.Pool({nproc}) with 1 BLAS thread per process "
                     f"+ fp32 U.T@X rotation on all cores; mean of {steps} runs after {warmup} warm-ups")
    return out


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


def ncu_traffic(kernel_substr: str):
    """DRAM bytes per SNP of a kernel from the committed `ncu --set full` captures (cuBLAS engine:
    profiles/ncu_r01_hot_kernels_8192snps.json, tools/prof_rot.py 10000 8192 10, first launch covers 3584 SNPs)."""
    p = os.path.join(ROOT, "profiles", "ncu_r01_hot_kernels_8192snps.json")
    snps = {"cutlass": 3584, "combine_i8": 3584, "compress_dmma": 8192, "reml_solve": 8192, "rotate_i8_tc2": 16384}
    # final-code captures (tools/run_ncu_tc2.sh): prof_tc.py 10000 16384 and prof_reml.py 10000 8192 10
    if kernel_substr == "rotate_i8_tc2":
        p = os.path.join(ROOT, "profiles", "ncu_r01_tc2_final_16384snps.json")
    elif kernel_substr in ("compress_dmma", "reml_solve"):
        p = os.path.join(ROOT, "profiles", "ncu_r01_reml_final_8192snps.json")
    if not os.path.exists(p):
        return None
    for e in json.load(open(p)):
        if kernel_substr in e["kernel"] and "dram_traffic_bytes" in e:
            return e["dram_traffic_bytes"] / snps[kernel_substr]
    return None


def make_gpu_problem(torch, dev, n, m, c0, seed, rank):
    """Synthetic UKB-shape inputs generated on the device: binomial dosages (int8, sample-major), kinship
    from an independent standardised panel (+1e-3 I), intercept + gaussian covariates, polygenic phenotype."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
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
    u = u / u.std()
    W = torch.cat([torch.ones(n, 1, device=dev, dtype=torch.float64),
                   torch.randn(n, c0 - 1, generator=g, device=dev, dtype=torch.float64)], dim=1)
    y = (0.5 ** 0.5) * u + (0.5 ** 0.5) * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    y = y + 0.05 * W[:, 1:].sum(dim=1)
    del G
    # genotypes of this rank (different SNPs per rank: weak scaling)
    gx = torch.Generator(device=dev)
    gx.manual_seed(seed * 1000 + 17 + rank)
    X = torch.empty((n, m), dtype=torch.int8, device=dev)
    step = 8192
    for a in range(0, m, step):
        bnd = min(m, a + step)
        mf = torch.rand(bnd - a, generator=gx, device=dev) * 0.45 + 0.05
        X[:, a:bnd] = ((torch.rand(n, bnd - a, generator=gx, device=dev) < mf).to(torch.int8)
                       + (torch.rand(n, bnd - a, generator=gx, device=dev) < mf).to(torch.int8))
    return K, W, y, X


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pygemma_b200 import _capi, multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pygemma_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    n, m, c0, grid = args.n, args.snps, args.c0, bool(args.grid)
    ctx = multi.context(local)

    K, W, y, X = make_gpu_problem(torch, dev, n, m, c0, seed=20240 + 2, rank=rank)
    h = _capi.Handle(n, c0, local)
    stream = torch.cuda.current_stream(dev)
    h.set_stream(stream.cuda_stream)
    rot_opt = {"auto": _capi.PG_ROT_AUTO, "fp64": _capi.PG_ROT_FP64, "i8split": _capi.PG_ROT_I8SPLIT,
               "i8tc": _capi.PG_ROT_I8TC}[args.rotation]
    h.set_options(rotation=rot_opt)

    # ---- setup (excluded from the metric): eigendecomposition on rank 0, NCCL broadcast of U and d
    t0 = time.perf_counter()
    K_host = K.cpu().numpy() if rank == 0 else None
    del K
    eig_ms = multi.setup_eigen(ctx, h, K_host)
    del K_host
    setup_s = time.perf_counter() - t0
    W_host, y_host = W.cpu().numpy(), y.cpu().numpy()
    design_ms = h.set_design(W_host, y_host)

    out_dev = torch.empty((6, m), dtype=torch.float64, device=dev)
    st_dev = torch.zeros((3, m), dtype=torch.int32, device=dev)
    out_ptrs = [out_dev[i].data_ptr() for i in range(6)]

    def step_resident():
        return h.scan_device(X.data_ptr(), _capi.PG_X_I8, m, _capi.PG_X_SAMPLE_MAJOR, m, grid, out_ptrs,
                             st_dev[0].data_ptr(), st_dev[1].data_ptr(), st_dev[2].data_ptr())

    X_host = torch.empty((n, m), dtype=torch.int8, pin_memory=True)
    X_host.copy_(X)
    X_np = X_host.numpy()

    def step_e2e():
        return h.scan(X_np, grid=grid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

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

    dev_ms, wall_ms, tms, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True, tag="pg_timed_resident")
    res_tm = tms[-1]
    ms_per_step = dev_ms / args.steps
    value = world * m / (ms_per_step * 1e-3)
    e2e_dev_ms, e2e_wall_ms, e2e_tms, _ = timed(step_e2e, args.steps, max(1, args.warmup - 1), tag="pg_timed_e2e")
    e2e_ms = e2e_wall_ms / args.steps  # host-visible time of the synchronous call (>= device time)
    e2e_value = world * m / (e2e_ms * 1e-3)
    counts = st_dev.cpu().numpy()
    bad = int((counts[0] != 0).sum())
    ev2, ev3 = float(counts[1].mean()), float(counts[2].mean())

    peaks = load_peaks()
    # ---- roofline of the dominant kernel (per launch = per SNP block; times are CUDA-event sums of the last step)
    rot_ms, reml_ms, cmp_ms, conv_ms = res_tm["rotate_ms"], res_tm["reml_ms"], res_tm["compress_ms"], res_tm["convert_ms"]
    solve_ms = reml_ms - cmp_ms
    step_ms = max(res_tm["total_ms"], 1e-9)  # stage spans may overlap (optimiser of block b under the rotation of b+1)
    nodes = res_tm["n_nodes"]
    i8 = res_tm.get("rot_engine") in (_capi.PG_ROT_I8SPLIT, _capi.PG_ROT_I8TC)
    fused = res_tm.get("rot_engine") == _capi.PG_ROT_I8TC
    n_planes = int(_capi.load().pg_rotation_planes())   # 7 unless the library was built with -DPG_SLICES=6
    rot_ops = (n_planes if i8 else 1) * 2.0 * n * n          # int8 (or fp64) multiply-add ops per SNP
    cmp_flops = 2.0 * n * 10 * (c0 + 2)                       # compression: n x kCq x (c0+2) FP64 FMAs per SNP
    stages = {
        "rotation": {"ms": rot_ms, "achieved": rot_ops * m / (rot_ms * 1e-3) / 1e12,
                     "peak": peaks["int8_gemm_tops"] if i8 else peaks["fp64_dmma_tflops"],
                     "kernel": (f"rotation U^T X, exact int8-split ({n_planes} base-256 digit planes): rotate_i8_tc2_kernel (hand-written TMA + "
                                "tcgen05 cta_group::2 kind::i8, recombination fused)") if (i8 and fused) else
                               (f"rotation U^T X, exact int8-split ({n_planes} base-256 digit planes): cuBLAS int8 GEMM (cutlass3x sm100 "
                                "tcgen05 2-SM kernel) + combine_i8_kernel") if i8 else
                               "rotation U^T X (cuBLAS DGEMM, FP64 tensor pipe)",
                     "peak_source": ("int8 tensor rate (cuBLAS int8 GEMM 16384x8192x8192) " if i8 else "FP64 DMMA rate ")
                                    + peaks["source_fp64"],
                     "algorithmic_ops_per_snp": rot_ops},
        "compress": {"ms": cmp_ms, "achieved": cmp_flops * m / (max(cmp_ms, 1e-9) * 1e-3) / 1e12,
                     "peak": peaks["fp64_dmma_tflops"], "kernel": "compress_dmma_kernel (FP64 tensor pipe, mma.sync m8n8k4)",
                     "peak_source": "FP64 DMMA rate " + peaks["source_fp64"], "algorithmic_ops_per_snp": cmp_flops},
    }
    dom = max(stages, key=lambda k: stages[k]["ms"])
    if solve_ms > stages[dom]["ms"]:
        # the optimiser on the compressed moments: FP64 FMA work = passes x nodes x (c0+2) x powers
        sol_flops = 2.0 * nodes * (c0 + 2) * (2 * ev2 + 3 * ev3)
        stages["solve"] = {"ms": solve_ms, "achieved": sol_flops * m / (solve_ms * 1e-3) / 1e12, "peak": peaks["fp64_fma_tflops"],
                           "kernel": "reml_solve_kernel", "peak_source": "FP64 FMA rate " + peaks["source_fp64"],
                           "algorithmic_ops_per_snp": sol_flops}
        dom = "solve"
    st = stages[dom]
    roofline = {"kernel": st["kernel"], "bound": "tensor" if dom != "solve" else "fp64", "achieved": st["achieved"],
                "peak": st["peak"], "unit": "TFLOP/s", "frac": st["achieved"] / st["peak"], "traffic": None,
                "peak_source": st["peak_source"], "algorithmic_ops_per_snp": st["algorithmic_ops_per_snp"],
                "ops": "int8 multiply-add ops counted 2 per MAC" if (dom == "rotation" and i8) else "fp64 flops",
                "share_of_step": st["ms"] / step_ms}
    # DRAM traffic of the dominant stage per launch (= per SNP block), from the committed ncu capture
    per_snp = None
    if dom == "rotation" and i8 and fused:
        per_snp = ncu_traffic("rotate_i8_tc2")
    elif dom == "rotation" and i8:
        a_, b_ = ncu_traffic("cutlass"), ncu_traffic("combine_i8")
        per_snp = (a_ + b_) if (a_ and b_) else None
    elif dom == "compress":
        per_snp = ncu_traffic("compress_dmma")
    elif dom == "solve":
        per_snp = ncu_traffic("reml_solve")
    roofline["traffic"] = per_snp * res_tm["block_snps"] if per_snp else None
    roofline["traffic_note"] = ("dram__bytes_read+write of the stage's kernels per SNP block (ncu --set full, n=10000, "
                                "profiles/ncu_r01_tc2_final_16384snps.json / ncu_r01_reml_final_8192snps.json); algorithmic HBM bytes "
                                "per SNP for the fused rotation: 10 KB int8 in + 80 KB fp64 out (+ the 70 MB digit planes "
                                "once per eigen-tile group); the cuBLAS form adds 2 x 280 KB of int32 partial products")
    roofline["per_kernel_ms_last_step"] = {"convert": conv_ms, "rotate": rot_ms, "compress": cmp_ms, "solve": solve_ms}
    roofline["rotate_tflops_fp64_equiv"] = 2.0 * n * n * m / (rot_ms * 1e-3) / 1e12
    roofline["compress_frac_of_dmma_peak"] = stages["compress"]["achieved"] / stages["compress"]["peak"]
    roofline["rotated_genotype_hbm_gbs"] = 2 * 8.0 * n * m / ((rot_ms + cmp_ms) * 1e-3) / 1e9  # written once, read once
    roofline["nodes_per_snp"] = nodes

    line = {
        "metric": METRIC, "value": value, "unit": "SNPs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": f"synthetic UKB-shape n={n} samples x {m} SNPs per GPU, c0={c0} covariates, "
                               f"{'grid-search' if grid else 'Brent+Newton'} lambda, int8 dosages (BASELINE.json configs[2])",
                   "n": n, "snps_per_gpu": m, "c0": c0, "grid": grid, "parallelism": f"snp-shard x{world}",
                   "l2": "inputs_larger_than_l2 (1 GB int8 genotypes + 80 KB/SNP rotated fp64 per step)",
                   "reml_engine": "compressed (eigenvalue-space moments)",
                   "rotation_engine": ("int8-split fused tcgen05" if fused else "int8-split cuBLAS") if i8 else "fp64"},
        "e2e": {"value": e2e_value, "unit": "SNPs/s", "h2d_bytes_per_step": int(n) * int(m) * world,
                "d2h_bytes_per_step": int(m) * (6 * 8 + 3 * 4) * world, "ms_per_step": e2e_ms,
                "device_ms_per_step": e2e_dev_ms / args.steps,
                "api": "pg_scan (C ABI, host buffers) as called by pygemma_b200.lmm.pygemma",
                "timing_last_call": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in e2e_tms[-1]["timing"].items()}},
        "gpu_launches": int(sum(t["convert_launches"] + t["reml_launches"] + t["rotate_launches"] for t in tms)),
        "clocks": clocks, "roofline": roofline,
        "setup": {"syevd_ms": eig_ms, "eigen_setup_s": setup_s, "design_tables_ms": design_ms},
        "evals_per_snp": {"two_power_passes": ev2, "three_power_passes": ev3}, "nan_rows": bad,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = run_cpu_reference(n, c0, grid, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": cb["snps_per_s"], "unit": "SNPs/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"], "scan_only_snps_per_s": cb["snps_per_s_scan_only"]}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "SNPs/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e}"}
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
    cb = run_cpu_reference(args.n, args.c0, bool(args.grid), steps=args.steps, warmup=args.warmup)
    line = {
        "metric": METRIC, "value": cb["snps_per_s"], "unit": "SNPs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (reference mix)", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"synthetic UKB-shape n={args.n} samples, c0={args.c0} covariates, "
                               f"{'grid-search' if args.grid else 'Brent+Newton'} lambda; bounded sample of "
                               f"{cb['m_sample']} SNPs per step (BASELINE.json configs[2])",
                   "n": args.n, "c0": args.c0, "grid": bool(args.grid)},
        "cpu_baseline": {"value": cb["snps_per_s"], "unit": "SNPs/s", "cores": cb["cores"], "kind": cb["kind"],
                         "sample": cb["sample"]},
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
    ap.add_argument("--snps", type=int, default=100000, help="SNPs per GPU per step")
    ap.add_argument("--n", type=int, default=N_SAMPLES)
    ap.add_argument("--c0", type=int, default=C0)
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
