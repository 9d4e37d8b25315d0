"""Multi-GPU plumbing for the scan: one process per GPU under torch.distributed.

The path shards by SNP exactly like the reference shards across processes
(SampleIter, reference lmm/lmm.py:413-436: `nproc` contiguous chunks of ceil(m/nproc) columns,
results concatenated in chunk order).  Collectives:
  * once per call: broadcast of U (n*n doubles) and d (n doubles) from the rank that ran the
    eigendecomposition (NCCL over NVLink when the backend is nccl);
  * at the end: all-gather of the per-SNP result rows (9 numbers per SNP).
There is no exchange inside the data path.  With no process group (or world_size 1) every function
here degenerates to the single-GPU behaviour.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

RESULT_KEYS = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald", "status", "n_eval2", "n_eval3"]
LRT_KEYS = ["lambda_ml", "loglik_ml", "D_lrt", "p_lrt"]   # present when the scan was asked for the likelihood-ratio outputs
_INT_KEYS = {"status", "n_eval2", "n_eval3"}


def _keys_of(res: dict):
    return RESULT_KEYS + (LRT_KEYS if "D_lrt" in res else [])


@dataclass
class Context:
    rank: int = 0
    world_size: int = 1
    device: int = 0
    backend: str = ""

    def shard(self, m: int):
        """Contiguous column range [a, b) of this rank: ceil(m / world) columns per rank (lmm/lmm.py:429-431)."""
        return shard_range(m, self.rank, self.world_size)


SHARD_ALIGN = 128  # columns: shard starts stay 128-byte aligned inside an int8 row, so device-resident shards of a
#                    sample-major matrix can be used in place as TMA / GEMM operands


def shard_len(m: int, world: int) -> int:
    """Columns per rank: ceil(m / world) as in the reference's SampleIter (lmm/lmm.py:429), rounded up to SHARD_ALIGN
    when every rank still gets work (results do not depend on where the chunk boundaries fall)."""
    per = -(-m // world) if world > 0 else m
    if world > 1 and per >= SHARD_ALIGN:
        al = -(-per // SHARD_ALIGN) * SHARD_ALIGN
        if al * (world - 1) < m:
            per = al
    return per


def shard_range(m: int, rank: int, world: int):
    per = shard_len(m, world)
    a = min(rank * per, m)
    b = min((rank + 1) * per, m)
    return a, b


def _dist():
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def context(device=None) -> Context:
    dist = _dist()
    if dist is None:
        return Context(0, 1, int(device) if device is not None else int(os.environ.get("PYGEMMA_B200_DEVICE", 0)), "")
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    return Context(rank, world, int(device), dist.get_backend())


def _coll_device(ctx: Context):
    import torch

    return torch.device(f"cuda:{ctx.device}") if ctx.backend == "nccl" else torch.device("cpu")


def check_backend(ctx: Context) -> None:
    """The eigen broadcast is a device collective: refuse any other backend BEFORE handles or K copies are made."""
    if ctx.world_size > 1 and ctx.backend != "nccl":
        raise RuntimeError("multi-GPU eigen broadcast needs the nccl backend (device buffers), "
                           f"the process group uses {ctx.backend!r}")


def agree_on_n(ctx: Context, n):
    """Rank 0 knows n (it holds K); the others may have been given K=None."""
    if ctx.world_size == 1:
        return n
    import torch
    import torch.distributed as dist

    t = torch.tensor([int(n) if ctx.rank == 0 else 0], dtype=torch.int64, device=_coll_device(ctx))
    dist.broadcast(t, src=0)
    return int(t.item())


# seconds spent in the collectives of the most recent call (bench.py reports them)
last_collective_s = {"broadcast": 0.0, "gather": 0.0}


def setup_eigen(ctx: Context, handle, K) -> float:
    """Eigendecomposition on rank 0, broadcast of U and d to the other ranks' handles.  Returns syevd ms.
    K is only read on rank 0."""
    if ctx.world_size == 1:
        _, ms = handle.set_kinship(K)
        return ms
    import time

    import torch
    import torch.distributed as dist

    n = handle.n
    check_backend(ctx)
    dev = _coll_device(ctx)
    U_t = torch.empty(n * n, dtype=torch.float64, device=dev)
    d_t = torch.empty(n, dtype=torch.float64, device=dev)
    ms = 0.0
    if ctx.rank == 0:
        _, ms = handle.set_kinship(K)
        handle.get_eigen_device(U_t.data_ptr(), d_t.data_ptr())
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    dist.broadcast(U_t, src=0)
    dist.broadcast(d_t, src=0)
    torch.cuda.synchronize(dev)
    last_collective_s["broadcast"] = time.perf_counter() - t0
    if ctx.rank != 0:
        handle.set_eigen_device(U_t.data_ptr(), False, d_t.data_ptr())
    del U_t, d_t
    return ms


def setup_eigen_from(ctx: Context, handle, U_t, d_t) -> None:
    """An eigen-system computed elsewhere (rank 0's device tensors: U column-major as n*n doubles, d) broadcast to
    every rank's handle -- the multi-GPU form of pg_set_eigen_device.  Other ranks pass their own empty tensors."""
    if ctx.world_size > 1:
        import time

        import torch
        import torch.distributed as dist

        check_backend(ctx)
        dev = _coll_device(ctx)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        dist.broadcast(U_t, src=0)
        dist.broadcast(d_t, src=0)
        torch.cuda.synchronize(dev)
        last_collective_s["broadcast"] = time.perf_counter() - t0
    handle.set_eigen_device(U_t.data_ptr(), False, d_t.data_ptr())


def gather_device(ctx: Context, out_dev, st_dev):
    """Device-resident form of gather_results: out_dev (6, per) float64 and st_dev (3, per) int32 of every rank are
    all-gathered into (world * 6, per) / (world * 3, per) tensors on the current stream (rank blocks stacked)."""
    import torch
    import torch.distributed as dist

    allo = torch.empty((ctx.world_size * out_dev.shape[0], out_dev.shape[1]), dtype=out_dev.dtype, device=out_dev.device)
    alls = torch.empty((ctx.world_size * st_dev.shape[0], st_dev.shape[1]), dtype=st_dev.dtype, device=st_dev.device)
    dist.all_gather_into_tensor(allo, out_dev)
    dist.all_gather_into_tensor(alls, st_dev)
    return allo, alls


def pack_results(res: dict, per: int, keys=None) -> np.ndarray:
    """(len(keys), per) float64 block of one rank's rows, NaN / 0 padded to the common shard length."""
    keys = keys or RESULT_KEYS
    k = res["beta"].shape[0]
    blk = np.full((len(keys), per), np.nan)
    for i, key in enumerate(keys):
        blk[i, :k] = res[key] if key in res else 0
    return blk


def unpack_results(blocks, m: int, world: int, keys=None) -> dict:
    """Concatenate rank blocks in rank order (= input column order) and trim the padding."""
    keys = keys or RESULT_KEYS
    out = {key: np.empty(m, dtype=np.int32 if key in _INT_KEYS else np.float64) for key in keys}
    for r, blk in enumerate(blocks):
        a, b = shard_range(m, r, world)
        for i, key in enumerate(keys):
            vals = blk[i, : b - a]
            out[key][a:b] = np.nan_to_num(vals).astype(np.int32) if key in _INT_KEYS else vals
    return out


_gather_cache: dict = {}


def _gather_buffers(ctx: Context, rows: int, per: int):
    """Pinned host staging + device buffers of one gather shape, kept across calls (a pageable round trip of the
    9 x m table costs several ms per call; the scan of an 8-GPU shard takes about as long)."""
    import torch

    key = (ctx.device, ctx.world_size, rows, per, ctx.backend)
    buf = _gather_cache.get(key)
    if buf is None:
        _gather_cache.clear()
        dev = _coll_device(ctx)
        pin = ctx.backend == "nccl"
        buf = {"mine_h": torch.empty((rows, per), dtype=torch.float64, pin_memory=pin),
               "all_h": torch.empty((ctx.world_size * rows, per), dtype=torch.float64, pin_memory=pin),
               "mine_d": torch.empty((rows, per), dtype=torch.float64, device=dev),
               "all_d": torch.empty((ctx.world_size * rows, per), dtype=torch.float64, device=dev)}
        _gather_cache[key] = buf
    return buf


def gather_results(ctx: Context, res: dict, m: int) -> dict:
    """All ranks end up with all m rows, in input order."""
    if ctx.world_size == 1:
        return res
    import time

    import torch
    import torch.distributed as dist

    per = shard_len(m, ctx.world_size)
    t0 = time.perf_counter()
    keys = _keys_of(res)
    buf = _gather_buffers(ctx, len(keys), per)
    mine = buf["mine_h"].numpy()
    k = res["beta"].shape[0]
    mine[:, k:] = np.nan
    for i, key in enumerate(keys):
        mine[i, :k] = res[key] if key in res else 0
    if ctx.backend == "nccl":
        buf["mine_d"].copy_(buf["mine_h"], non_blocking=True)
        dist.all_gather_into_tensor(buf["all_d"], buf["mine_d"])  # rank blocks stacked along dim 0
        buf["all_h"].copy_(buf["all_d"], non_blocking=True)
        torch.cuda.current_stream(buf["all_d"].device).synchronize()
    else:
        dist.all_gather_into_tensor(buf["all_h"], buf["mine_h"])
    blocks = buf["all_h"].numpy().reshape(ctx.world_size, len(keys), per)
    out = unpack_results(blocks, m, ctx.world_size, keys)
    last_collective_s["gather"] = time.perf_counter() - t0
    return out
