"""world_size-2 `gloo` test of the multi-GPU host logic (pygemma_b200/multi.py) on CPU: SNP sharding as the reference's
SampleIter does it (lmm/lmm.py:427-434), result gathering in input order, rank-0-only work.  The device collectives
(NCCL broadcast of U and d) need GPUs and are covered by the 2-GPU bench run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, tmpdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from pygemma_b200 import multi

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ctx = multi.context(device=0)
        assert (ctx.rank, ctx.world_size, ctx.backend) == (rank, world, "gloo")
        a, b = ctx.shard(m)
        per = multi.shard_len(m, world)
        assert per == (-(-m // world) if m < 256 else 512)  # small inputs: SampleIter's ceil(m / nproc); else 128-aligned
        assert (a, b) == (min(rank * per, m), min((rank + 1) * per, m))
        # every rank fabricates "its" rows of one global, deterministic result table
        rng = np.random.default_rng(123)
        full = {k: rng.standard_normal(m) for k in multi.RESULT_KEYS[:6]}
        full.update({k: rng.integers(0, 40, m).astype(np.int32) for k in multi.RESULT_KEYS[6:]})
        full["beta"][m // 2] = np.nan  # a failed row must survive the gather as NaN
        mine = {k: v[a:b].copy() for k, v in full.items()}
        out = multi.gather_results(ctx, mine, m)
        for k in multi.RESULT_KEYS:
            assert out[k].shape == (m,)
            assert np.array_equal(out[k], full[k], equal_nan=True), (rank, k)
        # the eigen broadcast is a device collective: under gloo it must refuse, not fall back to a CPU path
        class _H:
            n = 4
        with pytest.raises(RuntimeError, match="nccl"):
            multi.setup_eigen(ctx, _H(), None)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m", [11, 64, 1, 1000])
def test_world_size_2_gloo_shard_and_gather(m, tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, m, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
