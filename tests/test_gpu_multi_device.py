"""Several GPUs driven by one process (pg_multi_*, lmm.pygemma(..., gpus=N)): needs >= 2 devices, skipped otherwise.
The result must be the single-GPU result bit for bit (SNP shards are independent units; every row is written by the device
that owns its shard, in input order)."""
import numpy as np
import pytest

from conftest import COLS

pytestmark = pytest.mark.gpu


def _ndev():
    import torch

    return torch.cuda.device_count()


def test_one_process_two_gpus_equal_one_gpu_bit_for_bit():
    if _ndev() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    from oracle import oracle
    from pygemma_b200 import _capi, lmm
    from pygemma_b200.synth import make_problem

    n, m, c0 = 1024, 5000 + 37, 4
    p = make_problem(n, 64, c0, seed=3, m_k=2 * n)
    X = np.random.default_rng(2).integers(0, 3, size=(n, m), dtype=np.int8)
    one = lmm.pygemma(p["Y"], X, p["W"], p["K"])
    two = lmm.pygemma(p["Y"], X, p["W"], p["K"], gpus=2, snps=np.arange(m))
    assert list(two.columns) == COLS + ["SNPs"] and len(two) == m
    for c in COLS:
        assert np.array_equal(one[c].to_numpy(), two[c].to_numpy(), equal_nan=True), c
    assert lmm.last_timing["devices"] == [0, 1] and lmm.last_timing["bcast_ms"] > 0
    ref = oracle.pygemma(p["Y"], X[:, ::97], p["W"], p["K"])
    for c in COLS:
        e = np.abs(two[c].to_numpy()[::97] - ref[c]) / np.maximum(np.abs(ref[c]), 1e-300)
        assert e.max() < 1e-6, (c, float(e.max()))
    # SNP-major input, grid mode, fewer columns than two aligned shards, a strided view, rotated inputs (eigen=False)
    with _capi.MultiHandle(n, c0, [1, 0]) as mh, _capi.Handle(n, c0) as h:
        mh.set_kinship(p["K"])
        h.set_kinship(p["K"])
        mh.set_design(p["W"], p["Y"])
        h.set_design(p["W"], p["Y"])
        for Xi, layout, grid in ((np.ascontiguousarray(X.T), _capi.PG_X_SNP_MAJOR, False), (X[:, :150], _capi.PG_X_SAMPLE_MAJOR, True),
                                 (X[:, 3::2], _capi.PG_X_SAMPLE_MAJOR, False)):
            a, b = mh.scan(Xi, grid=grid, layout=layout), h.scan(Xi, grid=grid, layout=layout)
            for c in COLS + ["status", "n_eval2", "n_eval3"]:
                assert np.array_equal(a[c], b[c], equal_nan=True), c
            assert len(a["timing_per_device"]) == 2
        with pytest.raises(_capi.PgError):   # errors of a device surface with that device's message
            mh._ck(mh.L.pg_multi_set_design(mh.h, None, None, 0, None))
