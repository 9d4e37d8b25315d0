"""The C-ABI library loads on a CPU-only box and exports exactly what include/pygemma_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pygemma_b200 import _capi, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _capi.load()


def _declared():
    txt = open(os.path.join(ROOT, "include", "pygemma_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pygemma_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == names


def test_abi_version(lib):
    txt = open(os.path.join(ROOT, "include", "pygemma_b200.h")).read()
    assert lib.pg_abi_version() == int(re.search(r"#define PG_ABI_VERSION (\d+)", txt).group(1))


def test_default_build_keeps_seven_digit_planes(lib):
    """The shipped library is the default build: 56-bit fixed-point U (the 6-plane variant is a build option)."""
    assert lib.pg_rotation_planes() == 7


def test_no_cpu_fallback(lib):
    """Without a GPU pg_create must fail loudly (PG_ERR_NO_DEVICE), never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.PgError) as ei:
        _capi.Handle(100, 2, 0)
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_shapes(lib):
    h = ctypes.c_void_p()
    assert lib.pg_create(1, 0, 0, ctypes.byref(h)) == -1
    assert lib.pg_create(100, 63, 0, ctypes.byref(h)) == -1
    assert b"c0" in lib.pg_last_error(None)


def test_timing_struct_matches_header():
    assert ctypes.sizeof(_capi.PgTiming) == 6 * 4 + 6 * 4 + 3 * 4


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pygemma_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pygemma_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU or oracle", ""), os.path.join(dirpath, f)


def test_block_plan_covers_every_snp_once(lib):
    """pg_scan's SNP blocks (pg_probe_block_plan): contiguous, complete, none above the block size; host-resident
    genotypes start with short blocks growing by at most x 1.6 (+ rounding to 256), on whole 256-SNP boundaries, so
    that upload of block b+1 hides under the compute of block b."""
    for m, blk in [(100000, 25088), (12544, 3584), (1_000_000, 5120), (50177, 25088), (50176, 25088), (2 * 1536 + 1025, 1536),
                   (7, 32), (1, 32), (100000, 32768), (3000, 1024), (125056, 5120)]:
        for host in (True, False):
            for fixed in (True, False):
                b = _capi.block_plan(m, blk, host, fixed)
                assert b[0] == 0 and b[-1] == m and all(x < y for x, y in zip(b, b[1:])), (m, blk, b)
                sizes = [y - x for x, y in zip(b, b[1:])]
                assert max(sizes) <= blk
                if not host or fixed or m < 2 * blk:
                    assert sizes[:-1] == [blk] * (len(sizes) - 1)       # plain blocks
                    continue
                if blk > 1536:
                    assert sizes[0] == 1536
                for s0, s1 in zip(sizes, sizes[1:-1]):
                    assert s1 <= max(blk, s0 * 8 // 5 + 256) and s1 >= min(s0, blk), (m, blk, sizes)
                assert all(x % 256 == 0 for x in b[:-1]) or blk % 256, (m, blk, b)
    assert _capi.block_plan(0, 1024) == []
    assert lib.pg_probe_block_plan(-1, 1024, 1, 0, None, 0) == -1


def test_rotation_launch_plan_of_the_l2_pinned_planes(lib):
    """pg_probe_rotation_launches (host arithmetic of rotate_i8_tc2.cuh: persist_group_tiles): the bench's 25 088-SNP block at
    n = 10 000 takes one launch per 16 eigen tiles (36 MB of planes in the 40 MB set-aside); short blocks -- the ramp of
    host-resident input, one rank's shard of a problem split over 8 GPUs -- and n = 50 000, where not even one wave-filling
    group of planes fits, keep the single launch; so does a device without a persisting set-aside."""
    import ctypes

    B200 = dict(setaside=82903040, window=134217728, sms=148)   # cudaDevAttrMaxPersistingL2CacheSize / ...WindowSize on the pool

    def plan(n, mb, **kw):
        d = dict(B200, **kw)
        g = ctypes.c_int32(-1)
        k = lib.pg_probe_rotation_launches(n, mb, d["setaside"], d["window"], d["sms"], ctypes.byref(g))
        return k, g.value

    if os.environ.get("PG_TC2_PERSIST") == "0":
        assert plan(10000, 25088) == (1, 0)
        return
    assert plan(10000, 25088) == (20, 16)          # 313 eigen tiles in groups of 16
    assert plan(10000, 100000)[1] == 16
    assert plan(10000, 12544) == (1, 0)            # 25 SNP tiles x 16 < 8 waves of 74 clusters
    assert plan(10000, 1536) == (1, 0)
    assert plan(50000, 5120) == (1, 0)             # 11.2 MB per eigen tile: 3 tiles per 40 MB, far below 8 waves
    assert plan(10000, 25088, setaside=0) == (1, 0)
    assert plan(10000, 25088, window=1 << 20) == (1, 0)
    k, g = plan(4096, 65536)                       # small n: more tiles fit, the group stays at 16
    assert g == 16 and k == (4096 // 32 + 15) // 16
    assert lib.pg_probe_rotation_launches(0, 10, 1, 1, 148, None) < 0
