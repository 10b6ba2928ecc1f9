"""CPU-only: the C-ABI library loads and exports every symbol include/yolo3_b200.h declares, and
fails loudly (no fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="yolo3_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(y3_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from yolo3_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes out of sync with the header"
    assert lib.y3_abi_version() == _lib.ABI_VERSION == 3


def test_probe_library_is_separate():
    """the hardware probes of tests/probe_*.py are not in the product library"""
    from yolo3_b200 import _lib
    lib, probe = _lib.load(), _lib.load_probe()
    syms = declared_symbols("yolo3_b200_probe.h")
    assert sorted(_lib.PROBE_PROTOTYPES) == syms and len(syms) == 2
    for s in syms:
        assert hasattr(probe, s) and not hasattr(lib, s)


def test_tile_plan_host_logic(golden):
    from yolo3_b200 import tile_plan
    g = golden("tiling.npz")
    for tag, (h, w, tile, edge) in dict(u16=(700, 900, (512, 512), 96), u8rgb=(520, 1100, (256, 320), 64),
                                        small=(300, 280, (512, 512), 96), wide=(400, 1500, (512, 512), 96)).items():
        xs, ys = tile_plan(h, w, tile, edge)
        assert xs.tolist() == g[tag + "_xs"].tolist() and ys.tolist() == g[tag + "_ys"].tolist()
    from yolo3_b200 import tile_count
    assert tile_count(20000, 20000, (512, 512), 96) == 3969 and tile_count(20000, 20000, (512, 512), 64) == 2809


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from yolo3_b200 import Engine, Y3Error
    with pytest.raises(Y3Error) as e:
        Engine()
    assert e.value.code == -6 and "no CPU fallback" in str(e.value)


def test_batch_plan_host_logic():
    """y3_batch_plan: how the tiled entry points batch their tiles (pure host logic)"""
    from yolo3_b200 import batch_plan
    # the bench workload on 1 / 4 / 8 GPUs, image resident on the device and in host memory
    assert batch_plan(2809, 256, False) == [256] * 4 + [255] * 7
    assert batch_plan(352, 256, False) == [118, 117, 117]
    assert batch_plan(352, 256, True) == [48, 152, 152]
    assert batch_plan(703, 256, True) == [48, 219, 218, 218]
    for count in (0, 1, 2, 31, 47, 48, 49, 95, 96, 97, 200, 352, 703, 2809, 3969):
        for max_batch in (1, 4, 32, 118, 256):
            for host in (False, True):
                plan = batch_plan(count, max_batch, host)
                assert sum(plan) == count and all(0 < b <= max_batch for b in plan), (count, max_batch, host, plan)
                if count == 0:
                    assert plan == []
                    continue
                rest = plan[1:] if host else plan
                if host:
                    assert plan[0] == min(48, max_batch, count)
                if rest:
                    assert max(rest) - min(rest) <= 1                                 # even split, no small tail batch
                    assert len(rest) == max(-(-sum(rest) // max_batch), min(3 - (1 if host else 0), sum(rest) // 32), 1)
    import pytest
    with pytest.raises(ValueError):
        batch_plan(10, 0, False)
