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
