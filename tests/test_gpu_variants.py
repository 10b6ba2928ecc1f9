"""The kernel variants that are kept behind A/B switches must stay correct: each switch is
exercised in a subprocess and held to the same head tolerance as the default path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
HEAD_TOL = 2e-2

VARIANTS = [{}, {"Y3_NO_HALO": "1"}, {"Y3_NO_WS2": "1"}, {"Y3_NO_STEM_FUSE": "1"}, {"Y3_NO_STEM_FUSE": "1", "Y3_STEM_FP32": "1"},
            {"Y3_DISABLE_2CTA": "1"}, {"Y3_CONV2_OLD": "1"}, {"Y3_NO_IM2COL": "1", "Y3_NO_HALO": "1"}, {"Y3_NO_UPFUSE": "1"},
            {"Y3_PDL": "1"}, {"Y3_TAIL_BF16": "1"}]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: "+".join(sorted(e)) or "default")
def test_variant_heads(env):
    p = subprocess.run([sys.executable, os.path.join(HERE, "variant_heads.py")], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    errs = json.loads(line[len("RESULT "):])
    print(env, errs)
    assert all(max(v) <= HEAD_TOL for v in errs.values()), errs


def test_global_sort_pipeline_variant():
    """Y3_NMS_GLOBAL_SORT=1: the round-1 post-processing pipeline (global radix sort, host-synchronised) is kept as the route
    for very large segments and as an A/B switch - the NMS, detect and tiled suites must hold with it, bit for bit."""
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", "not k3_200k",
                        os.path.join(HERE, "test_gpu_nms.py"), os.path.join(HERE, "test_gpu_tiled_e2e.py"),
                        os.path.join(HERE, "test_gpu_net.py") + "::test_detect_pipeline_exact_on_own_boxes"],
                       env=dict(os.environ, Y3_NMS_GLOBAL_SORT="1"), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
