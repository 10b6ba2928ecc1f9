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


@pytest.mark.parametrize("env", [{"Y3_NMS_GLOBAL_SORT": "1"}, {"Y3_NO_POST_FUSE": "1", "Y3_NO_POST_PDL": "1"}, {"Y3_SEAM_SERIAL": "1"},
                                 {"Y3_GRAPH_MAX_BATCH": "0"}],
                         ids=lambda e: "+".join(sorted(e)))
def test_post_processing_pipeline_variants(env):
    """The post-processing A/B switches: Y3_NMS_GLOBAL_SORT (round 1's global radix sort, host-synchronised: kept as the route
    for very large segments), Y3_NO_POST_FUSE / Y3_NO_POST_PDL (separate scan / bin / out-scan / emit kernels - the form large
    segment tables always use - without programmatic dependent launch), Y3_SEAM_SERIAL (cross-seam stage through the general
    pipeline), Y3_GRAPH_MAX_BATCH=0 (no CUDA-graph forward).  The NMS, detect and tiled suites must hold with each, bit for bit."""
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", "not k3_200k and not fp32_oracle_network and not 2048",
                        os.path.join(HERE, "test_gpu_nms.py"), os.path.join(HERE, "test_gpu_tiled_e2e.py"),
                        os.path.join(HERE, "test_gpu_net.py") + "::test_detect_pipeline_exact_on_own_boxes",
                        os.path.join(HERE, "test_gpu_net.py") + "::test_detect_image_is_the_reference_flow_in_one_call"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
