"""Multi-GPU parity on hardware (needs >= 2 GPUs: `gpurun --gpus 2`): N-rank sharded output == 1-rank output, row for row."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_output_equals_single_gpu_on_every_rank():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the 1-rank code path is covered by test_sharded_entry_with_one_rank_equals_infer_tiled)")
    world = 2 if n < 4 else 4
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    res = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    print(res)
    assert res["all_ranks_equal_single"] and res["rows"] > 10 and res["rows_seam"] <= res["rows"]
