"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py launches it): the tile-sharded path with the library's own
NCCL all-gather must return, on EVERY rank, exactly the rows one GPU returns for the whole image - with and without the
cross-seam stage.  Prints RESULT {...} on rank 0."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "object-detection-yolov3_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import cases  # noqa: E402
from yolo3_b200 import infer_tiled_distributed  # noqa: E402
import test_gpu_tiled_e2e as t  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
img = cases.synthetic_image(1500, 1900, 1, np.uint16, seed=15, blobs=40)
import yolo3_b200.engine as E  # noqa: E402
_orig = E.Engine.__init__


def _on_my_gpu(self, *a, **k):
    k["device"] = local
    _orig(self, *a, **k)


E.Engine.__init__ = _on_my_gpu
eng = t.standardised_engine(img, 64, 8)
single = eng.infer_tiled(img, t.TILE, 24, edge_range=64)
sharded = infer_tiled_distributed(eng, img, t.TILE, 24, 64, out_device=None)
seam = infer_tiled_distributed(eng, img, t.TILE, 24, 64, cross_seam=True, out_device=None)
seam_single = eng.cross_seam_nms(single, img.shape[:2], t.TILE, 64, 0.3)
small_cap = eng.infer_tiled_sharded(img, t.TILE, 24, edge_range=64, cap=5)            # collective overflow + retry
ok = [bool(np.array_equal(single, sharded)), bool(np.array_equal(seam, seam_single)), bool(np.array_equal(small_cap, single))]
flags = torch.tensor([int(all(ok))], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RESULT " + json.dumps({"world": world, "rows": int(single.shape[0]), "rows_seam": int(seam.shape[0]), "all_ranks_equal_single": bool(flags.item()),
                                  "rank0": ok, "sha": sha(sharded), "ms_comm": eng.timings()["ms_comm"]}))
dist.destroy_process_group()
sys.exit(0 if flags.item() else 1)
