"""yolo3_b200 - host side of the B200-native tiled YOLOv3 inference path (ctypes over libyolo3_b200.so)."""
from ._lib import LIB_PATH, Y3Error  # noqa: F401
from .engine import (DEFAULT_ANCHORS, Engine, gather_rows, infer_tiled_distributed, pinned_copy, pinned_empty, post_engine,  # noqa: F401
                     seam_candidates, shard_range, tile_count, tile_plan, batch_plan)
from . import weights  # noqa: F401
