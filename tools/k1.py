"""K1 (416x416x3, batch 1, NC=80) latency through y3_detect_image - measurement helper (see bench.py aux_k1)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
import bench  # noqa: E402

r = bench.aux_k1(0, None, {}, None, None)
print(json.dumps({k: r[k] for k in ("ms_per_image_wall_median", "ms_device_total", "ms_device_conv", "ms_device_decode_nms", "ms_device_zscore",
                                    "pipeline_bit_exact_on_own_boxes", "heads_rel_err_vs_fp32_oracle")}))
