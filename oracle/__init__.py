"""CPU oracle for the tiled YOLOv3 inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`object-detection-yolov3_b200/`) may import, call, link or execute anything
in this directory.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and there only
as the checker or the timed CPU baseline - never as the thing shipped.

Parity status (also stated in DESIGN.md):
  * a11-a17 (normalise, tiling, small-box filter, IoU, NMS, stitching):
    PINNED.  `oracle/postproc_np.py`, `oracle/tiling_np.py` and
    `oracle/nms_oracle.c` are checked against the reference's own NumPy code
    executed verbatim in the build container (`oracle/ref_loader.py`), and the
    resulting vectors are committed under `tests/golden/`.
  * a1-a10 (network forward, decode): PARITY UNPINNED.  The arithmetic lives in
    TensorFlow 2.x / Keras (unpinned in the reference, absent from
    /root/reference and not installable offline).  `oracle/model_torch.py`
    restates model.py:29-59, 94-212, 356-464 in torch fp32/fp64 on the CPU.
"""
