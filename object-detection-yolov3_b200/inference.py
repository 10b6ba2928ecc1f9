"""Drop-in for the reference's inference.py (inference.py:24-135): same function, arguments and
command line; per-image forward -> decode -> small-box filter -> per-class NMS on the GPU.

Deliberate deviation: the reference's box clipping (inference.py:62-65) cannot run
as written (item assignment into an EagerTensor, batch axis indexed as the box axis) and it crashes
when NMS keeps nothing.  This implements the evident intent - clip x to [0, width] and y to
[0, height] before filtering - and writes a header-only CSV for an empty result.
"""
import argparse
import os

import numpy as np

import bbox_utils
import imagereader
import model


def inference(image_folder, image_format, saved_model_filepath, output_folder, min_box_size):
    os.makedirs(output_folder, exist_ok=True)
    image_format = image_format[1:] if image_format.startswith('.') else image_format
    files = [os.path.join(image_folder, fn) for fn in os.listdir(image_folder) if fn.endswith('.{}'.format(image_format))]
    yolo_model = model.load_saved_model(saved_model_filepath, max_batch=1)
    print('Starting inference of file list')
    for i, fp in enumerate(files):
        file_name = os.path.basename(fp)
        print('{}/{} : {}'.format(i, len(files), file_name))
        rois = _detect_one(yolo_model, imagereader.imread(fp), min_box_size)
        print('Found: {} rois'.format(rois.shape[0]))
        bbox_utils.write_boxes_from_xywhc(rois, os.path.join(output_folder, file_name.replace(image_format, 'csv')))


def _detect_one(yolo_model, img, min_box_size):
    """One image through the model: z-score, forward + decode, clip to the image, small-box filter, per-class NMS
    -> int32 [n, 5] rows x, y, w, h, class (inference.py:47-98)."""
    if img.ndim == 2:
        img = img[:, :, None]
    height, width = img.shape[0], img.shape[1]
    engine = yolo_model.engine_for((height, width)) if hasattr(yolo_model, "engine_for") else getattr(yolo_model, "engine", None)
    if engine is not None and engine.img_size is not None and tuple(engine.img_size) == tuple(img.shape):
        # the whole per-image body as ONE library call (y3_detect_image): z-score, forward, decode, clip, filter, NMS
        print('  img.shape={}'.format(img.shape))
        kept, _, labels = engine.detect_image(np.ascontiguousarray(img), min_box_size)
        if kept.shape[0] == 0:
            return np.zeros((0, 5), np.int32)
        xywh = kept.copy()
        xywh[:, 2] -= xywh[:, 0]
        xywh[:, 3] -= xywh[:, 1]
        return np.concatenate((xywh, labels.reshape(-1, 1)), axis=-1).astype(np.int32)
    # foreign model object: the reference's sequence of calls, each stage on the GPU
    norm = imagereader.zscore_normalize(img.astype(np.float32))
    print('  img.shape={}'.format(norm.shape))
    batch = np.ascontiguousarray(norm.transpose((2, 0, 1))[None], dtype=np.float32)
    dets = np.array(yolo_model(batch, training=False))[0]
    for col, hi in ((0, width), (1, height), (2, width), (3, height)):
        dets[:, col] = np.clip(dets[:, col], 0, hi)
    dets = bbox_utils.filter_small_boxes(dets, min_box_size)
    kept, _, labels = bbox_utils.per_class_nms(dets[:, 0:4], dets[:, 4:5], dets[:, 5:])
    if kept is None:
        return np.zeros((0, 5), np.int32)
    xywh = kept.copy()
    xywh[:, 2] -= xywh[:, 0]
    xywh[:, 3] -= xywh[:, 1]
    return np.concatenate((xywh, labels.reshape(-1, 1)), axis=-1).astype(np.int32)


def _parse_args(argv=None):
    """Same flags and defaults as the reference CLI (inference.py:103-114)."""
    ap = argparse.ArgumentParser(prog="inference", description="YOLOv3 detection over a folder of images (B200 path)")
    ap.add_argument("--saved-model-filepath", type=str, required=True, help="model directory (TF SavedModel or y3 side-car)")
    ap.add_argument("--output-folder", type=str, required=True, help="where the per-image CSV files go")
    ap.add_argument("--image-folder", dest="image_folder", type=str, required=True, help="folder with the input images")
    ap.add_argument("--image-format", dest="image_format", type=str, default="tif", help="image file extension, e.g. tif, jpg, png")
    ap.add_argument("--min-box-size", type=int, default=32, help="drop detections not larger than this in both directions")
    return ap.parse_args(argv)


if __name__ == "__main__":
    cli = _parse_args()
    print("Arguments:")
    for key in sorted(vars(cli)):
        print("{} = {}".format(key, getattr(cli, key)))
    # the reference pins inference.py to GPU 0 (inference.py:131-133)
    os.environ.update(CUDA_DEVICE_ORDER="PCI_BUS_ID", CUDA_VISIBLE_DEVICES="0")
    inference(cli.image_folder, cli.image_format, cli.saved_model_filepath, cli.output_folder, cli.min_box_size)
