"""Drop-in for the reference's inference_tiled.py: same functions, arguments, module constants and
command line, with the work done by libyolo3_b200.so on the GPU.

  convert_image_to_tiles(img, tile_size)                  inference_tiled.py:29-100   -> y3_tiles_raw
  inference_image_tiled(yolo_model, img, tile_size, min)  inference_tiled.py:185-310  -> y3_infer_tiled
                                                          (or y3_tiles_normalized + model + y3_stitch_tiles
                                                          when yolo_model is a foreign callable)
  inference_image_folder(...) + CLI                       inference_tiled.py:313-382
The private NMS copies of the reference (inference_tiled.py:103-182) are the same functions as
bbox_utils' and are re-exported from there.  Launched under torchrun, the tile grid is sharded
across the ranks (one GPU each) and the boxes are all-gathered with NCCL.
"""
import argparse
import os

import numpy as np

import bbox_utils
import imagereader
import model
from bbox_utils import compute_iou, filter_small_boxes, per_class_nms, single_class_nms  # noqa: F401
from yolo3_b200 import infer_tiled_distributed, post_engine, tile_plan

BATCH_SIZE = 256         # tiles per forward batch at 512x512 (scaled down for larger tiles; declared = 1, unused, in the reference)
EDGE_EFFECT_RANGE = 96
# Optional extra stage, off by default because the reference has none (seams are resolved by centre ownership
# only): greedy per-class NMS among the final boxes that straddle a tile-zone boundary.  Set the constant or
# Y3_CROSS_SEAM_NMS=1 to enable.
CROSS_SEAM_NMS = os.environ.get("Y3_CROSS_SEAM_NMS", "0") not in ("", "0")
CROSS_SEAM_IOU_THRESHOLD = 0.3


def convert_image_to_tiles(img, tile_size):
    img = np.ascontiguousarray(img)
    assert tile_size[0] % model.YoloV3.NETWORK_DOWNSAMPLE_FACTOR == 0
    assert tile_size[1] % model.YoloV3.NETWORK_DOWNSAMPLE_FACTOR == 0
    xs, ys = tile_plan(img.shape[0], img.shape[1], tile_size, EDGE_EFFECT_RANGE)
    tiles = post_engine().tiles_raw(img, tile_size, EDGE_EFFECT_RANGE)
    return [t for t in tiles], [int(v) for v in xs], [int(v) for v in ys]


def _distributed():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except ImportError:
        return False


def inference_image_tiled(yolo_model, img, tile_size, min_roi_size):
    img = np.ascontiguousarray(img)
    if hasattr(yolo_model, "engine_for"):
        engine = yolo_model.engine_for(tile_size)
    else:
        engine = getattr(yolo_model, "engine", None)
    if engine is not None and tuple(tile_size) == tuple(engine.img_size[:2]):
        if _distributed():
            # tile grid sharded over the ranks; rows all-gathered (and, if enabled, the cross-seam stage run) by the library
            pred = infer_tiled_distributed(engine, img, tile_size, min_roi_size, EDGE_EFFECT_RANGE, cross_seam=CROSS_SEAM_NMS,
                                           out_device=None)
        else:
            pred = engine.infer_tiled(img, tile_size, min_roi_size, EDGE_EFFECT_RANGE)
            if CROSS_SEAM_NMS:
                pred = engine.cross_seam_nms(pred, img.shape[:2], tile_size, EDGE_EFFECT_RANGE, CROSS_SEAM_IOU_THRESHOLD)
    else:
        # foreign model object: slice + normalise on the GPU, call the model per tile as the
        # reference does, stitch on the GPU
        post = post_engine()
        tiles = post.tiles_normalized(img, tile_size, EDGE_EFFECT_RANGE)
        dets = np.stack([np.asarray(yolo_model(t[None], training=False))[0] for t in tiles])
        pred = post.stitch_tiles(dets, img.shape[:2], tile_size, min_roi_size, EDGE_EFFECT_RANGE)
        if CROSS_SEAM_NMS:
            pred = post.cross_seam_nms(pred, img.shape[:2], tile_size, EDGE_EFFECT_RANGE, CROSS_SEAM_IOU_THRESHOLD,
                                       number_classes=dets.shape[2] - 5)
    print('Found: {} rois'.format(pred.shape[0]))
    return pred


def inference_image_folder(image_folder, image_format, saved_model_filepath, output_folder, tile_size, min_roi_size):
    if not os.path.exists(saved_model_filepath):
        raise RuntimeError('Missing saved_model_filepath File')
    image_format = image_format[1:] if image_format.startswith('.') else image_format
    files = [os.path.join(image_folder, fn) for fn in os.listdir(image_folder) if fn.endswith('.{}'.format(image_format))]
    device = int(os.environ.get("LOCAL_RANK", "0")) if _distributed() else 0
    max_batch = max(1, min(BATCH_SIZE, BATCH_SIZE * 512 * 512 // max(1, int(tile_size[0]) * int(tile_size[1]))))
    yolo_model = model.load_saved_model(saved_model_filepath, max_batch=max_batch, device=device)
    os.makedirs(output_folder, exist_ok=True)
    print('Starting inference of file list')
    for i, fp in enumerate(files):
        file_name = os.path.basename(fp)
        print('{}/{} : {}'.format(i, len(files), file_name))
        img = imagereader.imread(fp)
        if img.ndim == 2:
            img = img[:, :, None]
        print('  img.shape={}'.format(img.shape))
        predictions = inference_image_tiled(yolo_model, img, tile_size, min_roi_size)
        bbox_utils.write_boxes_from_ltrbpc(predictions, os.path.join(output_folder, file_name.replace(image_format, 'csv')))


if __name__ == "__main__":
    parser = argparse.ArgumentParser(prog='inference', description='Script to detect stars with the selected model')
    parser.add_argument('--saved-model-filepath', type=str, required=True, help='Filepath to the saved model to use')
    parser.add_argument('--image-folder', type=str, required=True, help='Filepath to the folder of images to inference')
    parser.add_argument('--output-folder', type=str, required=True)
    parser.add_argument('--tile-height', type=int, default=512)
    parser.add_argument('--tile-width', type=int, default=512)
    parser.add_argument('--min-box-size', type=int, default=32)
    parser.add_argument('--image-format', dest='image_format', type=str, default='tif',
                        help='format (extension) of the input images. E.g {tif, jpg, png)')
    args = parser.parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    print('Arguments:')
    for k, v in sorted(vars(args).items()):
        print('{} = {}'.format(k, v))
    inference_image_folder(args.image_folder, args.image_format, args.saved_model_filepath, args.output_folder,
                           [args.tile_height, args.tile_width], args.min_box_size)
