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
        img = imagereader.imread(fp)
        height, width, channels = img.shape
        img = imagereader.zscore_normalize(img.astype(np.float32))
        print('  img.shape={}'.format(img.shape))
        batch = np.ascontiguousarray(img.transpose((2, 0, 1))[None], dtype=np.float32)
        boxes = np.array(yolo_model(batch, training=False))[0]
        boxes[:, 0] = np.clip(boxes[:, 0], 0, width)
        boxes[:, 1] = np.clip(boxes[:, 1], 0, height)
        boxes[:, 2] = np.clip(boxes[:, 2], 0, width)
        boxes[:, 3] = np.clip(boxes[:, 3], 0, height)
        boxes = bbox_utils.filter_small_boxes(boxes, min_box_size)
        kept, scores, labels = bbox_utils.per_class_nms(boxes[:, 0:4], boxes[:, 4:5], boxes[:, 5:])
        if kept is None:
            out = np.zeros((0, 5), np.int32)
        else:
            kept = kept.copy()
            kept[:, 2] -= kept[:, 0]
            kept[:, 3] -= kept[:, 1]
            out = np.concatenate((kept, labels.reshape(-1, 1)), axis=-1).astype(np.int32)
        print('Found: {} rois'.format(out.shape[0]))
        bbox_utils.write_boxes_from_xywhc(out, os.path.join(output_folder, file_name.replace(image_format, 'csv')))


if __name__ == "__main__":
    parser = argparse.ArgumentParser(prog='inference', description='Script to detect stars with the selected model')
    parser.add_argument('--saved-model-filepath', type=str, required=True, help='Filepath to the saved model to use')
    parser.add_argument('--output-folder', type=str, required=True)
    parser.add_argument('--image-folder', dest='image_folder', type=str, required=True,
                        help='filepath to the folder containing tif images to inference (Required)')
    parser.add_argument('--image-format', dest='image_format', type=str, default='tif',
                        help='format (extension) of the input images. E.g {tif, jpg, png)')
    parser.add_argument('--min-box-size', type=int, default=32, help='Smallest detection to consider. Default (32, 32).')
    args = parser.parse_args()
    print('Arguments:')
    for k, v in sorted(vars(args).items()):
        print('{} = {}'.format(k, v))
    os.environ["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
    os.environ["CUDA_VISIBLE_DEVICES"] = "0"       # the reference pins inference.py to GPU 0 (inference.py:131-133)
    inference(args.image_folder, args.image_format, args.saved_model_filepath, args.output_folder, args.min_box_size)
