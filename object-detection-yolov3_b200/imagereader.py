"""Drop-in for the two imagereader.py helpers on the inference path (imagereader.py:34-50).
The LMDB training reader is out of scope."""
import numpy as np

from yolo3_b200 import post_engine


def zscore_normalize(image_data):
    """(x - mean) / std with population std over the whole array; x - mean when std <= 1.0.
    Runs on the GPU as a single full-size tile (same kernel as the tiled front-end)."""
    a = np.asarray(image_data)
    if a.dtype not in (np.uint8, np.uint16, np.int32, np.float32):
        a = a.astype(np.float32)
    shape = a.shape
    hwc = a.reshape(shape[0], -1, 1) if a.ndim != 3 else a
    h, w, c = hwc.shape
    ph, pw = (-h) % 32, (-w) % 32
    if ph or pw:
        raise ValueError("zscore_normalize on the GPU needs H and W to be multiples of 32 (got %dx%d)" % (h, w))
    out = post_engine().tiles_normalized(np.ascontiguousarray(hwc), (h, w), 0)[0]     # [C,H,W]
    return np.ascontiguousarray(out.transpose(1, 2, 0)).reshape(shape)


def imread(fp):
    """skimage.io.imread replacement (skimage is not a dependency): HxW or HxWxC array."""
    try:
        import cv2
        img = cv2.imread(fp, cv2.IMREAD_UNCHANGED)
        if img is not None:
            if img.ndim == 3 and img.shape[2] >= 3:
                img = img[:, :, [2, 1, 0] + list(range(3, img.shape[2]))]
            return img
    except ImportError:
        pass
    from PIL import Image
    return np.asarray(Image.open(fp))
