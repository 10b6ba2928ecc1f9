"""Drop-in for the two imagereader.py helpers on the inference path (imagereader.py:34-50).
The LMDB training reader is out of scope."""
import numpy as np

from yolo3_b200 import post_engine


def zscore_normalize(image_data):
    """(x - mean) / std with the population std over the whole array; x - mean when std <= 1.0 (any shape).
    Runs on the GPU (y3_zscore: the statistics / normalisation kernels of the tiled front-end)."""
    return post_engine().zscore(np.asarray(image_data))


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


def imwrite(img, fp):
    """skimage.io.imsave replacement (imagereader.py:52-53)."""
    img = np.asarray(img)
    try:
        import cv2
        out = img[:, :, [2, 1, 0] + list(range(3, img.shape[2]))] if img.ndim == 3 and img.shape[2] >= 3 else img
        if cv2.imwrite(fp, out):
            return
    except ImportError:
        pass
    from PIL import Image
    Image.fromarray(img).save(fp)


def format_image(image_data):
    """HWC -> CHW (imagereader.py:56-59)."""
    return np.transpose(image_data, [2, 0, 1])
