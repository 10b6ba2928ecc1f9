"""Weight containers for the B200 engine: the Keras variable inventory of the reference network,
Keras-default random initialisation (what `model.YoloV3(...)` gives before training), and a
TensorFlow-free on-disk format for `--saved-model-filepath`.

The reference keeps weights in a TF SavedModel (train.py:221, inference.py:35).  TensorFlow is not
available here, so a model directory is:
    <dir>/y3_config.json     {"img_size":[H,W,C], "number_classes":NC, "anchors":[[w,h],...]}
    <dir>/y3_weights.npz     one fp32 array per Keras variable name, Keras layouts
Reading the SavedModel's own `variables/` tensor bundle is the next item of SURVEY.md section 8(f).
"""
import json
import math
import os

import numpy as np

CONFIG_FILE = "y3_config.json"
WEIGHTS_FILE = "y3_weights.npz"


def variable_inventory(c_img, number_classes, number_anchors):
    """[(variable_name, shape)] in the reference's creation order (model.py:383-421, 356-380):
    72 conv_layers (Conv2D + BatchNormalization), 3 detection Conv2D, 2 Conv2DTranspose."""
    inv = []
    n = {"conv": 0, "bn": 0, "convt": 0}

    def nm(kind, base):
        k = n[kind]
        n[kind] += 1
        return base if k == 0 else "%s_%d" % (base, k)

    def conv(cin, cout, k):
        c, b = nm("conv", "conv2d"), nm("bn", "batch_normalization")
        inv.append((c + "/kernel", (k, k, cin, cout)))
        inv.append((c + "/bias", (cout,)))
        for v in ("gamma", "beta", "moving_mean", "moving_variance"):
            inv.append((b + "/" + v, (cout,)))

    def block(c, reps):
        for _ in range(reps):
            conv(c, c // 2, 1)
            conv(c // 2, c, 3)

    def yolo(cin, f):
        for a, b, k in ((cin, f // 2, 1), (f // 2, f, 3), (f, f // 2, 1), (f // 2, f, 3), (f, f // 2, 1), (f // 2, f, 3)):
            conv(a, b, k)

    def det(cin, idx):
        inv.append(("feature_map_%d/kernel" % idx, (1, 1, cin, number_anchors * (5 + number_classes))))
        inv.append(("feature_map_%d/bias" % idx, (number_anchors * (5 + number_classes),)))

    def convt(c):
        t = nm("convt", "conv2d_transpose")
        inv.append((t + "/kernel", (2, 2, c, c)))
        inv.append((t + "/bias", (c,)))

    conv(c_img, 32, 3)
    conv(32, 64, 3)
    block(64, 1)
    conv(64, 128, 3)
    block(128, 2)
    conv(128, 256, 3)
    block(256, 8)
    conv(256, 512, 3)
    block(512, 8)
    conv(512, 1024, 3)
    block(1024, 4)
    yolo(1024, 1024)
    det(1024, 1)
    conv(512, 512, 1)
    convt(512)
    yolo(1024, 512)
    det(512, 2)
    conv(256, 256, 1)
    convt(256)
    yolo(512, 256)
    det(256, 3)
    return inv


def random_init(c_img, number_classes, number_anchors, seed=0, randomize_bn=False):
    """Keras defaults: glorot-uniform kernels, zero biases, BatchNorm gamma=1 beta=0 mean=0 var=1,
    Conv2DTranspose kernel = ones (model.py:103).  randomize_bn draws non-trivial BN statistics so
    that the fused (scale, shift) epilogue is exercised by synthetic runs."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in variable_inventory(c_img, number_classes, number_anchors):
        layer, var = name.split("/")
        if var == "kernel":
            if layer.startswith("conv2d_transpose"):
                w[name] = np.ones(shape, np.float32)
            else:
                k, _, cin, cout = shape
                lim = math.sqrt(6.0 / (k * k * cin + k * k * cout))
                w[name] = rng.uniform(-lim, lim, shape).astype(np.float32)
        elif var in ("bias", "beta", "moving_mean"):
            w[name] = np.zeros(shape, np.float32)
            if randomize_bn and not layer.startswith(("feature_map", "conv2d_transpose")):
                w[name] = (rng.standard_normal(shape) * (0.05 if var == "bias" else 0.1)).astype(np.float32)
        else:  # gamma, moving_variance
            w[name] = np.ones(shape, np.float32)
            if randomize_bn:
                w[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
    return w


def save_model_dir(path, weights, img_size, number_classes, anchors):
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, CONFIG_FILE), "w") as fh:
        json.dump({"img_size": [int(v) for v in img_size], "number_classes": int(number_classes),
                   "anchors": [[float(a), float(b)] for a, b in anchors]}, fh)
    np.savez(os.path.join(path, WEIGHTS_FILE), **{k: np.asarray(v, np.float32) for k, v in weights.items()})


def load_model_dir(path):
    """-> (config dict, {name: fp32 array}).  Raises with a precise message for a raw TF SavedModel."""
    cfg_p, w_p = os.path.join(path, CONFIG_FILE), os.path.join(path, WEIGHTS_FILE)
    if not (os.path.exists(cfg_p) and os.path.exists(w_p)):
        if os.path.exists(os.path.join(path, "saved_model.pb")):
            raise RuntimeError(
                "%s is a TensorFlow SavedModel without the %s / %s side-car.  Reading the TF variable bundle "
                "without TensorFlow is not implemented yet; export the Keras variables with "
                "yolo3_b200.weights.save_model_dir()." % (path, CONFIG_FILE, WEIGHTS_FILE))
        raise RuntimeError("%s does not contain %s and %s" % (path, CONFIG_FILE, WEIGHTS_FILE))
    with open(cfg_p) as fh:
        cfg = json.load(fh)
    with np.load(w_p) as z:
        weights = {k: z[k] for k in z.files}
    return cfg, weights
