"""Weight containers for the B200 engine: the Keras variable inventory of the reference network,
Keras-default random initialisation (what `model.YoloV3(...)` gives before training), and a
TensorFlow-free on-disk format for `--saved-model-filepath`.

The reference keeps weights in a TF SavedModel (train.py:221, inference.py:35).  TensorFlow is not
available here, so a model directory is either
  * the reference's own SavedModel: `saved_model.pb` + `variables/variables.{index,data-*}` - the variable bundle
    is parsed without TensorFlow by tf_bundle.py (anchors are graph constants that the bundle does not hold: they
    come from y3_config.json next to saved_model.pb, else from the constants in saved_model.pb, else from the
    Y3_ANCHORS environment variable - never from a silent default), or
  * the TensorFlow-free side-car format written by save_model_dir():
    <dir>/y3_config.json     {"img_size":[H,W,C], "number_classes":NC, "anchors":[[w,h],...]}
    <dir>/y3_weights.npz     one fp32 array per Keras variable name, Keras layouts
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


def load_model_dir(path, anchors=None):
    """-> (config dict, {name: fp32 array}).  cfg["img_size"] may hold None for H/W when the directory is a
    TF SavedModel whose signature does not pin them (the engine is then built for the size it is called with)."""
    cfg_p, w_p = os.path.join(path, CONFIG_FILE), os.path.join(path, WEIGHTS_FILE)
    cfg = None
    if os.path.exists(cfg_p):
        with open(cfg_p) as fh:
            cfg = json.load(fh)
    if cfg is not None and os.path.exists(w_p):
        with np.load(w_p) as z:
            weights = {k: z[k] for k in z.files}
        return cfg, weights
    prefix = os.path.join(path, "variables", "variables")
    if os.path.exists(prefix + ".index"):
        return _load_tf_saved_model(path, prefix, cfg, anchors)
    raise RuntimeError("%s holds neither %s + %s nor a TensorFlow SavedModel (variables/variables.index)"
                       % (path, CONFIG_FILE, WEIGHTS_FILE))


def _anchors_from_env():
    """Y3_ANCHORS="64,384;384,64" - for SavedModels whose graph constants cannot be read."""
    txt = os.environ.get("Y3_ANCHORS")
    if not txt:
        return None
    try:
        out = [tuple(float(v) for v in pair.split(",")) for pair in txt.replace(" ", "").split(";") if pair]
    except ValueError:
        out = None
    if not out or any(len(a) != 2 for a in out):
        raise RuntimeError("Y3_ANCHORS must look like '64,384;384,64', got %r" % txt)
    return out


def _load_tf_saved_model(path, prefix, cfg, anchors=None):
    """anchors: explicit [(w,h), ...] (wins over everything else).  Otherwise, in this order: y3_config.json next to
    saved_model.pb, the anchor constants of the graph in saved_model.pb (model.py:163, 432-436), the Y3_ANCHORS
    environment variable.  There is NO silent default: anchors are not variables, a wrong table gives wrong boxes
    (and a wrong class count - train.py:33 uses 2 anchors) without any error."""
    from . import tf_bundle
    found = tf_bundle.normalize_keras_names(tf_bundle.read_keras_variables(prefix))
    stem = found.get("conv2d/kernel")
    head = found.get("feature_map_1/kernel")
    if stem is None or head is None or stem.ndim != 4 or head.ndim != 4:
        raise RuntimeError("%s: the variable bundle does not hold the YOLOv3 layers (conv2d/kernel, feature_map_1/kernel)" % path)
    c_img, det_c = int(stem.shape[2]), int(head.shape[3])
    pb = os.path.join(path, "saved_model.pb")
    source = "argument"
    if anchors is None and cfg and "anchors" in cfg:
        anchors, source = [tuple(a) for a in cfg["anchors"]], CONFIG_FILE
    if anchors is None:
        anchors, source = tf_bundle.saved_model_anchors(pb), "saved_model.pb"
    if anchors is None:
        anchors, source = _anchors_from_env(), "Y3_ANCHORS"
    if anchors is None:
        raise RuntimeError("%s: the anchors of this model are unknown - they are graph constants, not variables, and could not "
                           "be read from saved_model.pb.  Write them into %s ({\"anchors\": [[w,h],...]}) next to "
                           "saved_model.pb or set Y3_ANCHORS='w,h;w,h' (the reference trainer uses 64,384;384,64 - "
                           "train.py:33; model.py:433 defaults to 32,32;128,128;256,256)" % (path, CONFIG_FILE))
    anchors = [(float(a), float(b)) for a, b in anchors]
    if det_c % len(anchors) or det_c // len(anchors) < 6:
        raise RuntimeError("%s: %d detection channels do not fit the %d anchors from %s"
                           % (path, det_c, len(anchors), source))
    nc = det_c // len(anchors) - 5
    if cfg and "number_classes" in cfg and int(cfg["number_classes"]) != nc:
        raise RuntimeError("%s: %s says %d classes but %d detection channels / %d anchors give %d"
                           % (path, CONFIG_FILE, int(cfg["number_classes"]), det_c, len(anchors), nc))
    weights = {}
    for name, shape in variable_inventory(c_img, nc, len(anchors)):
        if name not in found:
            raise RuntimeError("%s: variable %s is missing from the SavedModel bundle" % (path, name))
        if tuple(found[name].shape) != tuple(shape):
            raise RuntimeError("%s: variable %s has shape %s, expected %s" % (path, name, found[name].shape, shape))
        weights[name] = np.ascontiguousarray(found[name], np.float32)
    hw = [None, None]
    if cfg and "img_size" in cfg:
        hw = [int(cfg["img_size"][0]), int(cfg["img_size"][1])]
    else:
        sig = tf_bundle.signature_input_shape(pb)       # [-1, C, H, W]
        if sig and sig[2] > 0 and sig[3] > 0:
            hw = [int(sig[2]), int(sig[3])]
    return {"img_size": [hw[0], hw[1], c_img], "number_classes": nc, "anchors": [[a, b] for a, b in anchors],
            "anchors_source": source}, weights
