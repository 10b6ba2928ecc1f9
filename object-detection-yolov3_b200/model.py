"""Drop-in for the inference surface of the reference's model.py (model.py:423-470): the YoloV3
class with the same constructor and the two `get_keras_*_model()` accessors, whose return values
are callables `m(batch_nchw_f32, training=False)` backed by the B200 engine.

Training (compute_loss / train_step / optimizer, model.py:214-354, 481-540) is out of scope and
raises NotImplementedError.
"""
import numpy as np

from yolo3_b200 import DEFAULT_ANCHORS, Engine, weights as _weights


class _Callable:
    def __init__(self, fn):
        self._fn = fn

    def __call__(self, batch, training=False):
        if training:
            raise NotImplementedError("the B200 path is inference-only")
        return self._fn(np.ascontiguousarray(np.asarray(batch), dtype=np.float32))


class YoloV3:
    BLOCK_COUNT = 8
    FILTER_COUNT = 1024
    KERNEL_SIZE = 3
    NETWORK_DOWNSAMPLE_FACTOR = 32
    WEIGHT_DECAY = 5e-4

    def __init__(self, global_batch_size, img_size, number_classes, anchors=None, learning_rate=1e-4, weights=None,
                 device=0, seed=0):
        self.number_classes = number_classes
        self.learning_rate = learning_rate
        self.global_batch_size = global_batch_size
        self.img_size = img_size
        self.score_threshold = 0.1
        self.iou_threshold = 0.5
        self.anchors = list(anchors) if anchors is not None else list(DEFAULT_ANCHORS)
        self.number_anchors = len(self.anchors)
        f = YoloV3.NETWORK_DOWNSAMPLE_FACTOR
        cells = sum((img_size[0] // (f >> s)) * (img_size[1] // (f >> s)) for s in range(3))
        self.number_output_boxes = self.number_anchors * cells
        self.output_shape = [self.number_output_boxes, 5 + self.number_classes]
        self.engine = Engine(img_size, number_classes, self.anchors, max_batch=max(1, int(global_batch_size)), device=device)
        # like Keras, a freshly constructed model is randomly initialised
        self.weights = weights if weights is not None else _weights.random_init(img_size[2], number_classes,
                                                                                self.number_anchors, seed=seed)
        self.engine.load_weights(self.weights)
        self.model = _Callable(self.engine.forward_boxes)
        self.model_feature_maps = _Callable(self.engine.forward_heads)

    def get_keras_model(self):
        return self.model

    def get_keras_feature_map_model(self):
        return self.model_feature_maps

    def save(self, path, fmt="y3"):
        """fmt="y3": y3_config.json + y3_weights.npz; fmt="tf": the reference's SavedModel layout (variables/ bundle +
        a saved_model.pb that carries the input signature) plus y3_config.json for the anchors."""
        if fmt == "tf":
            import json
            import os
            from yolo3_b200 import tf_bundle
            tf_bundle.write_saved_model_variables(path, self.weights, input_shape=[-1, self.img_size[2], self.img_size[0], self.img_size[1]],
                                                  anchors=self.anchors)
            with open(os.path.join(path, _weights.CONFIG_FILE), "w") as fh:
                json.dump({"anchors": [[float(a), float(b)] for a, b in self.anchors]}, fh)
            return
        _weights.save_model_dir(path, self.weights, self.img_size, self.number_classes, self.anchors)

    def get_optimizer(self):
        raise NotImplementedError("training is outside the B200 inference path")

    set_learning_rate = get_learning_rate = train_step = test_step = get_optimizer


class LoadedModel:
    """What `tf.saved_model.load(path)` is to the reference scripts: callable as
    yolo_model(batch, training=False) -> [B, N, 5+NC]; also carries the engine for the fused paths.
    Reads the reference's TF SavedModel directory (variable bundle parsed without TensorFlow) or the
    y3_config.json + y3_weights.npz side-car format.  The network is fully convolutional: when the model
    directory does not pin the input size the engine is built for the first size it is asked for."""

    def __init__(self, saved_model_filepath, max_batch=1, device=0, anchors=None):
        cfg, w = _weights.load_model_dir(saved_model_filepath, anchors=anchors)
        self._weights, self._max_batch, self._device = w, max_batch, device
        self._c_img = int(cfg["img_size"][2])
        self.number_classes = int(cfg["number_classes"])
        self.anchors = [tuple(a) for a in cfg["anchors"]]
        self._engines = {}
        self.engine = None
        self.img_size = None
        if cfg["img_size"][0] and cfg["img_size"][1]:
            self.engine_for(cfg["img_size"][:2])

    def engine_for(self, hw):
        """the engine for H x W inputs (built on first use, weights uploaded once per size)"""
        key = (int(hw[0]), int(hw[1]))
        if key not in self._engines:
            eng = Engine(key + (self._c_img,), self.number_classes, self.anchors, max_batch=self._max_batch, device=self._device)
            eng.load_weights(self._weights)
            self._engines[key] = eng
        self.engine = self._engines[key]
        self.img_size = key + (self._c_img,)
        return self.engine

    def __call__(self, batch, training=False):
        if training:
            raise NotImplementedError("the B200 path is inference-only")
        batch = np.ascontiguousarray(np.asarray(batch), dtype=np.float32)
        return self.engine_for(batch.shape[2:4]).forward_boxes(batch)


def load_saved_model(saved_model_filepath, max_batch=1, device=0, anchors=None):
    """anchors: only needed for a TF SavedModel whose anchor constants cannot be read from saved_model.pb and that has no
    y3_config.json (see yolo3_b200/weights.py); there is no silent default."""
    return LoadedModel(saved_model_filepath, max_batch=max_batch, device=device, anchors=anchors)
