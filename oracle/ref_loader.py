"""Import the reference's NumPy code VERBATIM from /root/reference.

Build-container only: /root/reference does not exist on the GPU box, so nothing
in `-m gpu` tests, smoke() or bench.py may call this.  It is used by
`oracle/gen_golden.py` (to mint `tests/golden/*.npz`) and by the `not gpu`
tests that pin the restatements in this directory to the reference
(those tests skip when /root/reference is absent).

tensorflow / skimage / lmdb / isg_ai_pb2 are not installed, so they are
stubbed in sys.modules exactly far enough for the modules to import
(SURVEY.md section 8c).  No reference source is copied.
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("Y3_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "bbox_utils.py"))


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__y3_stub__", False):
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__y3_stub__ = True
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns dict(bbox_utils=..., inference_tiled=..., imagereader=...)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)

    tf = _stub("tensorflow", __version__="2.3.0", function=lambda f=None, **k: f,
               convert_to_tensor=lambda x: x)
    tf.nn = types.SimpleNamespace(leaky_relu=None, sigmoid=None)
    tf.keras = types.SimpleNamespace()
    sk = _stub("skimage")
    sk.io = _stub("skimage.io", imread=None)
    sk.measure = _stub("skimage.measure")
    sk.transform = _stub("skimage.transform")
    _stub("lmdb")
    _stub("isg_ai_pb2", ImageYoloBoxesPair=object)
    _stub("scipy.ndimage") if "scipy.ndimage" not in sys.modules else None

    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  ("bbox_utils", "inference_tiled", "imagereader", "model", "augment")}
    # a stand-in for reference model.py (it builds Keras graphs at import of
    # inference_tiled only through `model.YoloV3.NETWORK_DOWNSAMPLE_FACTOR`)
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, REFERENCE_DIR)
    try:
        fake_model = types.ModuleType("model")
        fake_model.YoloV3 = type("YoloV3", (), {"NETWORK_DOWNSAMPLE_FACTOR": 32})
        sys.modules["model"] = fake_model
        fake_aug = types.ModuleType("augment")
        sys.modules["augment"] = fake_aug
        out = {}
        for name in ("bbox_utils", "imagereader", "inference_tiled"):
            out[name] = importlib.import_module(name)
            assert os.path.dirname(out[name].__file__) == REFERENCE_DIR
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
        for k in ("tensorflow", "skimage", "skimage.io", "skimage.measure",
                  "skimage.transform", "lmdb", "isg_ai_pb2"):
            if getattr(sys.modules.get(k), "__y3_stub__", False):
                del sys.modules[k]
    _loaded.update(out)
    return _loaded
