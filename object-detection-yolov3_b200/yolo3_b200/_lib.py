"""ctypes binding of libyolo3_b200.so (include/yolo3_b200.h).

The library is the product: there is NO Python / NumPy / CPU fallback.  If the shared object is
missing or no sm_100 device is present, every compute call raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libyolo3_b200.so")

Y3_MAX_ANCHORS = 8
MEM_HOST, MEM_DEVICE = 0, 1
U8, U16, I32, F32 = 0, 1, 2, 3
ABI_VERSION = 3            # include/yolo3_b200.h Y3_ABI_VERSION this binding was written against
OK, ERR_INVALID, ERR_CUDA, ERR_NOSPACE, ERR_STATE, ERR_UNSUPPORTED, ERR_NODEVICE = 0, -1, -2, -3, -4, -5, -6


class Y3Config(ctypes.Structure):
    _fields_ = [("struct_size", c_int32), ("img_h", c_int32), ("img_w", c_int32), ("img_c", c_int32),
                ("num_classes", c_int32), ("num_anchors", c_int32),
                ("anchors", (c_float * 2) * Y3_MAX_ANCHORS), ("max_batch", c_int32), ("device", c_int32),
                ("max_candidates", c_int64)]


class Y3Timings(ctypes.Structure):
    _fields_ = [("ms_total", c_float), ("ms_h2d", c_float), ("ms_prep", c_float), ("ms_conv", c_float),
                ("ms_decode", c_float), ("ms_nms", c_float), ("ms_stitch", c_float), ("ms_d2h", c_float),
                ("ms_comm", c_float), ("reserved_", c_float),
                ("kernels_launched", c_int64), ("candidates", c_int64), ("kept", c_int64), ("tiles", c_int64)]


# name -> (restype, argtypes); exactly the symbols include/yolo3_b200.h declares
PROTOTYPES = {
    "y3_abi_version": (c_int32, []),
    "y3_last_error": (c_char_p, [c_void_p]),
    "y3_create": (c_int32, [POINTER(Y3Config), POINTER(c_void_p)]),
    "y3_destroy": (None, [c_void_p]),
    "y3_load_weights": (c_int32, [c_void_p, c_int32, POINTER(c_char_p), POINTER(c_void_p)]),
    "y3_forward_heads": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32]),
    "y3_forward_boxes": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32]),
    "y3_boxes_per_image": (c_int64, [c_void_p]),
    "y3_detect": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_float, c_float, c_float, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_int64, POINTER(c_int64)]),
    "y3_detect_image": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_float, c_int32,
                                  c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_int64)]),
    "y3_compute_iou": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "y3_filter_small": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p, c_int64,
                                  POINTER(c_int64)]),
    "y3_single_class_nms": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p, POINTER(c_int64)]),
    "y3_per_class_nms": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_int64)]),
    "y3_tile_plan": (c_int64, [c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64]),
    "y3_batch_plan": (c_int64, [c_int64, c_int32, c_int32, c_void_p, c_int64]),
    "y3_tiles_normalized": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_int32, c_int32,
                                      c_int32, c_int32, c_int64, c_int64, c_void_p, c_int32]),
    "y3_zscore": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_int32]),
    "y3_tiles_raw": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_int32, c_int32,
                               c_int32, c_int32, c_int64, c_int64, c_void_p, c_int32]),
    "y3_stitch_tiles": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int64, c_int64, c_int32, c_int32,
                                  c_int32, c_int64, c_int64, c_float, c_float, c_float, c_void_p, c_int32, c_int64,
                                  POINTER(c_int64)]),
    "y3_infer_tiled": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_int32, c_int32, c_int32,
                                 c_int32, c_int64, c_int64, c_float, c_float, c_float, c_void_p, c_int32, c_int64,
                                 POINTER(c_int64)]),
    "y3_host_alloc": (c_int32, [c_int64, POINTER(c_void_p)]),
    "y3_host_free": (None, [c_void_p]),
    "y3_cross_seam_nms": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int64, c_int64, c_int32, c_int32, c_int32,
                                    c_float, c_void_p, c_int32, c_int64, POINTER(c_int64)]),
    "y3_comm_unique_id": (c_int32, [c_void_p]),
    "y3_comm_init": (c_int32, [c_void_p, c_int32, c_int32, c_void_p]),
    "y3_comm_size": (c_int32, [c_void_p]),
    "y3_infer_tiled_sharded": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_int32, c_int32, c_int32,
                                         c_int32, c_float, c_float, c_float, c_int32, c_void_p, c_int32, c_int64,
                                         POINTER(c_int64)]),
    "y3_get_timings": (c_int32, [c_void_p, POINTER(Y3Timings)]),
    "y3_bench_forward": (c_int32, [c_void_p, c_int32, c_int32, POINTER(c_float)]),
    "y3_profile_layers": (c_int32, [c_void_p, c_int32, c_int32, c_char_p, c_int64]),
    "y3_debug_layer_output": (c_int32, [c_void_p, c_char_p, c_int32, c_void_p, c_int64, POINTER(c_int32)]),
}

# include/yolo3_b200_probe.h - the hardware probes of tests/probe_*.py live in their own shared object
PROBE_LIB_PATH = os.path.join(os.path.dirname(_HERE), "libyolo3_b200_probe.so")
PROBE_PROTOTYPES = {
    "y3_debug_umma_rowshift": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "y3_debug_im2col": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  c_void_p, c_int32, c_void_p]),
}

_lib = None
_probe = None


def load_probe():
    """libyolo3_b200_probe.so (test infrastructure): takes handles created by the main library."""
    global _probe
    if _probe is None:
        load()
        if not os.path.exists(PROBE_LIB_PATH):
            raise RuntimeError("libyolo3_b200_probe.so not built - run `make -C object-detection-yolov3_b200 all`")
        lib = ctypes.CDLL(PROBE_LIB_PATH)
        for name, (res, args) in PROBE_PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _probe = lib
    return _probe


def load():
    """dlopen the library and bind every prototype.  Raises if the .so has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libyolo3_b200.so not built (%s) - run `python __graft_entry__.py` or "
                           "`make -C object-detection-yolov3_b200`; there is no fallback path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.y3_abi_version() < ABI_VERSION:
        raise RuntimeError("%s has ABI version %d, this package needs >= %d - rebuild it (`python __graft_entry__.py`)"
                           % (LIB_PATH, lib.y3_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


class Y3Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libyolo3_b200 error %d: %s" % (code, msg))
        self.code = code


def check(status, handle=None):
    if status != OK:
        msg = load().y3_last_error(handle)
        raise Y3Error(status, msg.decode("utf-8", "replace") if msg else "?")
