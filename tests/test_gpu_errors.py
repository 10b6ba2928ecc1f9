"""Error behaviour of the C ABI (include/yolo3_b200.h): negative y3_status + y3_last_error message, surfaced as
RuntimeError (Y3Error) by the Python facades like the reference's own RuntimeErrors; the handle stays usable."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SIZE = (96, 128, 3)


@pytest.fixture(scope="module")
def eng():
    from yolo3_b200 import Engine, weights
    e = Engine(SIZE, 2, None, max_batch=2)
    e.load_weights(weights.random_init(3, 2, 3, seed=0, randomize_bn=True))
    return e


def test_forward_before_weights_is_a_state_error():
    from yolo3_b200 import Engine, Y3Error, _lib
    e = Engine(SIZE, 2, None, max_batch=1)
    with pytest.raises(Y3Error) as ex:
        e.forward_heads(np.zeros((1, 3, 96, 128), np.float32))
    assert ex.value.code == _lib.ERR_STATE and "weights" in str(ex.value)


def test_batch_and_shape_checks(eng):
    from yolo3_b200 import Y3Error, _lib
    with pytest.raises(Y3Error) as ex:
        eng.forward_heads(np.zeros((3, 3, 96, 128), np.float32))          # max_batch is 2
    assert ex.value.code == _lib.ERR_INVALID and "batch" in str(ex.value)
    with pytest.raises(Y3Error) as ex:
        eng.infer_tiled(np.zeros((300, 300, 3), np.uint8), (64, 64), 8, edge_range=0)   # tile != network input
    assert ex.value.code == _lib.ERR_INVALID and "tile" in str(ex.value)
    with pytest.raises(Y3Error) as ex:
        eng.load_weights({"conv2d/kernel": np.zeros((3, 3, 3, 31), np.float32)})          # wrong Cout
    assert ex.value.code == _lib.ERR_INVALID
    with pytest.raises(Y3Error):
        eng.load_weights({"no_such_layer/kernel": np.zeros((1, 1, 1, 1), np.float32)})
    # the handle is still good after the errors
    out = eng.forward_heads(np.zeros((2, 3, 96, 128), np.float32))
    assert out[0].shape == (2, 21, 3, 4) and np.isfinite(out[0]).all()


def test_unsupported_configurations_fail_at_create():
    from yolo3_b200 import Engine, Y3Error, _lib
    with pytest.raises(Y3Error) as ex:
        Engine((96, 128, 3), 80, [(10, 10)] * 4, max_batch=1)             # 4 x 85 = 340 head channels > 256
    assert ex.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(Y3Error):
        Engine((100, 128, 3), 1, None, max_batch=1)                        # H not a multiple of 32


def test_output_capacity_is_reported_not_overrun(eng):
    """caller-allocated outputs: a too small capacity returns Y3_ERR_NOSPACE and the required size"""
    from yolo3_b200 import _lib
    rng = np.random.default_rng(0)
    c = rng.uniform(0, 500, (300, 2)); wh = rng.uniform(5, 10, (300, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    obj = np.ones(300, np.float32); cls = np.ones((300, 1), np.float32)
    cap = 4
    ob = np.full((cap + 8, 4), -7.0, np.float32); os_ = np.empty(cap + 8, np.float32)
    ol = np.empty(cap + 8, np.int32); osrc = np.empty(cap + 8, np.int32)
    n = ctypes.c_int64()
    st = eng.lib.y3_per_class_nms(eng.h, boxes.ctypes.data, obj.ctypes.data, cls.ctypes.data, 300, 1, ctypes.c_float(0.3),
                                  ctypes.c_float(0.1), ob.ctypes.data, os_.ctypes.data, ol.ctypes.data, osrc.ctypes.data,
                                  cap, ctypes.byref(n))
    assert st == _lib.ERR_NOSPACE and n.value > cap
    assert (ob[cap:] == -7.0).all()                                          # nothing written past the capacity


def test_empty_inputs(eng):
    import bbox_utils
    assert bbox_utils.single_class_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5) == []
    b, s, l = bbox_utils.per_class_nms(np.zeros((5, 4), np.float32), np.zeros((5, 1), np.float32), np.zeros((5, 2), np.float32))
    assert b is None and s is None and l is None                             # the reference's (None, None, None)
    assert bbox_utils.filter_small_boxes(np.zeros((0, 7), np.float32), 3).shape == (0, 7)


def test_round2_entry_points_check_state_and_arguments(eng):
    """y3_infer_tiled_sharded before y3_comm_init, y3_detect_image with a foreign size, y3_cross_seam_nms with bad arguments,
    a candidate capacity overflow in the segmented pipeline: status + message, and the handle stays usable"""
    from yolo3_b200 import Engine, Y3Error, _lib, weights
    img = np.zeros((200, 230, 3), np.uint8)
    with pytest.raises(Y3Error) as ex:
        eng.infer_tiled_sharded(img, (96, 128), 8, edge_range=32)
    assert ex.value.code == _lib.ERR_STATE and "y3_comm_init" in str(ex.value)
    with pytest.raises(Y3Error) as ex:
        eng.detect_image(np.zeros((64, 64, 3), np.uint8))
    assert ex.value.code == _lib.ERR_INVALID and "network input" in str(ex.value)
    n = ctypes.c_int64()
    st = eng.lib.y3_cross_seam_nms(eng.h, None, 0, 5, 2, 100, 100, 32, 32, 0, ctypes.c_float(0.3), None, 0, 0, ctypes.byref(n))
    assert st == _lib.ERR_INVALID
    assert eng.lib.y3_comm_unique_id(None) == _lib.ERR_INVALID
    p = ctypes.c_void_p()
    assert eng.lib.y3_host_alloc(0, ctypes.byref(p)) == _lib.ERR_INVALID
    # candidate list overflow: every (row, class) passes but the handle was created with room for 100 candidates
    small = Engine(SIZE, 2, None, max_batch=1, max_candidates=100)
    w = weights.random_init(3, 2, 3, seed=0, randomize_bn=True)
    for i in (1, 2, 3):
        b = np.zeros((3, 7), np.float32)
        b[:, 4:] = 9.0                                   # objectness and class logits far above the threshold
        w["feature_map_%d/bias" % i] = b.reshape(-1)
        w["feature_map_%d/kernel" % i] = np.zeros_like(w["feature_map_%d/kernel" % i])
    small.load_weights(w)
    with pytest.raises(Y3Error) as ex:
        small.detect(np.zeros((1, 3, 96, 128), np.float32), 0, 0.3, 0.1)
    assert ex.value.code == _lib.ERR_NOSPACE and "max_candidates" in str(ex.value)
    with pytest.raises(Y3Error) as ex:
        small.infer_tiled(np.zeros((96, 128, 3), np.uint8), (96, 128), 0, edge_range=32)
    assert ex.value.code == _lib.ERR_NOSPACE
    # the first handle still works
    assert eng.forward_heads(np.zeros((1, 3, 96, 128), np.float32))[0].shape == (1, 21, 3, 4)
    b, s, l = eng.detect_image(np.zeros(SIZE, np.uint8))
    assert b.shape[1] == 4 and len(s) == len(l) == len(b)
