"""Host-side logic that needs no GPU."""
import numpy as np
import torch

from yolo3_b200 import _lib
from yolo3_b200.engine import _image_arg


def test_image_dtypes_are_converted_like_the_reference():
    """inference_tiled.py:202 accepts any dtype via astype(np.float32); the library reads u8/u16/i32/f32, everything
    else is converted on the host - exactly where that is possible (int16 used to be reinterpreted as uint16)."""
    for dt, want in ((np.uint8, _lib.U8), (np.uint16, _lib.U16), (np.int32, _lib.I32), (np.float32, _lib.F32)):
        a = np.zeros((4, 4, 1), dt)
        b, code = _image_arg(a)
        assert b is a and code == want
    a = np.array([[-5, 7], [300, -32768]], np.int16).reshape(2, 2, 1)
    b, code = _image_arg(a)
    assert code == _lib.I32 and b.dtype == np.int32 and np.array_equal(b, a)
    for dt in (np.float64, np.uint32, np.int64):
        a = (np.arange(12).reshape(3, 4, 1) * 1000).astype(dt)
        b, code = _image_arg(a)
        assert code == _lib.F32 and np.array_equal(b, a.astype(np.float32))
    b, code = _image_arg(np.ones((2, 2, 1), bool))
    assert code == _lib.I32 and b.sum() == 4
    t = torch.tensor([[-3, 9]], dtype=torch.int16).reshape(1, 2, 1)
    b, code = _image_arg(t)
    assert code == _lib.I32 and b.dtype == torch.int32 and b.flatten().tolist() == [-3, 9]
    b, code = _image_arg(torch.zeros(2, 2, 1, dtype=torch.float64))
    assert code == _lib.F32 and b.dtype == torch.float32
    # non-contiguous input is made contiguous
    a = np.zeros((4, 6, 2), np.uint16)[:, ::2]
    b, _ = _image_arg(a)
    assert b.flags["C_CONTIGUOUS"] and b.shape == (4, 3, 2)
