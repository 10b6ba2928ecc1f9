"""Host-side wrapper of one y3_handle: construction, DLPack weight hand-over, forward / detect /
tiled inference.  Pure plumbing - all arithmetic happens in libyolo3_b200.so on the GPU."""
import ctypes

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, Y3Config, Y3Timings, check

DEFAULT_ANCHORS = [(32, 32), (128, 128), (256, 256)]          # model.py:433
_DTYPES = {np.dtype(np.uint8): _lib.U8, np.dtype(np.uint16): _lib.U16, np.dtype(np.int32): _lib.I32,
           np.dtype(np.float32): _lib.F32}
_USED = b"used_dltensor"

_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_PyCapsule_SetName = ctypes.pythonapi.PyCapsule_SetName
_PyCapsule_SetName.restype = ctypes.c_int
_PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]


def _ptr(a):
    """(address, y3_mem) of a numpy array or a torch tensor."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, MEM_HOST
    if hasattr(a, "data_ptr"):                                  # torch.Tensor
        assert a.is_contiguous()
        return a.data_ptr(), (MEM_DEVICE if a.is_cuda else MEM_HOST)
    raise TypeError(type(a))


def _image_arg(img):
    """(array, y3_dtype) for an HWC image the tiling entry points accept.  The library reads u8 / u16 / i32 / f32;
    anything else is converted the way the reference does for every tile (`astype(np.float32)`, inference_tiled.py:202),
    or exactly to int32 where that is lossless (int8 / int16 / bool)."""
    if isinstance(img, np.ndarray):
        if img.dtype not in _DTYPES:
            img = img.astype(np.int32 if img.dtype in (np.dtype(np.int8), np.dtype(np.int16), np.dtype(np.bool_)) else np.float32)
        return np.ascontiguousarray(img), _DTYPES[img.dtype]
    import torch
    table = {torch.uint8: _lib.U8, torch.int32: _lib.I32, torch.float32: _lib.F32}
    if getattr(torch, "uint16", None) is not None:
        table[torch.uint16] = _lib.U16
    if img.dtype not in table:
        img = img.to(torch.int32 if img.dtype in (torch.int8, torch.int16, torch.bool) else torch.float32)
    return img.contiguous(), table[img.dtype]


def _capsule(t):
    if isinstance(t, np.ndarray):
        return np.ascontiguousarray(t, dtype=np.float32).__dlpack__()
    import torch
    from torch.utils.dlpack import to_dlpack
    return to_dlpack(t.detach().to(torch.float32).contiguous())


class Engine:
    """One library handle.  img_size=None creates a post-processing-only handle."""

    def __init__(self, img_size=None, number_classes=1, anchors=None, max_batch=1, device=0, max_candidates=0):
        self.lib = _lib.load()
        cfg = Y3Config()
        cfg.struct_size = ctypes.sizeof(Y3Config)
        anchors = list(anchors) if anchors is not None else list(DEFAULT_ANCHORS)
        if img_size is not None:
            cfg.img_h, cfg.img_w, cfg.img_c = int(img_size[0]), int(img_size[1]), int(img_size[2])
        cfg.num_classes = int(number_classes)
        cfg.num_anchors = len(anchors)
        for i, (w, h) in enumerate(anchors):
            cfg.anchors[i][0], cfg.anchors[i][1] = float(w), float(h)
        cfg.max_batch = int(max_batch)
        cfg.device = int(device)
        cfg.max_candidates = int(max_candidates)
        self.img_size = None if img_size is None else tuple(int(v) for v in img_size)
        self.number_classes, self.anchors, self.max_batch, self.device = int(number_classes), anchors, int(max_batch), int(device)
        h = ctypes.c_void_p()
        check(self.lib.y3_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.y3_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights (DLPack)
    def load_weights(self, weights):
        """weights: {keras_variable_name: fp32 array (numpy or torch, host or CUDA)} in Keras layouts."""
        names = list(weights.keys())
        caps = [_capsule(weights[k]) for k in names]
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (ctypes.c_void_p * n)(*[_PyCapsule_GetPointer(c, b"dltensor") for c in caps])
        check(self.lib.y3_load_weights(self.h, n, c_names, c_ptrs), self.h)
        for c in caps:                                          # consumed: the library ran the deleters
            _PyCapsule_SetName(c, _USED)

    # ------------------------------------------------------------------ network
    @property
    def boxes_per_image(self):
        return int(self.lib.y3_boxes_per_image(self.h))

    def forward_heads(self, batch):
        """NCHW fp32 [B,C,H,W] -> three NCHW fp32 heads (model.py:462 get_keras_feature_map_model)."""
        batch = np.ascontiguousarray(batch, dtype=np.float32) if isinstance(batch, np.ndarray) else batch
        B = batch.shape[0]
        H, W, _ = self.img_size
        dc = len(self.anchors) * (5 + self.number_classes)
        outs = [np.empty((B, dc, H // s, W // s), np.float32) for s in (32, 16, 8)]
        p, mem = _ptr(batch)
        check(self.lib.y3_forward_heads(self.h, p, mem, B, outs[0].ctypes.data, outs[1].ctypes.data,
                                        outs[2].ctypes.data, MEM_HOST), self.h)
        return outs

    def forward_boxes(self, batch):
        """yolo_model(batch, training=False): [B,C,H,W] -> [B,N,5+NC] fp32."""
        batch = np.ascontiguousarray(batch, dtype=np.float32) if isinstance(batch, np.ndarray) else batch
        B = batch.shape[0]
        out = np.empty((B, self.boxes_per_image, 5 + self.number_classes), np.float32)
        p, mem = _ptr(batch)
        check(self.lib.y3_forward_boxes(self.h, p, mem, B, out.ctypes.data, MEM_HOST), self.h)
        return out

    def __call__(self, batch, training=False):
        return self.forward_boxes(np.asarray(batch) if not hasattr(batch, "data_ptr") else batch)

    def detect(self, batch, min_box_size=32, iou_threshold=0.3, score_threshold=0.1, cap=None):
        """forward -> filter_small_boxes -> per_class_nms on the device.  -> boxes, scores, labels, img_index."""
        batch = np.ascontiguousarray(batch, dtype=np.float32) if isinstance(batch, np.ndarray) else batch
        B = batch.shape[0]
        cap = int(cap) if cap else getattr(self, "_detect_cap", 1 << 16)
        p, mem = _ptr(batch)
        while True:
            ob = np.empty((cap, 4), np.float32)
            os_ = np.empty(cap, np.float32)
            ol = np.empty(cap, np.int32)
            oi = np.empty(cap, np.int32)
            n = ctypes.c_int64()
            st = self.lib.y3_detect(self.h, p, mem, B, float(min_box_size), float(iou_threshold), float(score_threshold),
                                    ob.ctypes.data, os_.ctypes.data, ol.ctypes.data, oi.ctypes.data, cap, ctypes.byref(n))
            if st == _lib.ERR_NOSPACE and n.value > cap:
                cap = self._detect_cap = 2 * int(n.value)      # sticky: later calls do not run twice
                continue
            check(st, self.h)
            k = n.value
            return ob[:k], os_[:k], ol[:k], oi[:k]

    def detect_image(self, img, min_box_size=32, iou_threshold=0.3, score_threshold=0.1, clip=True, cap=None):
        """inference.py:47-79 for one HWC image of the network's size in ONE library call: whole-image z-score, forward
        (CUDA graph), decode, clip to the image, small-box filter, per-class NMS.  -> boxes [n,4] f32, scores, labels."""
        H, W, C = (int(v) for v in img.shape)
        img, dt = _image_arg(img)
        p, mem = _ptr(img)
        cap = int(cap) if cap else 1 << 14
        while True:
            ob = np.empty((cap, 4), np.float32)
            os_ = np.empty(cap, np.float32)
            ol = np.empty(cap, np.int32)
            n = ctypes.c_int64()
            st = self.lib.y3_detect_image(self.h, p, dt, mem, H, W, C, float(min_box_size), float(iou_threshold), float(score_threshold),
                                          1 if clip else 0, ob.ctypes.data, os_.ctypes.data, ol.ctypes.data, cap, ctypes.byref(n))
            if st == _lib.ERR_NOSPACE and n.value > cap:
                cap = int(n.value)
                continue
            check(st, self.h)
            k = n.value
            return ob[:k], os_[:k], ol[:k]

    # ------------------------------------------------------------------ tiled
    def infer_tiled(self, img, tile_size, min_box_size=32, edge_range=96, iou_threshold=0.3, score_threshold=0.1,
                    tile_first=0, tile_count=-1, cap=None, out_device=None):
        """inference_tiled.inference_image_tiled for tiles [tile_first, tile_first+tile_count).
        img: HWC numpy (host) or torch CUDA tensor.  Returns float64 [n,6] (numpy, or a torch CUDA
        tensor when out_device is a torch device - used for the NCCL all-gather)."""
        H, W, C = (int(v) for v in img.shape)
        img, dt = _image_arg(img)
        p, mem = _ptr(img)
        cap = int(cap) if cap else getattr(self, "_tiled_cap", 1 << 16)
        while True:
            if out_device is None:
                out = _result_rows(self, cap)
                op, om = out.ctypes.data, MEM_HOST
            else:
                import torch
                out = torch.empty((cap, 6), dtype=torch.float64, device=out_device)
                op, om = out.data_ptr(), MEM_DEVICE
            n = ctypes.c_int64()
            st = self.lib.y3_infer_tiled(self.h, p, dt, mem, H, W, C, int(tile_size[0]), int(tile_size[1]), int(edge_range),
                                         int(tile_first), int(tile_count), float(min_box_size), float(iou_threshold),
                                         float(score_threshold), op, om, cap, ctypes.byref(n))
            if st == _lib.ERR_NOSPACE and n.value > cap:
                cap = self._tiled_cap = 2 * int(n.value)       # sticky: later calls do not recompute
                continue
            check(st, self.h)
            return out[:n.value]

    def tiles_normalized(self, img, tile_size, edge_range=96, first=0, count=None):
        H, W, C = (int(v) for v in img.shape)
        total = tile_count(H, W, tile_size, edge_range)
        count = total - first if count is None else count
        out = np.empty((count, C, int(tile_size[0]), int(tile_size[1])), np.float32)
        img, dt = _image_arg(img)
        p, mem = _ptr(img)
        check(self.lib.y3_tiles_normalized(self.h, p, dt, mem, H, W, C, int(tile_size[0]),
                                           int(tile_size[1]), int(edge_range), first, count, out.ctypes.data, MEM_HOST), self.h)
        return out

    def zscore(self, data):
        """imagereader.zscore_normalize for an array of any shape -> float32 array of the same shape."""
        a = np.ascontiguousarray(data)
        if a.dtype not in _DTYPES:
            a = a.astype(np.float32)
        out = np.empty(a.shape, np.float32)
        if a.size:
            check(self.lib.y3_zscore(self.h, a.ctypes.data, _DTYPES[a.dtype], MEM_HOST, a.size, out.ctypes.data, MEM_HOST), self.h)
        return out

    def tiles_raw(self, img, tile_size, edge_range=96, first=0, count=None):
        H, W, C = (int(v) for v in img.shape)
        total = tile_count(H, W, tile_size, edge_range)
        count = total - first if count is None else count
        src_dtype = img.dtype
        img, dt = _image_arg(img)
        out = np.empty((count, int(tile_size[0]), int(tile_size[1]), C), img.dtype)
        p, mem = _ptr(img)
        check(self.lib.y3_tiles_raw(self.h, p, dt, mem, H, W, C, int(tile_size[0]),
                                    int(tile_size[1]), int(edge_range), first, count, out.ctypes.data, MEM_HOST), self.h)
        return out if out.dtype == src_dtype else out.astype(src_dtype)      # tiles come back in the source dtype

    def stitch_tiles(self, dets, img_hw, tile_size, min_box_size=32, edge_range=96, iou_threshold=0.3,
                     score_threshold=0.1, first=0):
        dets = np.ascontiguousarray(dets, dtype=np.float32)
        T, N, E = dets.shape
        cap = 1 << 16
        while True:
            out = _result_rows(self, cap)
            n = ctypes.c_int64()
            st = self.lib.y3_stitch_tiles(self.h, dets.ctypes.data, MEM_HOST, N, E - 5, int(img_hw[0]), int(img_hw[1]),
                                          int(tile_size[0]), int(tile_size[1]), int(edge_range), first, T, float(min_box_size),
                                          float(iou_threshold), float(score_threshold), out.ctypes.data, MEM_HOST, cap, ctypes.byref(n))
            if st == _lib.ERR_NOSPACE and n.value > cap:
                cap = int(n.value)
                continue
            check(st, self.h)
            return out[:n.value]

    # ------------------------------------------------------------------ post-processing
    def cross_seam_nms(self, pred, img_hw, tile_size, edge_range=96, iou_threshold=0.3, number_classes=None):
        """Optional final stage (north_star; NOT in the reference, which resolves seams by centre ownership
        only, inference_tiled.py:235-254): boxes whose extent crosses a zone boundary of the tile grid are
        candidates, greedy per-class NMS runs among them ON THE DEVICE (y3_cross_seam_nms), suppressed rows
        are dropped, everything else and the row order are untouched.  pred: float64 [n,6] as returned by
        infer_tiled (numpy, or a torch CUDA tensor) -> same kind, float64 [n',6]."""
        is_np = isinstance(pred, np.ndarray) or not hasattr(pred, "data_ptr")
        if is_np:
            pred = np.ascontiguousarray(np.asarray(pred, np.float64).reshape(-1, 6))
        n = int(pred.shape[0])
        if n == 0:
            return pred
        nc = int(number_classes) if number_classes else (self.number_classes if self.img_size is not None else
                                                         int(float(pred[:, 5].max())) + 1)
        if is_np:
            out = np.empty((n, 6), np.float64)
            op, om = out.ctypes.data, MEM_HOST
        else:
            import torch
            pred = pred.contiguous()
            out = torch.empty((n, 6), dtype=torch.float64, device=pred.device)
            op, om = out.data_ptr(), MEM_DEVICE
        p, mem = _ptr(pred)
        k = ctypes.c_int64()
        check(self.lib.y3_cross_seam_nms(self.h, p, mem, n, nc, int(img_hw[0]), int(img_hw[1]), int(tile_size[0]), int(tile_size[1]),
                                         int(edge_range), float(iou_threshold), op, om, n, ctypes.byref(k)), self.h)
        return out[:k.value]

    # ------------------------------------------------------------------ multi-GPU (NCCL inside the library)
    def comm_init(self, group=None):
        """Creates the library's NCCL communicator over the ranks of a torch.distributed group: rank 0 draws the
        unique id (y3_comm_unique_id), torch.distributed only carries its 128 bytes to the other ranks."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            check(self.lib.y3_comm_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if world > 1:
            backend = dist.get_backend(group)
            carrier = ident.cuda(self.device) if backend == "nccl" else ident
            dist.broadcast(carrier, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = carrier.cpu()
        raw = (ctypes.c_uint8 * 128)(*ident.tolist())
        check(self.lib.y3_comm_init(self.h, rank, world, raw), self.h)
        self._comm_world = world

    def infer_tiled_sharded(self, img, tile_size, min_box_size=32, edge_range=96, iou_threshold=0.3, score_threshold=0.1,
                            cross_seam=False, cap=None, out_device=None):
        """y3_infer_tiled_sharded: every rank of the communicator runs its row band of tiles, the rows are
        all-gathered by the library (NCCL on the handle's stream) - every rank gets the single-GPU result."""
        H, W, C = (int(v) for v in img.shape)
        img, dt = _image_arg(img)
        p, mem = _ptr(img)
        cap = int(cap) if cap else getattr(self, "_tiled_cap", 1 << 16)
        while True:
            if out_device is None:
                out = _result_rows(self, cap)
                op, om = out.ctypes.data, MEM_HOST
            else:
                import torch
                out = torch.empty((cap, 6), dtype=torch.float64, device=out_device)
                op, om = out.data_ptr(), MEM_DEVICE
            n = ctypes.c_int64()
            st = self.lib.y3_infer_tiled_sharded(self.h, p, dt, mem, H, W, C, int(tile_size[0]), int(tile_size[1]), int(edge_range),
                                                 float(min_box_size), float(iou_threshold), float(score_threshold),
                                                 1 if cross_seam else 0, op, om, cap, ctypes.byref(n))
            if st == _lib.ERR_NOSPACE and n.value > cap:          # reported by every rank together: all of them retry
                cap = self._tiled_cap = 2 * int(n.value)
                continue
            check(st, self.h)
            return out[:n.value]

    def compute_iou(self, box, boxes):
        box = np.ascontiguousarray(box, dtype=np.float32)
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        out = np.empty(boxes.shape[0], np.float32)
        check(self.lib.y3_compute_iou(self.h, box.ctypes.data, boxes.ctypes.data, boxes.shape[0], out.ctypes.data), self.h)
        return out

    def filter_small(self, rows, min_size):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n, L = rows.shape
        out = np.empty((max(n, 1), L), np.float32)
        k = ctypes.c_int64()
        check(self.lib.y3_filter_small(self.h, rows.ctypes.data, n, L, float(min_size), out.ctypes.data, None, n,
                                       ctypes.byref(k)), self.h)
        return out[:k.value]

    def single_class_nms(self, boxes, scores, iou_threshold):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        m = boxes.shape[0]
        keep = np.empty(max(m, 1), np.int32)
        k = ctypes.c_int64()
        check(self.lib.y3_single_class_nms(self.h, boxes.ctypes.data, scores.ctypes.data, m, float(iou_threshold),
                                           keep.ctypes.data, ctypes.byref(k)), self.h)
        return keep[:k.value]

    def per_class_nms(self, boxes, objectness, class_probs, iou_threshold=0.3, score_threshold=0.1, with_src=False):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        obj = np.ascontiguousarray(objectness, dtype=np.float32).reshape(-1)
        cls = np.ascontiguousarray(class_probs, dtype=np.float32)
        n, nc = cls.shape
        cap = max(min(n * nc, 1 << 20), 1)
        while True:
            ob = np.empty((cap, 4), np.float32)
            os_ = np.empty(cap, np.float32)
            ol = np.empty(cap, np.int32)
            osrc = np.empty(cap, np.int32)
            k = ctypes.c_int64()
            st = self.lib.y3_per_class_nms(self.h, boxes.ctypes.data, obj.ctypes.data, cls.ctypes.data, n, nc,
                                           float(iou_threshold), float(score_threshold), ob.ctypes.data, os_.ctypes.data,
                                           ol.ctypes.data, osrc.ctypes.data, cap, ctypes.byref(k))
            if st == _lib.ERR_NOSPACE and k.value > cap:
                cap = int(k.value)
                continue
            check(st, self.h)
            kk = k.value
            if kk == 0:
                return (None, None, None, None) if with_src else (None, None, None)
            res = (ob[:kk], os_[:kk], ol[:kk])
            return res + (osrc[:kk],) if with_src else res

    # ------------------------------------------------------------------ measurement
    def timings(self):
        t = Y3Timings()
        check(self.lib.y3_get_timings(self.h, ctypes.byref(t)), self.h)
        return {f: getattr(t, f) for f, _ in Y3Timings._fields_}

    def profile_layers(self, batch, iters=5):
        """Measurement hook: per-layer CSV (name, shapes, tile config, ms, TFLOP/s, minimum GB/s)."""
        buf = ctypes.create_string_buffer(1 << 16)
        check(self.lib.y3_profile_layers(self.h, int(batch), int(iters), buf, len(buf)), self.h)
        return buf.value.decode()

    def debug_layer_output(self, layer, batch):
        """Test hook: NCHW fp32 output of one layer of the last forward (needs Y3_DEBUG_NO_REUSE at create)."""
        H, W, _ = self.img_size
        cap = batch * 32 * H * W
        out = np.empty(cap, np.float32)
        dims = (ctypes.c_int32 * 3)()
        check(self.lib.y3_debug_layer_output(self.h, layer.encode(), batch, out.ctypes.data, cap, dims), self.h)
        c, h, w = dims[0], dims[1], dims[2]
        return out[:batch * c * h * w].reshape(batch, c, h, w).copy()

    def bench_forward(self, batch, iters):
        ms = ctypes.c_float()
        check(self.lib.y3_bench_forward(self.h, int(batch), int(iters), ctypes.byref(ms)), self.h)
        return ms.value


def seam_candidates(pred, img_hw, tile_size, edge_range):
    """Rows of pred [n,6] (integer pixel corners, inclusive) that straddle a zone boundary of the tile grid
    (zone = tile - 2*edge_range on every axis that is actually tiled, inference_tiled.py:41-47)."""
    pred = np.asarray(pred)
    cand = np.zeros(pred.shape[0], bool)
    for axis, (lo, hi) in enumerate(((1, 3), (0, 2))):           # axis 0 = rows (y), axis 1 = columns (x)
        if int(tile_size[axis]) >= int(img_hw[axis]):
            continue
        zone = int(tile_size[axis]) - 2 * int(edge_range)
        cand |= np.floor_divide(pred[:, lo], zone) != np.floor_divide(pred[:, hi], zone)
    return cand


def tile_count(img_h, img_w, tile_size, edge_range):
    n = _lib.load().y3_tile_plan(int(img_h), int(img_w), int(tile_size[0]), int(tile_size[1]), int(edge_range), None, None, 0)
    if n < 0:
        raise _lib.Y3Error(int(n), "bad tile geometry")
    return int(n)


def tile_plan(img_h, img_w, tile_size, edge_range):
    n = tile_count(img_h, img_w, tile_size, edge_range)
    xs = np.empty(n, np.int32)
    ys = np.empty(n, np.int32)
    _lib.load().y3_tile_plan(int(img_h), int(img_w), int(tile_size[0]), int(tile_size[1]), int(edge_range),
                             xs.ctypes.data, ys.ctypes.data, n)
    return xs, ys


def batch_plan(tile_count, max_batch, host_image):
    """tile counts of the batches y3_infer_tiled runs for tile_count tiles (y3_batch_plan; host-only)"""
    sizes = np.empty(max(1, int(tile_count)), np.int32)
    n = _lib.load().y3_batch_plan(int(tile_count), int(max_batch), 1 if host_image else 0, sizes.ctypes.data, sizes.size)
    if n < 0:
        raise ValueError("y3_batch_plan(%r, %r): bad arguments" % (tile_count, max_batch))
    return sizes[:n].tolist()


class _PinnedBlock:
    """One y3_host_alloc block, exposed through the array interface: numpy arrays made from it keep it alive
    (arr.base), and it is returned to CUDA when the last of them goes away."""

    def __init__(self, nbytes):
        self.lib = _lib.load()
        p = ctypes.c_void_p()
        check(self.lib.y3_host_alloc(int(nbytes), ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                self.lib.y3_host_free(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class _Lease:
    """Array-interface owner of a pooled page-locked block: numpy arrays made from it keep it alive; when the last of
    them goes away the block goes back to its pool instead of back to CUDA."""

    def __init__(self, block, state):
        self.block, self.state = block, state
        state["out"] += 1
        self.__array_interface__ = block.__array_interface__

    def __del__(self):
        try:
            self.state["out"] -= 1
            self.state["free"].append(self.block)
        except Exception:
            pass


_LEASES_MAX = 4


def _result_rows(owner, cap):
    """float64 [cap, 6] result buffer.  Page-locked host memory recycled between calls, so that the device-to-host copy
    of the result rows runs at link speed and touches no fresh pages (a pageable np.empty costs 0.3 ms per 4 MB of rows
    on one GPU and more when eight ranks of one host copy at once).  The caller owns the array like any other; a
    caller that keeps more than _LEASES_MAX results alive gets ordinary pageable arrays for the rest (page-locked
    memory is not for hoarding)."""
    nbytes = max(1, int(cap) * 48)
    pools = owner.__dict__.setdefault("_row_pool", {})
    for k in [k for k in pools if k != nbytes and pools[k]["out"] == 0]:       # the capacity changed: drop the old blocks
        del pools[k]
    state = pools.setdefault(nbytes, {"free": [], "out": 0})
    if not state["free"] and state["out"] >= _LEASES_MAX:
        return np.empty((int(cap), 6), np.float64)
    block = state["free"].pop() if state["free"] else _PinnedBlock(nbytes)
    return np.asarray(_Lease(block, state))[:int(cap) * 48].view(np.float64).reshape(int(cap), 6)


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (y3_host_alloc): images held in it are uploaded by y3_infer_tiled on a copy
    stream, overlapped with compute.  Raises when the library has no CUDA device (no fallback)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    block = _PinnedBlock(max(1, n * dtype.itemsize))
    return np.asarray(block)[:n * dtype.itemsize].view(dtype).reshape(shape)


def pinned_copy(a):
    """copy of a numpy array in page-locked host memory"""
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    np.copyto(out, a)
    return out


_post = {}


def post_engine(device=0):
    """Process-wide post-processing handle (bbox_utils facade)."""
    if device not in _post:
        _post[device] = Engine(img_size=None, device=device)
    return _post[device]


def shard_range(n_tiles, rank, world):
    """contiguous (row-band) tile range of one rank"""
    per = (n_tiles + world - 1) // world
    first = min(rank * per, n_tiles)
    return first, min(per, n_tiles - first)


def gather_rows(local, group=None):
    """All-gather a variable number of [n, 6] rows from every rank and concatenate them in rank order
    (counts first, then padded records) - the only exchange step of the tiled path.  Works on whatever
    device `local` lives on (NCCL for CUDA tensors, gloo for CPU tensors in the host-logic tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    return torch.cat([g[:c] for g, c in zip(gathered, counts)], 0)


def infer_tiled_distributed(eng, img, tile_size, min_box_size=32, edge_range=96, iou_threshold=0.3,
                            score_threshold=0.1, group=None, cross_seam=False, out_device="auto"):
    """inference_image_tiled with the tile grid sharded across the ranks of a torch.distributed group: every rank
    runs its row band of tiles on its own GPU (slice, normalise, conv stack, decode, NMS, ownership filter - no
    data-path collective), then the surviving boxes are all-gathered over NVLink BY THE LIBRARY (ncclAllGather on
    the handle's stream, y3_infer_tiled_sharded) and laid end to end in rank (= tile) order; cross_seam appends
    the cross-seam NMS stage on the device.  torch.distributed is used once, to hand the NCCL unique id around.
    -> float64 [n,6]: a torch CUDA tensor (out_device "auto" / a device) or numpy (out_device None)."""
    import torch
    if getattr(eng, "_comm_world", None) is None:
        eng.comm_init(group)
    dev = torch.device("cuda", eng.device) if out_device == "auto" else out_device
    return eng.infer_tiled_sharded(img, tile_size, min_box_size, edge_range, iou_threshold, score_threshold,
                                   cross_seam=cross_seam, out_device=dev)
