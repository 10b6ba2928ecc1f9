"""torch-CPU restatement of the reference network forward + decode (ORACLE - tests only).

PARITY UNPINNED: the reference computes this inside TensorFlow 2.x / Keras
(unpinned, not vendored, not installable here), so there is nothing to run it
against.  This file restates /root/reference/model.py:
  conv_layer                 29-39    Conv2D(+bias, SAME) -> leaky_relu(0.2) -> BatchNorm(eps 1e-3)   (Q1-Q4)
  feature_block              42-48    layer = X + conv3(conv1(layer)), X = block input             (Q5)
  yolo_block                 51-59
  upsample_2x                94-105   Conv2DTranspose k2 s2 (kernel from the weights; ones at init) (Q6)
  detection_layer            108-120  linear 1x1 conv, A*(5+NC) channels
  reorg_layer                122-167  decode
  convert_feature_map_...    169-212  sigmoid(obj/cls), centre -> corners, concat scales 32,16,8  (Q9)
  build_feature_maps         356-380  bridges keep the route's width, concat [up, route]          (Q7)
  darknet53_feature_extractor 383-421
Keras layer auto-names follow creation order: conv2d, conv2d_1 ... conv2d_71,
batch_normalization ... _71, conv2d_transpose, conv2d_transpose_1, feature_map_1/2/3.
Weight layouts are Keras': Conv2D kernel [kh,kw,Cin,Cout]; Conv2DTranspose kernel [kh,kw,Cout,Cin].
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LEAKY = 0.2
BN_EPS = 1e-3
DEFAULT_ANCHORS = [(32, 32), (128, 128), (256, 256)]


def _suffix(k):
    return "" if k == 0 else "_%d" % k


class _Namer:
    def __init__(self):
        self.conv = 0
        self.bn = 0
        self.convt = 0

    def next_conv(self):
        n = "conv2d" + _suffix(self.conv)
        self.conv += 1
        return n

    def next_bn(self):
        n = "batch_normalization" + _suffix(self.bn)
        self.bn += 1
        return n

    def next_convt(self):
        n = "conv2d_transpose" + _suffix(self.convt)
        self.convt += 1
        return n


def layer_table(c_img, num_classes, num_anchors):
    """Flat list of every weighted layer in creation order:
    dicts(kind in {conv,det,convt}, name, bn, cin, cout, k, stride)."""
    nm = _Namer()
    tab = []

    def conv(cin, cout, k, s=1):
        tab.append(dict(kind="conv", name=nm.next_conv(), bn=nm.next_bn(), cin=cin, cout=cout, k=k, stride=s))
        return cout

    def block(c, reps):
        for _ in range(reps):
            conv(c, c // 2, 1)
            conv(c // 2, c, 3)

    def yolo(cin, f):
        conv(cin, f // 2, 1)
        conv(f // 2, f, 3)
        conv(f, f // 2, 1)
        conv(f // 2, f, 3)
        conv(f, f // 2, 1)
        conv(f // 2, f, 3)

    det_c = num_anchors * (5 + num_classes)
    conv(c_img, 32, 3)
    conv(32, 64, 3, 2)
    block(64, 1)
    conv(64, 128, 3, 2)
    block(128, 2)
    conv(128, 256, 3, 2)
    block(256, 8)
    conv(256, 512, 3, 2)
    block(512, 8)
    conv(512, 1024, 3, 2)
    block(1024, 4)
    yolo(1024, 1024)
    tab.append(dict(kind="det", name="feature_map_1", cin=1024, cout=det_c, k=1, stride=1))
    conv(512, 512, 1)
    tab.append(dict(kind="convt", name=nm.next_convt(), cin=512, cout=512, k=2, stride=2))
    yolo(1024, 512)
    tab.append(dict(kind="det", name="feature_map_2", cin=512, cout=det_c, k=1, stride=1))
    conv(256, 256, 1)
    tab.append(dict(kind="convt", name=nm.next_convt(), cin=256, cout=256, k=2, stride=2))
    yolo(512, 256)
    tab.append(dict(kind="det", name="feature_map_3", cin=256, cout=det_c, k=1, stride=1))
    return tab


def conv_flops_per_image(h, w, c_img, num_classes, num_anchors):
    """Sum of 2*M*N*K over the 75 Conv2D (ConvT excluded) - SURVEY 8(d)."""
    res = {}
    g = (h // 32, w // 32)
    # output resolution of each layer, walked in creation order
    tab = layer_table(c_img, num_classes, num_anchors)
    cur = (h, w)
    total = 0
    scale_after = {"feature_map_1": None}
    stage_res = None
    for L in tab:
        if L["kind"] == "convt":
            cur = (cur[0] * 2, cur[1] * 2)
            continue
        if L["stride"] == 2:
            cur = (cur[0] // 2, cur[1] // 2)
        total += 2 * cur[0] * cur[1] * L["cout"] * L["k"] * L["k"] * L["cin"]
    return total


def init_weights(c_img, num_classes, num_anchors, seed=0, randomize_bn=False, obj_bias=None,
                 head_gain=1.0):
    """Keras-default random init (glorot-uniform kernels, zero bias, BN identity, ConvT ones),
    optionally with randomised BatchNorm statistics so the (s,t) epilogue is exercised, and an
    objectness-bias shift on the detection layers for a sparse detection regime."""
    g = torch.Generator().manual_seed(seed)
    W = {}
    for L in layer_table(c_img, num_classes, num_anchors):
        k, cin, cout = L["k"], L["cin"], L["cout"]
        if L["kind"] == "convt":
            W[L["name"] + "/kernel"] = torch.ones(k, k, cout, cin)
            W[L["name"] + "/bias"] = torch.zeros(cout)
            continue
        lim = math.sqrt(6.0 / (k * k * cin + k * k * cout))
        W[L["name"] + "/kernel"] = (torch.rand(k, k, cin, cout, generator=g) * 2 - 1) * lim
        W[L["name"] + "/bias"] = torch.zeros(cout)
        if L["kind"] == "det":
            if head_gain != 1.0:
                W[L["name"] + "/kernel"] *= head_gain
            if obj_bias is not None:
                b = W[L["name"] + "/bias"].view(num_anchors, 5 + num_classes)
                b[:, 4] = obj_bias
            continue
        bn = L["bn"]
        if randomize_bn:
            gamma = torch.rand(cout, generator=g) + 0.5
            flip = torch.rand(cout, generator=g) < 0.03
            gamma[flip] = -gamma[flip]
            W[bn + "/gamma"] = gamma
            W[bn + "/beta"] = torch.randn(cout, generator=g) * 0.1
            W[bn + "/moving_mean"] = torch.randn(cout, generator=g) * 0.1
            W[bn + "/moving_variance"] = torch.rand(cout, generator=g) + 0.5
            W[L["name"] + "/bias"] = torch.randn(cout, generator=g) * 0.05
        else:
            W[bn + "/gamma"] = torch.ones(cout)
            W[bn + "/beta"] = torch.zeros(cout)
            W[bn + "/moving_mean"] = torch.zeros(cout)
            W[bn + "/moving_variance"] = torch.ones(cout)
    return {k: v.contiguous() for k, v in W.items()}


class OracleNet:
    """Callable restatement of YoloV3.model / model_feature_maps."""

    def __init__(self, weights, img_size, num_classes, anchors=None, dtype=torch.float32,
                 round_activations=None):
        self.H, self.W, self.C = int(img_size[0]), int(img_size[1]), int(img_size[2])
        self.nc = int(num_classes)
        self.anchors = [tuple(a) for a in (anchors if anchors is not None else DEFAULT_ANCHORS)]
        self.dtype = dtype
        self.w = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in weights.items()}
        # optional: emulate bf16 storage of every activation (used to size tolerances)
        self.round = round_activations
        self._it = None
        self.trace = None          # set to {} to record every layer output by Keras name

    # -- layers ------------------------------------------------------------
    def _q(self, x):
        return x if self.round is None else x.to(self.round).to(self.dtype)

    def _conv_layer(self, x, L):
        k, s = L["k"], L["stride"]
        w = self.w[L["name"] + "/kernel"].permute(3, 2, 0, 1)
        b = self.w[L["name"] + "/bias"]
        if k == 3:
            # TF SAME: stride 1 -> (1,1); stride 2 on an even extent -> (0 before, 1 after)  (Q4)
            x = F.pad(x, (1, 1, 1, 1)) if s == 1 else F.pad(x, (0, 1, 0, 1))
        z = F.conv2d(x, w, b, stride=s)
        a = F.leaky_relu(z, LEAKY)
        bn = L["bn"]
        sc = self.w[bn + "/gamma"] / torch.sqrt(self.w[bn + "/moving_variance"] + BN_EPS)
        sh = self.w[bn + "/beta"] - self.w[bn + "/moving_mean"] * sc
        return a * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)

    def _next(self, kind):
        L = next(self._it)
        assert L["kind"] == kind, (L, kind)
        return L

    def _rec(self, L, y):
        if self.trace is not None:
            self.trace[L["name"]] = y
        return y

    def _cl(self, x):
        L = self._next("conv")
        return self._rec(L, self._q(self._conv_layer(x, L)))

    def _block(self, x, reps):
        y = x
        for _ in range(reps):
            y = self._cl(y)
            L = self._next("conv")
            y = self._rec(L, self._q(x + self._conv_layer(y, L)))       # Q5: adds the BLOCK input
        return y

    def _yolo(self, x):
        for _ in range(5):
            x = self._cl(x)
        return x, self._cl(x)

    def _det(self, x):
        L = self._next("det")
        w = self.w[L["name"] + "/kernel"].permute(3, 2, 0, 1)
        return F.conv2d(x, w, self.w[L["name"] + "/bias"])

    def _up(self, x):
        L = self._next("convt")
        w = self.w[L["name"] + "/kernel"].permute(3, 2, 0, 1)            # [Cin, Cout, kh, kw]
        return self._rec(L, self._q(F.conv_transpose2d(x, w, self.w[L["name"] + "/bias"], stride=2)))

    # -- graph -------------------------------------------------------------
    @torch.no_grad()
    def feature_maps(self, x):
        """[B,C,H,W] -> (fm1 [B,A(5+NC),H/32,W/32], fm2 [.. /16], fm3 [.. /8])  NCHW."""
        x = torch.as_tensor(np.asarray(x)).to(self.dtype)
        self._it = iter(layer_table(self.C, self.nc, len(self.anchors)))
        x = self._q(x)
        x = self._cl(x)
        x = self._cl(x)
        x = self._block(x, 1)
        x = self._cl(x)
        x = self._block(x, 2)
        x = self._cl(x)
        r1 = x = self._block(x, 8)
        x = self._cl(x)
        r2 = x = self._block(x, 8)
        x = self._cl(x)
        x = self._block(x, 4)
        route, x = self._yolo(x)
        fm1 = self._det(x)
        x = self._up(self._cl(route))
        route, x = self._yolo(torch.cat([x, r2], 1))
        fm2 = self._det(x)
        x = self._up(self._cl(route))
        route, x = self._yolo(torch.cat([x, r1], 1))
        fm3 = self._det(x)
        assert next(self._it, None) is None
        return fm1, fm2, fm3

    @torch.no_grad()
    def decode(self, fms):
        """model.py:122-212.  -> [B, N, 5+NC], rows (i*gw+j)*A+a, scales 32,16,8."""
        A, nc = len(self.anchors), self.nc
        anc = torch.tensor(self.anchors, dtype=self.dtype)
        out = []
        for fm in fms:
            fm = torch.as_tensor(np.asarray(fm)).to(self.dtype)
            B, _, gh, gw = fm.shape
            # model.py:127 - the (h, w) stride pair multiplies the (x, y) pair as written
            stride = torch.tensor([float(self.H // gh), float(self.W // gw)], dtype=self.dtype)
            t = fm.permute(0, 2, 3, 1).reshape(B, gh, gw, A, 5 + nc)
            gy, gx = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
            off = torch.stack([gx, gy], -1).view(gh, gw, 1, 2).to(self.dtype)
            xy = (torch.sigmoid(t[..., 0:2]) + off) * stride
            wh = torch.exp(t[..., 2:4]) * anc
            obj = torch.sigmoid(t[..., 4:5])
            cls = torch.sigmoid(t[..., 5:])
            half = wh / 2.0
            rows = torch.cat([xy - half, xy + half, obj, cls], -1)
            out.append(rows.reshape(B, gh * gw * A, 5 + nc))
        return torch.cat(out, 1)

    def __call__(self, x, training=False):
        return self.decode(self.feature_maps(x)).to(torch.float32).numpy()


def heads_rel_err(a, b):
    """max|a-b| / max|b| for one head (the north_star tolerance metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
