"""Experiment (CPU): which roundings drive the fm3 head error in the 2-anchor / 1-class config.
Emulates the GPU path (bf16 weights, bf16 activation storage, fp32 accumulate) with per-layer policies."""
import sys, os, time
import torch, numpy as np
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_torch as mt

torch.set_num_threads(os.cpu_count())

def bf(x): return x.to(torch.bfloat16).to(torch.float32)
def f16(x): return x.to(torch.float16).to(torch.float32)
def split2(x):  # hi+lo bf16 pair ~ 16 mantissa bits
    hi = bf(x); return hi + bf(x - hi)

class Emu(mt.OracleNet):
    """policy(name) -> (act_round_fn, weight_round_fn); rank1: fp32 up term"""
    def __init__(self, W, img, nc, anchors, act_policy, w_policy, rank1=False):
        super().__init__(W, img, nc, anchors)
        self.act_policy = act_policy; self.w_policy = w_policy; self.rank1 = rank1
        self.idx = 0
    def _conv_layer(self, x, L, pre=None):
        k, s = L["k"], L["stride"]
        w = self.w_policy(L["name"])(self.w[L["name"] + "/kernel"]).permute(3, 2, 0, 1)
        b = self.w[L["name"] + "/bias"]
        if k == 3:
            x = F.pad(x, (1, 1, 1, 1)) if s == 1 else F.pad(x, (0, 1, 0, 1))
        z = F.conv2d(x, w, b, stride=s)
        if pre is not None: z = z + pre
        a = F.leaky_relu(z, mt.LEAKY)
        bn = L["bn"]
        sc = self.w[bn + "/gamma"] / torch.sqrt(self.w[bn + "/moving_variance"] + mt.BN_EPS)
        sh = self.w[bn + "/beta"] - self.w[bn + "/moving_mean"] * sc
        return a * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    def _cl(self, x):
        L = self._next("conv")
        return self._rec(L, self.act_policy(L["name"])(self._conv_layer(x, L)))
    def _block(self, x, reps):
        y = x
        for _ in range(reps):
            y = self._cl(y)
            L = self._next("conv")
            y = self._rec(L, self.act_policy(L["name"])(x + self._conv_layer(y, L)))
        return y
    def _det(self, x):
        L = self._next("det")
        w = self.w_policy(L["name"])(self.w[L["name"] + "/kernel"]).permute(3, 2, 0, 1)
        return F.conv2d(x, w, self.w[L["name"] + "/bias"])
    def _upcat_conv(self, route, rskip):
        """bridge conv -> convT -> concat -> 1x1 conv, the way the GPU composes it"""
        Lb = self._next("conv")
        xb32 = self._conv_layer(route, Lb)              # fp32 bridge output
        xb = self.act_policy(Lb["name"])(xb32)
        Lt = self._next("convt")
        Lc = self._next("conv")
        Kt = self.w[Lt["name"] + "/kernel"]             # [2,2,Cup,Cx]
        Wy = self.w[Lc["name"] + "/kernel"][0, 0]       # [Cup+Cr, cout]
        cup = Kt.shape[2]
        wr = self.w_policy(Lc["name"])
        B, _, h, w = xb.shape
        cout = Wy.shape[1]
        out = torch.zeros(B, cout, 2 * h, 2 * w)
        for i in range(2):
            for j in range(2):
                Wc = torch.einsum('ok,oc->kc', Wy[:cup], Kt[i, j])     # [cout, Cx] fp32 composed
                if self.rank1:
                    # all-ones: Wc[k,c] = csum[k] for all c; up term = csum[k] * S, S = fp32 channel sum of the fp32 bridge output
                    S = xb32.sum(1, keepdim=True)
                    up = Wc[:, 0].view(1, -1, 1, 1) * S
                else:
                    up = F.conv2d(xb, wr(Wc)[:, :, None, None])
                rt = F.conv2d(rskip[:, :, i::2, j::2], wr(Wy[cup:].t().contiguous())[:, :, None, None])
                out[:, :, i::2, j::2] = up + rt
        z = out + self.w[Lc["name"] + "/bias"].view(1, -1, 1, 1)
        a = F.leaky_relu(z, mt.LEAKY)
        bn = Lc["bn"]
        sc = self.w[bn + "/gamma"] / torch.sqrt(self.w[bn + "/moving_variance"] + mt.BN_EPS)
        sh = self.w[bn + "/beta"] - self.w[bn + "/moving_mean"] * sc
        y = a * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        return self.act_policy(Lc["name"])(y)
    @torch.no_grad()
    def feature_maps(self, x):
        x = torch.as_tensor(np.asarray(x)).float()
        self._it = iter(mt.layer_table(self.C, self.nc, len(self.anchors)))
        x = bf(x)
        x = self._cl(x); x = self._cl(x); x = self._block(x, 1); x = self._cl(x); x = self._block(x, 2)
        x = self._cl(x); r1 = x = self._block(x, 8); x = self._cl(x); r2 = x = self._block(x, 8)
        x = self._cl(x); x = self._block(x, 4)
        route, x = self._yolo(x); fm1 = self._det(x)
        x = self._upcat_conv(route, r2)
        for _ in range(4): x = self._cl(x)
        route = x; x = self._cl(x); fm2 = self._det(x)
        x = self._upcat_conv(route, r1)
        for _ in range(4): x = self._cl(x)
        x = self._cl(x); fm3 = self._det(x)
        return fm1, fm2, fm3

def tail_names():
    tab = mt.layer_table(1, 1, 2)
    names = [L["name"] for L in tab]
    i2 = names.index("conv2d_transpose"); i3 = names.index("conv2d_transpose_1")
    return names, i2, i3

if __name__ == "__main__":
    img = (512, 512, 1); nc = 1; anchors = [(64, 384), (384, 64)]
    names, i2, i3 = tail_names()
    tail3 = set(names[i3 - 1:])      # bridge conv + everything after the second upsample
    tail2 = set(names[i2 - 1:])
    print("tail3", sorted(tail3))
    for seed in (1, 2, 0):
        W = mt.init_weights(1, nc, 2, seed=seed, randomize_bn=True)
        x = torch.randn(1, 1, 512, 512, generator=torch.Generator().manual_seed(2))
        want = mt.OracleNet(W, img, nc, anchors).feature_maps(x)
        def run(tag, ap, wp, rank1=False):
            t = time.time()
            got = Emu(W, img, nc, anchors, ap, wp, rank1).feature_maps(x)
            print("seed", seed, "%-44s" % tag, ["%.4f" % mt.heads_rel_err(a.numpy(), b.numpy()) for a, b in zip(got, want)], "%.1fs" % (time.time() - t), flush=True)
        allbf = lambda n: bf
        run("all bf16", allbf, allbf)
        run("rank1 fp32 up-term", allbf, allbf, True)
        run("act bf16, weights fp32", allbf, lambda n: (lambda t: t))
        run("tail3 act fp32", lambda n: (lambda t: t) if n in tail3 else bf, allbf)
        run("tail3 act fp32 + rank1", lambda n: (lambda t: t) if n in tail3 else bf, allbf, True)
        run("tail3 act+w fp32 + rank1", lambda n: (lambda t: t) if n in tail3 else bf, lambda n: (lambda t: t) if n in tail3 else bf, True)
        run("tail3 act fp16 + rank1", lambda n: f16 if n in tail3 else bf, allbf, True)
        run("tail3 act split2 + rank1", lambda n: split2 if n in tail3 else bf, allbf, True)
        run("tail2 act fp32 + rank1", lambda n: (lambda t: t) if n in tail2 else bf, allbf, True)

    print("---- fp16 variants")
    for seed in (0, 1, 2, 3, 4, 5):
        W = mt.init_weights(1, nc, 2, seed=seed, randomize_bn=True)
        x = torch.randn(1, 1, 512, 512, generator=torch.Generator().manual_seed(2))
        ora = mt.OracleNet(W, img, nc, anchors); ora.trace = {}
        want = ora.feature_maps(x)
        if seed == 0:
            for n in names:
                if n in ora.trace and (n in tail2):
                    print("  max|act|", n, float(ora.trace[n].abs().max()))
        def run(tag, ap, wp, rank1=False):
            got = Emu(W, img, nc, anchors, ap, wp, rank1).feature_maps(x)
            print("seed", seed, "%-44s" % tag, ["%.4f" % mt.heads_rel_err(a.numpy(), b.numpy()) for a, b in zip(got, want)], flush=True)
        allbf = lambda n: bf
        t3 = tail3 - {"conv2d_65"}
        run("all bf16", allbf, allbf)
        run("tail3 act fp16 w bf16", lambda n: f16 if n in tail3 else bf, allbf)
        run("tail3 act+w fp16", lambda n: f16 if n in tail3 else bf, lambda n: f16 if n in t3 else bf)
        run("tail2+3 act+w fp16", lambda n: f16 if n in tail2 else bf, lambda n: f16 if n in tail2 else bf)
        run("all fp16", lambda n: f16, lambda n: f16)
