"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the network behind
``model.predict`` in the reference's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path (``sbb_textline_detection_b200``) never does.

PARITY UNPINNED (this module only -- the tiling / stitch / pre-post / deskew / XML oracles are pinned to the
reference's own execution): the reference ships no tests, golden vectors or weights for this path and its own
arithmetic lives in third-party tensorflow-gpu==1.15.* / keras==2.3.* (requirements.txt:5,10) which
cannot be installed here.  The architecture is not in /root/reference either: it is whatever
``keras.models.load_model`` deserialises (main.py:216-223); the reference only fixes the call sites
``model.layers[-1].output_shape`` (main.py:227-229) and ``model.predict`` (main.py:287-288, 373-374).
(An independent witness exists for the encoder: tests/test_oracle_witness.py checks it against torchvision's
ResNet-50 reconfigured to the Keras-v1 conventions, and the decoder against a float64 numpy restatement with explicit
slices -- that rules out a misreading of ResNet-50 and of torch's conv conventions, it is not a pin to the reference.)  What follows restates the published ``resnet50_unet`` of qurator-spk/sbb_pixelwise_segmentation
(README.md:16 names that repo as the training code) with Keras 2.3 inference numerics:

  * channels_last, every Conv2D has a bias, kernels stored HWIO
  * BatchNormalization inference: y = gamma*(x-mean)/sqrt(var+1e-3)+beta  (Keras eps, not torch's)
  * MaxPooling2D((3,3), strides=2) is 'valid'  (224 -> 111)
  * stride sits on the FIRST 1x1 of a conv_block and on its shortcut (Keras-v1 ResNet50)
  * one_side_pad(x) = ZeroPadding2D(1)(x)[:, :-1, :-1, :]  (111 -> 112, data shifted by (+1,+1))
  * decoder: UpSampling2D(2) nearest, concatenate([up, skip]), ZeroPadding2D(1), Conv2D 3x3 valid
  * head: Conv2D 1x1 -> BatchNormalization -> softmax over channels

Weights live in a flat dict of numpy arrays keyed ``<layer>/kernel|bias`` and
``<bn>/gamma|beta|mean|var`` (kernel layout HWIO, like the .h5 files the reference loads).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # keras.layers.BatchNormalization default

# (stage, blocks, (f1, f2, f3), stride of block 'a')
STAGES = (
    (2, "abc", (64, 64, 256), 1),
    (3, "abcd", (128, 128, 512), 2),
    (4, "abcdef", (256, 256, 1024), 2),
    (5, "abc", (512, 512, 2048), 2),
)


def conv_specs(n_classes: int = 2):
    """Ordered list of (conv_name, bn_name_or_None, kh, kw, cin, cout) for every Conv2D."""
    specs = [("conv1", "bn_conv1", 7, 7, 3, 64)]
    cin = 64
    for stage, blocks, (f1, f2, f3), _ in STAGES:
        for b in blocks:
            base = f"res{stage}{b}_branch"
            bnb = f"bn{stage}{b}_branch"
            specs.append((base + "2a", bnb + "2a", 1, 1, cin, f1))
            specs.append((base + "2b", bnb + "2b", 3, 3, f1, f2))
            specs.append((base + "2c", bnb + "2c", 1, 1, f2, f3))
            if b == "a":
                specs.append((base + "1", bnb + "1", 1, 1, cin, f3))
            cin = f3
    specs += [
        ("dec_v5", "bn_dec_v5", 1, 1, 2048, 512),
        ("dec_v4", "bn_dec_v4", 1, 1, 1024, 512),
        ("dec1", "bn_dec1", 3, 3, 1024, 512),
        ("dec2", "bn_dec2", 3, 3, 1024, 256),
        ("dec3", "bn_dec3", 3, 3, 512, 128),
        ("dec4", "bn_dec4", 3, 3, 192, 64),
        ("dec5", "bn_dec5", 3, 3, 67, 32),
        ("cls", "bn_cls", 1, 1, 32, n_classes),
    ]
    return specs


def conv_flops(tile_h: int, tile_w: int, n_classes: int = 2):
    """2*MACs of every Conv2D for one tile: returns (total, encoder, decoder) in FLOP."""
    def o2(v):  # 7x7 s2 pad 3
        return (v + 6 - 7) // 2 + 1
    h1, w1 = o2(tile_h), o2(tile_w)
    h2, w2 = (h1 - 3) // 2 + 1, (w1 - 3) // 2 + 1
    hw = {1: (h1, w1), 2: (h2, w2)}
    h, w = h2, w2
    for s in (3, 4, 5):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        hw[s] = (h, w)
    enc = 2 * h1 * w1 * 147 * 64
    cin = 64
    for stage, blocks, (f1, f2, f3), _ in STAGES:
        m = hw[stage][0] * hw[stage][1]
        for b in blocks:
            enc += 2 * m * (cin * f1 + 9 * f1 * f2 + f2 * f3)
            if b == "a":
                enc += 2 * m * cin * f3
            cin = f3
    m5 = hw[5][0] * hw[5][1]
    m4 = hw[4][0] * hw[4][1]
    dec = 2 * m5 * 2048 * 512 + 2 * m4 * 1024 * 512
    dec += 2 * m4 * 9 * 1024 * 512
    dec += 2 * (4 * m4) * 9 * 1024 * 256
    dec += 2 * (16 * m4) * 9 * 512 * 128
    dec += 2 * (64 * m4) * 9 * 192 * 64
    dec += 2 * (256 * m4) * 9 * 67 * 32
    dec += 2 * (256 * m4) * 32 * n_classes
    return enc + dec, enc, dec


class OracleNet:
    """fp32 (or fp64) CPU forward of the ResNet50-U-Net.  ``quant`` optionally rounds every conv
    operand (activations and BN-folded weights) to fp16/bf16 to SIMULATE the GPU numerics; it is a
    diagnostic, the oracle proper is quant=None."""

    def __init__(self, weights: dict, n_classes: int, dtype=torch.float32, quant=None):
        self.n_classes = n_classes
        self.dtype = dtype
        self.quant = quant
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}
        self._k = {}
        for name, bn, kh, kw, cin, cout in conv_specs(n_classes):
            k = self.w[name + "/kernel"]
            assert tuple(k.shape) == (kh, kw, cin, cout), (name, tuple(k.shape))
            self._k[name] = k.permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
        self.calibrating = False
        self.taps = None  # optional dict name -> tensor (NCHW) captured during forward

    # -- primitives ---------------------------------------------------------------------------
    def _q(self, t):
        if self.quant is None:
            return t
        return t.to(self.quant).to(self.dtype)

    def _conv(self, x, name, stride=1, pad=0, bn=None, relu=False, add=None):
        """conv (+bias) [+BN] [+residual] [+ReLU].  With quant set, BN is folded into the weights
        first (as the GPU path does) and operands are rounded; otherwise Keras order is kept."""
        k, b = self._k[name], self.w[name + "/bias"]
        if self.calibrating and bn is not None:
            y = F.conv2d(x, k, b, stride=stride, padding=pad)
            mean = y.mean(dim=(0, 2, 3))
            var = y.var(dim=(0, 2, 3), unbiased=False)
            self.w[bn + "/mean"] = mean.clone()
            self.w[bn + "/var"] = var.clone()
        if self.quant is not None and bn is not None:
            s = self.w[bn + "/gamma"] / torch.sqrt(self.w[bn + "/var"] + BN_EPS)
            kf = self._q(k * s[:, None, None, None])
            bf = (b - self.w[bn + "/mean"]) * s + self.w[bn + "/beta"]
            y = F.conv2d(self._q(x), kf, bf, stride=stride, padding=pad)
        else:
            y = F.conv2d(self._q(x), self._q(k), b, stride=stride, padding=pad)
            if bn is not None:
                y = self._bn(y, bn)
        if add is not None:
            y = y + add
        if relu:
            y = F.relu(y)
        return y

    def _bn(self, x, bn):
        g, b = self.w[bn + "/gamma"], self.w[bn + "/beta"]
        m, v = self.w[bn + "/mean"], self.w[bn + "/var"]
        s = g / torch.sqrt(v + BN_EPS)
        return x * s[None, :, None, None] + (b - m * s)[None, :, None, None]

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t

    # -- network ------------------------------------------------------------------------------
    def logits(self, x_nhwc: np.ndarray) -> torch.Tensor:
        """x: [N,H,W,3] float in [0,1] (BGR/255).  Returns pre-softmax BN output [N,H,W,C]."""
        x = torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(self.dtype).permute(0, 3, 1, 2)
        inp = x
        # stem: ZeroPad(3) + 7x7 s2; f1 is the RAW conv output (before BN/ReLU)
        f1 = self._conv(x, "conv1", stride=2, pad=3)
        if self.calibrating:
            self.w["bn_conv1/mean"] = f1.mean(dim=(0, 2, 3)).clone()
            self.w["bn_conv1/var"] = f1.var(dim=(0, 2, 3), unbiased=False).clone()
        f1 = self._q(f1)  # the GPU path stores f1 in half precision
        self._tap("conv1", f1)
        x = F.relu(self._bn(f1, "bn_conv1"))
        x = F.max_pool2d(x, 3, 2)
        self._tap("pool1", x)
        feats = {}
        for stage, blocks, _, stride in STAGES:
            for b in blocks:
                base, bnb = f"res{stage}{b}_branch", f"bn{stage}{b}_branch"
                s = stride if b == "a" else 1
                h = self._conv(x, base + "2a", stride=s, bn=bnb + "2a", relu=True)
                h = self._conv(h, base + "2b", pad=1, bn=bnb + "2b", relu=True)
                if b == "a":
                    sc = self._conv(x, base + "1", stride=s, bn=bnb + "1")
                else:
                    sc = x
                x = self._conv(h, base + "2c", bn=bnb + "2c", add=sc, relu=True)
                self._tap(f"res{stage}{b}", x)
            feats[stage] = x
        f2 = F.pad(feats[2], (1, 0, 1, 0))  # one_side_pad: zero row on top, zero col on the left
        f3, f4, f5 = feats[3], feats[4], feats[5]
        v5 = self._conv(f5, "dec_v5", bn="bn_dec_v5", relu=True)
        v4 = self._conv(f4, "dec_v4", bn="bn_dec_v4", relu=True)
        self._tap("dec_v5", v5)
        self._tap("dec_v4", v4)
        o = v5
        for i, skip in enumerate((v4, f3, f2, f1, inp), start=1):
            o = F.interpolate(o, scale_factor=2, mode="nearest")
            o = torch.cat([o, skip], dim=1)
            o = self._conv(o, f"dec{i}", pad=1, bn=f"bn_dec{i}", relu=True)
            self._tap(f"dec{i}", o)
        o = self._conv(o, "cls", bn="bn_cls")
        return o.permute(0, 2, 3, 1).contiguous()

    def predict(self, x_nhwc: np.ndarray) -> np.ndarray:
        """Equivalent of keras ``model.predict``: softmax probabilities [N,H,W,C] float32."""
        with torch.no_grad():
            z = self.logits(x_nhwc)
            p = torch.softmax(z, dim=-1)
        return p.to(torch.float32).numpy()

    # -- the two attributes do_prediction touches (main.py:227-229) --------------------------
    def as_keras_like(self, tile_h: int, tile_w: int):
        return KerasLikeModel(self, tile_h, tile_w)


class _Layer:
    def __init__(self, shape):
        self.output_shape = shape


class KerasLikeModel:
    """Duck-types what do_prediction uses of a Keras model (main.py:227-229, 287-288)."""

    def __init__(self, net: OracleNet, tile_h: int, tile_w: int):
        self.net = net
        self.layers = [_Layer((None, tile_h, tile_w, net.n_classes))]

    def predict(self, x):
        return self.net.predict(np.asarray(x, dtype=np.float32))
