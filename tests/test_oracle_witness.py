"""Independent witness for the oracle's ENCODER arithmetic.  The network oracle (oracle/resnet50_unet.py) is a
restatement that nothing of the reference can pin (no TF 1.15 / Keras 2.3, no .h5 here).
What can be done is to check it against an implementation written by someone else: torchvision's ResNet-50
graph, reconfigured to the Keras-v1 conventions the reference's models were trained with --

  * stride 2 on the FIRST 1x1 conv of a down-sampling block and on its shortcut (torchvision: on the 3x3)
  * 3x3/2 max-pool without padding (224 -> 111)
  * BatchNorm eps 1e-3; Keras convs carry a bias, torchvision's do not: BN(x + b) == BN with mean - b

-- loaded with the SAME seeded weights.  All four stage outputs (the skip tensors of the U-Net) must agree.
This covers 53 of the 61 convolutions; the decoder is specific to sbb_pixelwise_segmentation and has no
independent implementation to compare with."""
import numpy as np
import pytest
import torch

from oracle.resnet50_unet import BN_EPS, STAGES, OracleNet
from sbb_textline_detection_b200 import synth

tv = pytest.importorskip("torchvision")


def _load(conv, bn, w, name, bn_name):
    k = torch.from_numpy(w[name + "/kernel"]).permute(3, 2, 0, 1).contiguous()     # HWIO -> OIHW
    assert conv.weight.shape == k.shape, (name, conv.weight.shape, k.shape)
    conv.weight.data.copy_(k)
    bn.eps = BN_EPS
    bn.weight.data.copy_(torch.from_numpy(w[bn_name + "/gamma"]))
    bn.bias.data.copy_(torch.from_numpy(w[bn_name + "/beta"]))
    bn.running_mean.data.copy_(torch.from_numpy(w[bn_name + "/mean"] - w[name + "/bias"]))   # folds the conv bias
    bn.running_var.data.copy_(torch.from_numpy(w[bn_name + "/var"]))


def torchvision_encoder(w):
    net = tv.models.resnet50(weights=None)
    net.eval()
    net.maxpool = torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=0)
    _load(net.conv1, net.bn1, w, "conv1", "bn_conv1")
    for (stage, blocks, _, stride), layer in zip(STAGES, (net.layer1, net.layer2, net.layer3, net.layer4)):
        for b, blk in zip(blocks, layer):
            base, bnb = f"res{stage}{b}_branch", f"bn{stage}{b}_branch"
            _load(blk.conv1, blk.bn1, w, base + "2a", bnb + "2a")
            _load(blk.conv2, blk.bn2, w, base + "2b", bnb + "2b")
            _load(blk.conv3, blk.bn3, w, base + "2c", bnb + "2c")
            if b == "a":
                s = (stride, stride)
                blk.conv1.stride, blk.conv2.stride = s, (1, 1)                      # Keras v1: stride on the first 1x1
                _load(blk.downsample[0], blk.downsample[1], w, base + "1", bnb + "1")
                blk.downsample[0].stride = s
            else:
                assert blk.downsample is None
    return net


def test_oracle_encoder_equals_reconfigured_torchvision_resnet50(textline_weights):
    w, nc = textline_weights
    x = np.stack([synth.document_page(160, 160, seed=3), synth.uniform_page(160, 160, 1)]).astype(np.float32) / np.float32(255)
    oracle = OracleNet(w, nc)
    oracle.taps = {}
    with torch.no_grad():
        oracle.logits(x)
    net = torchvision_encoder(w)
    feats = {}
    with torch.no_grad():
        t = torch.from_numpy(x).permute(0, 3, 1, 2)
        t = net.maxpool(net.relu(net.bn1(net.conv1(t))))
        assert t.shape[2:] == oracle.taps["pool1"].shape[2:] == (39, 39)           # (160/2 - 3)//2 + 1: 'valid' pooling
        for name, layer in (("res2c", net.layer1), ("res3d", net.layer2), ("res4f", net.layer3), ("res5c", net.layer4)):
            t = layer(t)
            feats[name] = t
    for name, t in feats.items():
        ref = oracle.taps[name]
        assert t.shape == ref.shape, name
        err = (t - ref).abs().max().item()
        assert err <= 2e-4 * max(1.0, ref.abs().max().item()), (name, err)


def _np_conv_bn(x, w, name, bn, relu, pad):
    """Keras Conv2D (cross-correlation, HWIO kernel, bias) + BatchNormalization(eps 1e-3) (+ ReLU) on NHWC float64,
    written with explicit shifted slices: no torch conv / pad / layout permutation is involved."""
    k = w[name + "/kernel"].astype(np.float64)
    kh, kw = k.shape[:2]
    if pad:
        x = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)))
    H, W = x.shape[1] - kh + 1, x.shape[2] - kw + 1
    y = np.zeros(x.shape[:1] + (H, W, k.shape[3]))
    for dy in range(kh):
        for dx in range(kw):
            y += x[:, dy:dy + H, dx:dx + W, :] @ k[dy, dx]
    y += w[name + "/bias"].astype(np.float64)
    y = (y - w[bn + "/mean"]) / np.sqrt(w[bn + "/var"].astype(np.float64) + BN_EPS) * w[bn + "/gamma"] + w[bn + "/beta"]
    return np.maximum(y, 0.0) if relu else y


def test_oracle_decoder_equals_numpy_restatement(textline_weights):
    """The decoder has no third-party implementation to compare with (it is specific to sbb_pixelwise_segmentation), so
    its witness is a second, differently written statement of SURVEY App. A: float64 numpy, NHWC, explicit slices --
    UpSampling2D(2) as np.repeat, concatenate([up, skip]), ZeroPadding2D(1) + 3x3 'valid' conv, BN eps 1e-3, ReLU; the
    stride-2 skip of the 111-grid padded by one zero row on top and one zero column on the left; the classifier 1x1 +
    BN.  It starts from the oracle's OWN skip tensors (the encoder is covered by the torchvision witness) and must
    reproduce every decoder activation and the logits."""
    w, nc = textline_weights
    x = np.stack([synth.document_page(64, 96, seed=5), synth.uniform_page(64, 96, 7)]).astype(np.float32) / np.float32(255)
    oracle = OracleNet(w, nc)
    oracle.taps = {}
    with torch.no_grad():
        z = oracle.logits(x).numpy()
    nhwc = lambda name: oracle.taps[name].permute(0, 2, 3, 1).numpy().astype(np.float64)   # noqa: E731
    f1, f2, f3, f4, f5 = nhwc("conv1"), nhwc("res2c"), nhwc("res3d"), nhwc("res4f"), nhwc("res5c")
    f2 = np.pad(f2, ((0, 0), (1, 0), (1, 0), (0, 0)))                                     # one_side_pad
    v5 = _np_conv_bn(f5, w, "dec_v5", "bn_dec_v5", True, 0)
    v4 = _np_conv_bn(f4, w, "dec_v4", "bn_dec_v4", True, 0)
    o = v5
    for i, skip in enumerate((v4, f3, f2, f1, x.astype(np.float64)), start=1):
        o = np.repeat(np.repeat(o, 2, axis=1), 2, axis=2)
        assert o.shape[1:3] == skip.shape[1:3], (i, o.shape, skip.shape)
        o = _np_conv_bn(np.concatenate([o, skip], axis=3), w, f"dec{i}", f"bn_dec{i}", True, 1)
        ref = nhwc(f"dec{i}")
        assert np.abs(o - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), f"dec{i}"
    got = _np_conv_bn(o, w, "cls", "bn_cls", False, 0)
    assert got.shape == z.shape
    assert np.abs(got - z).max() <= 5e-4 * max(1.0, np.abs(z).max())
