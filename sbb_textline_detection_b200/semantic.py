"""Document-like synthetic models (test / bench support, like synth.py): the REAL ResNet50-U-Net architecture
with hand-placed weights on ONE channel per layer so that the three models of the pipeline (main.py:58-60)
produce plausible label maps on ``synth.document_page`` -- a page box, text-region blobs, text-line bars --
instead of the noise a randomly initialised network emits.  There are no trained ``.h5`` files in this
environment; these stand in for them wherever the reference's host glue (contours, deskew, line separation,
PAGE-XML: main.py:456-2053) has to see a document, e.g. the run()/PAGE-XML gate of BASELINE.json's north_star.

How: channel 0 of every layer is a "semantic" channel that carries a smoothed ink-density (or, for the border
model, brightness) map through the network -- 1x1 convs pass it on, 3x3 convs average it (horizontally for
text lines, isotropically for regions / the page), shortcuts halve and add it, BatchNorm is the identity on it,
every value stays >= 0 so ReLU is inert -- and the classifier thresholds it.  All OTHER channels keep the seeded
He-normal weights and calibrated BatchNorm statistics of ``weights.random_init`` (they read the semantic
channel, never write it), so every MMA still sees dense, O(1) operands, and a small seeded mix of the last
block's other channels enters the class-1 logit, which makes the label boundaries irregular like a real model's.

The BatchNorm statistics of the non-semantic channels and the classifier threshold come from one pass of the
CPU oracle over a sample page (oracle/calibrate_semantic.py -> data/sem_stats_<kind>.npz); everything else is a
function of the seed.
"""
from __future__ import annotations

import os

import numpy as np

from . import weights as W
from .arch import BN_EPS, conv_specs

KINDS = {"page": (2236, 2), "region": (2235, 4), "textline": (2234, 2)}   # kind -> (seed, n_classes)
_ISO = np.full((3, 3), 1.0 / 9.0, np.float32)
_HOR = np.zeros((3, 3), np.float32)
_HOR[1, :] = 1.0 / 3.0
_STAGE_BLOCKS = {2: "abc", 3: "abcd", 4: "abcdef", 5: "abc"}


def _bn_identity(w, bn, ch=0, shift=0.0):
    w[bn + "/gamma"][ch] = 1.0
    w[bn + "/beta"][ch] = shift
    w[bn + "/mean"][ch] = 0.0
    w[bn + "/var"][ch] = 1.0 - BN_EPS


def _stage(w, stage, smooth):
    """Semantic channel through one ResNet stage: conv block 0.5*x_sub + 0.5*smooth, identity blocks x + 0.5*smooth."""
    for b in _STAGE_BLOCKS[stage]:
        base = f"res{stage}{b}_branch"
        w[base + "2a/kernel"][0, 0, 0, 0] = 1.0
        w[base + "2b/kernel"][:, :, 0, 0] = smooth
        w[base + "2c/kernel"][0, 0, 0, 0] = 0.5
        if b == "a":
            w[base + "1/kernel"][0, 0, 0, 0] = 0.5


def semantic_init(kind: str) -> dict:
    """Weight dict (layout of weights.random_init) with the semantic channel installed; BatchNorm moving
    statistics of the other channels still at the Keras defaults (see ``semantic_weights``)."""
    seed, nc = KINDS[kind]
    w = W.random_init(seed, nc)
    for s in conv_specs(nc):
        if s.name != "cls":
            w[s.name + "/kernel"][:, :, :, 0] = 0.0     # nothing but the taps placed below writes channel 0
            w[s.name + "/bias"][0] = 0.0
            _bn_identity(w, s.bn)
    k1 = w["conv1/kernel"]                               # [7, 7, 3, 64], input = BGR / 255
    if kind == "page":
        k1[:, :, :, 0] = 1.0 / 147.0                     # mean brightness; the zero padding reads as dark border
        smooth = _ISO
    else:
        if kind == "textline":
            k1[2:5, :, :, 0] = -1.0 / 63.0               # 3 rows x 7 columns: darkness, averaged along the line
            smooth = _HOR
        else:
            k1[:, :, :, 0] = -1.0 / 147.0
            smooth = _ISO
        w["conv1/bias"][0] = 1.0                         # darkness = 1 - brightness
        _bn_identity(w, "bn_conv1", shift=-0.3)          # ReLU(darkness - 0.3): paper noise -> 0, ink stays
    deepest = {"textline": 2, "region": 3, "page": 3}[kind]
    for stage in range(2, deepest + 1):
        _stage(w, stage, smooth)
    # decoder: pick the semantic channel up from the skip of the deepest level used, then up-sample + smooth
    # cin layout of decK = [up-sampled previous output | skip] (concatenate([up, skip]))
    up_c = {"dec1": 512, "dec2": 512, "dec3": 256, "dec4": 128, "dec5": 64}
    entry = {2: "dec3", 3: "dec2", 4: "dec1"}[deepest]
    if deepest == 4:
        w["dec_v4/kernel"][0, 0, 0, 0] = 1.0             # v4[0] = f4[0]
    names = ["dec1", "dec2", "dec3", "dec4", "dec5"]
    for name in names[names.index(entry):]:
        src = up_c[name] if name == entry else 0         # skip channel 0 at the entry block, up channel 0 after it
        w[name + "/kernel"][:, :, src, 0] = smooth
    return w


def semantic_weights(kind: str, leak: float = 0.08):
    """-> (weights dict, n_classes): ``semantic_init`` + the calibrated BatchNorm statistics of the
    non-semantic channels + the classifier fitted by oracle/calibrate_semantic.py.  ``leak``: weight scale of the
    non-semantic channels in the class-1 logit (0.08 = the committed models; smaller = logits dominated by the
    semantic path, i.e. a better conditioned model -- used by the precision-planner test)."""
    seed, nc = KINDS[kind]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", f"sem_stats_{kind}.npz")
    stats = np.load(path)
    w = W.apply_bn_stats(semantic_init(kind), stats)
    for s in conv_specs(nc):
        if s.name != "cls":
            _bn_identity(w, s.bn, shift=float(w[s.bn + "/beta"][0]))   # calibration touched channel 0's statistics
    install_classifier(w, nc, float(stats["cls_scale"]), float(stats["cls_threshold"]), seed, leak=leak)
    return w, nc


def install_classifier(w, nc, scale, threshold, seed, leak=0.08):
    """logit[1] - logit[0] = scale * (D - threshold) + leak * <r, other channels of the last block>;
    classes >= 2 never win."""
    rng = np.random.default_rng([seed, 77])
    k = np.zeros((1, 1, 32, nc), np.float32)
    b = np.zeros(nc, np.float32)
    k[0, 0, 0, 1] = scale
    k[0, 0, 1:, 1] = leak * rng.standard_normal(31).astype(np.float32)
    b[1] = -scale * threshold
    b[2:] = -30.0
    w["cls/kernel"], w["cls/bias"] = k, b
    for c in range(nc):
        _bn_identity(w, "bn_cls", ch=c)
