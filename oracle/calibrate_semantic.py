"""ORACLE tooling: calibrate the document-like synthetic models of sbb_textline_detection_b200/semantic.py.

For each of the three models one oracle pass in calibrating mode over sample inputs of the kind that model
sees in the pipeline (main.py:384-503: the border model the whole page resized to the tile; the region model
Otsu-binarised tiles; the textline model raw tiles) sets the BatchNorm statistics of the non-semantic channels;
a second pass reads the semantic channel of the last decoder block and fits the classifier threshold to a
target mask derived from the page's ink (the kind of blobs the reference's glue expects from each model).

    python -m oracle.calibrate_semantic      # writes sbb_textline_detection_b200/data/sem_stats_*.npz
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import do_prediction as odp  # noqa: E402
from oracle.resnet50_unet import OracleNet  # noqa: E402
from sbb_textline_detection_b200 import semantic, synth  # noqa: E402

TILE = 448


def sample_inputs(kind):
    """-> (tiles uint8 [n, TILE, TILE, 3], target masks bool [n, TILE, TILE])"""
    page = synth.document_page(1400, 1000, seed=5)
    page = odp.resize_nearest(page, 2800, 2000)                      # get_image_and_scales, main.py:196-214
    ink = (page[:, :, 0] < 128).astype(np.uint8)
    if kind == "page":
        tiles, masks = [], []
        for seed in (11, 12, 13):
            p = synth.framed_page(2800, 2000, seed)
            tiles.append(odp.resize_nearest(p, TILE, TILE))
            m = np.zeros((2800, 2000), np.uint8)
            m[160:-160, 160:-160] = 1      # target a little INSIDE the paper: the crop must not keep scanner border
            masks.append(odp.resize_nearest(m, TILE, TILE) > 0)
        return np.stack(tiles), np.stack(masks)
    if kind == "region":
        src = odp.otsu_copy(page).astype(np.uint8)
        target = cv2.dilate(ink, np.ones((45, 45), np.uint8)) > 0
    else:
        src = page
        target = cv2.dilate(ink, np.ones((5, 25), np.uint8)) > 0
    org = [(0, 0), (700, 300), (1400, 800), (2352, 1552), (1000, 1200), (300, 1500)]
    tiles = np.stack([src[y:y + TILE, x:x + TILE] for y, x in org])
    masks = np.stack([target[y:y + TILE, x:x + TILE] for y, x in org])
    return tiles, masks


def calibrate(kind):
    seed, nc = semantic.KINDS[kind]
    w = semantic.semantic_init(kind)
    semantic.install_classifier(w, nc, 1.0, 0.0, seed)
    tiles, masks = sample_inputs(kind)
    x = tiles.astype(np.float64) / 255.0
    net = OracleNet(w, nc, dtype=torch.float64)
    net.calibrating = True
    with torch.no_grad():
        net.logits(x)
    net.calibrating = False
    stats = {k: v.numpy().astype(np.float32) for k, v in net.w.items() if k.endswith("/mean") or k.endswith("/var")}
    # second pass with the statistics installed the way semantic_weights installs them
    from sbb_textline_detection_b200 import weights as W
    from sbb_textline_detection_b200.arch import conv_specs
    w2 = W.apply_bn_stats(semantic.semantic_init(kind), stats)
    for s in conv_specs(nc):
        if s.name != "cls":
            semantic._bn_identity(w2, s.bn, shift=float(w2[s.bn + "/beta"][0]))
    semantic.install_classifier(w2, nc, 1.0, 0.0, seed)
    net = OracleNet(w2, nc, dtype=torch.float32)
    net.taps = {}
    with torch.no_grad():
        net.logits(tiles.astype(np.float32) / np.float32(255))
    D = net.taps["dec5"][:, 0].numpy()                                 # semantic channel of the last block
    inner = np.zeros_like(masks)
    if kind == "page":
        inner[:] = True                                                # patches=False: the whole map is used
    else:
        inner[:, 44:-44, 44:-44] = True                                # the part of a tile the stitch keeps
    best = (-1.0, 0.0)
    for thr in np.quantile(D[inner], np.linspace(0.02, 0.98, 97)):
        pred = D > thr
        iou = (pred & masks & inner).sum() / max(((pred | masks) & inner).sum(), 1)
        if iou > best[0]:
            best = (iou, float(thr))
    spread = float(D[inner & masks].mean() - D[inner & ~masks].mean())
    scale = 8.0 / max(spread, 1e-6)                                    # +-4 logit units between the class means
    print(f"{kind}: threshold {best[1]:.4f} (IoU vs target {best[0]:.3f}), D text {D[inner & masks].mean():.3f} / "
          f"bg {D[inner & ~masks].mean():.3f}, scale {scale:.3f}")
    stats["cls_threshold"] = np.float32(best[1])
    stats["cls_scale"] = np.float32(scale)
    return stats


if __name__ == "__main__":
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           "sbb_textline_detection_b200", "data")
    for kind in (sys.argv[1:] or list(semantic.KINDS)):
        stats = calibrate(kind)
        path = os.path.join(out_dir, f"sem_stats_{kind}.npz")
        np.savez_compressed(path, **stats)
        print(kind, "->", path, os.path.getsize(path))
