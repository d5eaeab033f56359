"""ORACLE tooling: mint the calibrated BatchNorm statistics for the seeded random-init models
(SURVEY.md 8(d) "calibrated random init").  One oracle forward in calibrating mode over tiles of
a document-like synthetic page sets every BN's moving mean/var to the observed per-channel stats,
so activations stay O(1) through the 61 convs and class margins are not degenerate.

    python -m oracle.calibrate            # writes sbb_textline_detection_b200/data/bn_stats_*.npz

The .npz files are committed: they make the synthetic weights bit-reproducible on any machine
(numpy Generator streams + stored stats), which is what lets tests/golden/ fixtures travel."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.resnet50_unet import OracleNet  # noqa: E402
from sbb_textline_detection_b200 import synth, weights  # noqa: E402

MODELS = {  # name -> (seed, n_classes)   (file names of main.py:58-60)
    "textline": (1234, 2),
    "region": (1235, 4),
    "page": (1236, 2),
}


def calibrate(seed: int, n_classes: int, tile: int = 448):
    w = weights.random_init(seed, n_classes)
    net = OracleNet(w, n_classes, dtype=torch.float64)
    page = synth.document_page(1344, 1344, seed=7)
    tiles = np.stack([page[y:y + tile, x:x + tile] for y in (0, 448, 896) for x in (0, 448)][:4])
    net.calibrating = True
    with torch.no_grad():
        net.logits(tiles.astype(np.float64) / 255.0)
    net.calibrating = False
    return {k: v.numpy().astype(np.float32) for k, v in net.w.items()
            if k.endswith("/mean") or k.endswith("/var")}


if __name__ == "__main__":
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           "sbb_textline_detection_b200", "data")
    os.makedirs(out_dir, exist_ok=True)
    for name, (seed, nc) in MODELS.items():
        stats = calibrate(seed, nc)
        path = os.path.join(out_dir, f"bn_stats_{name}.npz")
        np.savez_compressed(path, **stats)
        print(name, seed, nc, "->", path, os.path.getsize(path))
