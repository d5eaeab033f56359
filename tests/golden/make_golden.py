"""Mint the golden fixtures in this directory FROM THE ORACLE (the reference itself cannot run here:
no tensorflow/keras/h5py, no .h5 weights -- PARITY UNPINNED, see oracle/resnet50_unet.py).

    python tests/golden/make_golden.py

Writes
  tile96_textline.npz   one 96x96 tile: input (uint8), oracle logits fp32, labels
  tile448_textline.npz  one 448x448 tile: input, labels (bit-packed), logits at 512 sampled pixels
  page96_region.npz     do_prediction(patches=True) of a 300x260 page with a 96x96 4-class model:
                        input page, stitched label map
  stitch_hash.npz       do_prediction stitch replay with a closed-form fake model (exact integers):
                        pins tile grid + 9-case crop + overwrite order without any float arithmetic
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.do_prediction import do_prediction, stitch_replay, tile_grid  # noqa: E402
from oracle.resnet50_unet import OracleNet  # noqa: E402
from sbb_textline_detection_b200 import synth  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402


def fake_seg(t, i, j, x0, y0, mh, mw):
    """Closed-form per-tile class map: depends on the tile index AND the in-tile position, so any
    mistake in crop offsets or overwrite order changes the stitched result."""
    yy, xx = np.mgrid[0:mh, 0:mw]
    return ((yy * 7 + xx * 13 + t * 31 + i * 3 + j * 5) % 251).astype(np.int64)


def main():
    torch.manual_seed(0)
    w, nc = synthetic_weights("textline")
    net = OracleNet(w, nc, torch.float32)
    page = synth.document_page(700, 600, seed=11)
    t96 = page[100:196, 200:296]
    with torch.no_grad():
        z = net.logits(t96[None].astype(np.float32) / np.float32(255.0)).numpy()[0]
    np.savez_compressed(os.path.join(HERE, "tile96_textline.npz"), tile=t96, logits=z,
                        labels=z.argmax(-1).astype(np.uint8))

    big = synth.document_page(2800, 2000, seed=0)[360:808, 360:808]
    with torch.no_grad():
        z = net.logits(big[None].astype(np.float32) / np.float32(255.0)).numpy()[0]
    rng = np.random.default_rng(5)
    ys, xs = rng.integers(0, 448, 512), rng.integers(0, 448, 512)
    np.savez_compressed(os.path.join(HERE, "tile448_textline.npz"), tile=big,
                        labels_packed=np.packbits(z.argmax(-1).astype(np.uint8)), ys=ys, xs=xs,
                        logits_sampled=z[ys, xs])

    wr, ncr = synthetic_weights("region")
    netr = OracleNet(wr, ncr, torch.float32).as_keras_like(96, 96)
    pg = synth.document_page(300, 260, seed=12)
    lab = do_prediction(True, pg, netr, predict_batch=8)[:, :, 0]
    np.savez_compressed(os.path.join(HERE, "page96_region.npz"), page=pg, labels=lab)

    cases = []
    for (H, W, mh, mw, margin) in [(2800, 2000, 448, 448, None), (4600, 3400, 672, 672, None),
                                   (4600, 3400, 672, 672, 168), (448, 448, 448, 448, None),
                                   (1000, 449, 448, 448, None), (901, 1203, 448, 448, 0)]:
        m, nxf, nyf, tiles = tile_grid(H, W, mh, mw, margin)
        out = stitch_replay(H, W, mh, mw, m, nxf, nyf, tiles,
                            lambda t, i, j, x0, y0: fake_seg(t, i, j, x0, y0, mh, mw))[:, :, 0]
        # store a strided sample + checksums (full maps would be tens of MB)
        cases.append(dict(H=H, W=W, mh=mh, mw=mw, margin=-1 if margin is None else margin, nxf=nxf, nyf=nyf,
                          rowsum=out.sum(axis=1, dtype=np.int64), colsum=out.sum(axis=0, dtype=np.int64),
                          sample=out[::37, ::41].copy()))
    np.savez_compressed(os.path.join(HERE, "stitch_hash.npz"),
                        **{f"c{k}_{name}": np.asarray(v) for k, c in enumerate(cases) for name, v in c.items()})
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
