"""Mint golden fixtures by EXECUTING THE UNMODIFIED REFERENCE CLASS in this container
(tests/golden/ref_import.py stubs tensorflow/keras, which only provide load_model / session objects on
this path; every numpy / cv2 statement of the reference runs as written, from /root/reference).

    python tests/golden/make_golden_from_reference.py

What this pins (and what it cannot): the reference's own ``do_prediction`` (tile grid, clamp, /255,
argmax, 9-case crop, overwrite order, uint8 cast, the no-patch resize path), ``otsu_copy``,
``resize_image``, ``get_image_and_scales`` and the three stage drivers ``extract_page`` /
``extract_text_regions`` / ``textline_contours`` are the REAL code (main.py:112-113, 178-214, 225-503).
The network behind ``model.predict`` is third-party Keras/TF and stays a restatement: where a fixture
needs a network, the oracle network is plugged into the reference through its duck-typed model
interface (main.py:227-229, 287-288).

Writes (tests/golden/)
  ref_stitch_fake.npz     reference do_prediction(patches=True) with a closed-form, position-dependent
                          fake model: full label maps for small pages, SHA-256 of the label map for the
                          BASELINE grids (2800x2000/448, 4600x3400/672) and other ragged shapes
  ref_nopatch_fake.npz    reference do_prediction(patches=False) with the fake model
  ref_prepost.npz         reference otsu_copy / resize_image / get_image_and_scales on seeded images
  ref_pipeline96.npz      reference extract_page / extract_text_regions / textline_contours on a
                          420x330 synthetic page with the three ORACLE networks at tile 96 plugged in
  ref_page96_textline.npz reference do_prediction(patches=True) on a 300x260 page, oracle textline net
"""
import hashlib
import os
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from fake_model import FakeModel, seeded_page  # noqa: E402
from oracle.resnet50_unet import OracleNet  # noqa: E402
from sbb_textline_detection_b200 import synth  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


STITCH_CASES = [  # (H, W, mh, mw, n_classes, seed, keep_full)
    (300, 260, 96, 96, 2, 1, True),
    (97, 96, 96, 96, 4, 2, True),       # one pixel taller than a tile
    (96, 96, 96, 96, 2, 3, True),       # exactly one tile: nxf = nyf = 2, both clamped to origin 0
    (231, 417, 64, 96, 3, 4, True),     # non-square tile: margin comes from the WIDTH only (main.py:233)
    (1000, 777, 224, 224, 4, 5, False),
    (2800, 2000, 448, 448, 2, 6, False),  # BASELINE config 2 grid
    (4600, 3400, 672, 672, 2, 7, False),  # BASELINE config 5 grid (reference margin rule)
]


def main():
    ref = ref_import.load_reference_main()
    tmp = tempfile.mkdtemp()
    det = ref.textline_detector(os.path.join(tmp, "x.png"), tmp, "x", tmp)

    # ---- do_prediction(patches=True), fake model
    out = {}
    for k, (H, W, mh, mw, nc, seed, keep) in enumerate(STITCH_CASES):
        page = seeded_page(H, W, seed)
        lab = det.do_prediction(True, page, FakeModel(mh, mw, nc))
        assert lab.dtype == np.uint8 and lab.shape == (H, W, 3)
        assert (lab[:, :, 0] == lab[:, :, 1]).all() and (lab[:, :, 0] == lab[:, :, 2]).all()
        out[f"case{k}_params"] = np.array([H, W, mh, mw, nc, seed], np.int64)
        out[f"case{k}_sha"] = np.array(sha(lab[:, :, 0]))
        if keep:
            out[f"case{k}_labels"] = lab[:, :, 0]
        print("stitch", (H, W, mh, mw, nc), out[f"case{k}_sha"])
    np.savez_compressed(os.path.join(HERE, "ref_stitch_fake.npz"), **out)

    # ---- do_prediction(patches=False), fake model (needs self.image for the resize back, main.py:378)
    out = {}
    for k, (H, W, mh, mw, nc, seed) in enumerate([(420, 330, 96, 96, 2, 11), (2800, 2000, 448, 448, 2, 12),
                                                   (131, 517, 64, 96, 4, 13)]):
        page = seeded_page(H, W, seed)
        det.image = page
        lab = det.do_prediction(False, page, FakeModel(mh, mw, nc))
        assert lab.dtype == np.uint8 and lab.shape == (H, W, 3)
        out[f"case{k}_params"] = np.array([H, W, mh, mw, nc, seed], np.int64)
        out[f"case{k}_sha"] = np.array(sha(lab))
        if H * W < 300000:
            out[f"case{k}_labels"] = lab[:, :, 0]
    np.savez_compressed(os.path.join(HERE, "ref_nopatch_fake.npz"), **out)

    # ---- otsu_copy / resize_image / get_image_and_scales
    out = {}
    doc = synth.document_page(500, 380, seed=31)
    o = det.otsu_copy(doc)
    out["otsu_in_seed"] = np.array([500, 380, 31])
    out["otsu_dtype"] = np.array(str(o.dtype))
    out["otsu_packed"] = np.packbits((o[:, :, 0] > 0))
    out["otsu_allch_equal"] = np.array(bool((o[:, :, 0] == o[:, :, 1]).all() and (o[:, :, 0] == o[:, :, 2]).all()))
    out["otsu_values"] = np.unique(o)
    rnd = seeded_page(333, 211, 32)
    for k, (oh, ow) in enumerate([(448, 448), (96, 96), (700, 500), (100, 641)]):
        out[f"resize{k}_hw"] = np.array([oh, ow])
        out[f"resize{k}_sha"] = np.array(sha(det.resize_image(rnd, oh, ow)))
    for k, (h, w) in enumerate([(420, 330), (2499, 1800), (2500, 1800), (3000, 2113)]):
        png = os.path.join(tmp, f"s{k}.png")
        cv2.imwrite(png, seeded_page(h, w, 40 + k))
        d = ref.textline_detector(png, tmp, None, tmp)
        d.get_image_and_scales()
        out[f"scale{k}"] = np.array([h, w, d.img_hight_int, d.img_width_int, d.height_org, d.width_org], np.int64)
        out[f"scale{k}_f"] = np.array([d.scale_y, d.scale_x], np.float64)
        out[f"scale{k}_sha"] = np.array(sha(d.image))
        out[f"scale{k}_fname"] = np.array(d.f_name)
    np.savez_compressed(os.path.join(HERE, "ref_prepost.npz"), **out)

    # ---- the three stage drivers with the oracle networks plugged in (tile 96), BASELINE config 3
    T = 96
    nets = {}
    for kind, fname in (("page", "model_page_mixed_best.h5"), ("region", "model_strukturerkennung.h5"),
                        ("textline", "model_textline_new.h5")):
        w, nc = synthetic_weights(kind)
        nets[kind] = OracleNet(w, nc).as_keras_like(T, T)
        ref_import.MODEL_FACTORY[tmp + "/" + fname] = (lambda m=nets[kind]: m)
    page = synth.document_page(420, 330, seed=7)
    det = ref.textline_detector(os.path.join(tmp, "p.png"), tmp, "p", tmp)
    det.image = page.copy()
    border = det.do_prediction(False, det.image, nets["page"])
    image_page, page_coord = det.extract_page()       # deletes det.image (main.py:431)
    regions = det.extract_text_regions(image_page)
    textline = det.textline_contours(image_page)
    print("pipeline: page_coord", page_coord, "regions classes", np.unique(regions), "textline", np.unique(textline))
    np.savez_compressed(os.path.join(HERE, "ref_pipeline96.npz"), page_seed=np.array([420, 330, 7]),
                        border=border[:, :, 0], page_coord=np.array(page_coord), cont_page=det.cont_page[0],
                        regions=regions[:, :, 0], regions_allch_equal=np.array(bool((regions[:, :, 0] == regions[:, :, 1]).all())),
                        textline=textline)

    pg = synth.document_page(300, 260, seed=21)
    lab = det.do_prediction(True, pg, nets["textline"])
    np.savez_compressed(os.path.join(HERE, "ref_page96_textline.npz"), page_seed=np.array([300, 260, 21]),
                        labels=lab[:, :, 0])
    print("done")


if __name__ == "__main__":
    main()
