"""Parity against fixtures minted by EXECUTING THE UNMODIFIED REFERENCE CLASS
(tests/golden/make_golden_from_reference.py: /root/reference/.../main.py run in place with only
tensorflow/keras stubbed out).  These pin the oracle restatement, the product's host logic and --
on the GPU -- the product's page path to the reference's own code for main.py:112-113, 178-214,
225-503; the one thing the reference cannot pin here is the Keras network arithmetic itself."""
import hashlib
import os
import sys

import cv2
import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import do_prediction as odp
from oracle.resnet50_unet import OracleNet
from sbb_textline_detection_b200 import synth
from sbb_textline_detection_b200.model import compute_tile_grid

sys.path.insert(0, GOLDEN)
from fake_model import FakeModel, seeded_page  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _cases(name):
    g = golden(name)
    n = len([k for k in g.files if k.endswith("_params")])
    return g, n


# ----------------------------------------------------------------------------- CPU: oracle + host logic
def test_oracle_do_prediction_equals_reference_stitch():
    """oracle.do_prediction (restatement) == reference do_prediction, bit for bit, incl. the BASELINE
    config 2 and config 5 grids."""
    g, n = _cases("ref_stitch_fake.npz")
    assert n == 7
    for k in range(n):
        H, W, mh, mw, nc, seed = (int(v) for v in g[f"case{k}_params"])
        lab = odp.do_prediction(True, seeded_page(H, W, seed), FakeModel(mh, mw, nc), predict_batch=16)
        assert lab.dtype == np.uint8 and lab.shape == (H, W, 3)
        assert sha(lab[:, :, 0]) == str(g[f"case{k}_sha"]), (H, W, mh, mw)
        if f"case{k}_labels" in g.files:
            assert (lab[:, :, 0] == g[f"case{k}_labels"]).all()
            assert (lab[:, :, 1] == lab[:, :, 0]).all() and (lab[:, :, 2] == lab[:, :, 0]).all()


def test_product_tile_grid_and_owner_tables_equal_reference_stitch(built_lib):
    """The product's host-side integer logic (sbb_compute_tile_grid: origins + separable owner tables,
    the test the fused GPU epilogue applies) reproduces the reference's stitched label map exactly."""
    g, n = _cases("ref_stitch_fake.npz")
    for k in range(n):
        H, W, mh, mw, nc, seed = (int(v) for v in g[f"case{k}_params"])
        page = seeded_page(H, W, seed)
        fm = FakeModel(mh, mw, nc)
        nx, ny, org, ox, oy = compute_tile_grid(H, W, mh, mw, -1)
        out = np.zeros((H, W), np.uint8)
        for (x0, y0, i, j) in org.tolist():
            seg = fm.classes(page[None, y0:y0 + mh, x0:x0 + mw].astype(np.float64) / 255.0)[0]
            own = (oy[y0:y0 + mh, None] == j) & (ox[None, x0:x0 + mw] == i)
            out[y0:y0 + mh, x0:x0 + mw][own] = seg[own]
        assert sha(out) == str(g[f"case{k}_sha"]), (H, W, mh, mw)


def test_detector_generic_path_equals_reference(tmp_path):
    """The drop-in class with a duck-typed (non-GPU) model == the reference class, both branches."""
    from sbb_textline_detection_b200 import detector as D
    det = D.textline_detector(str(tmp_path / "x.png"), str(tmp_path), "x", str(tmp_path))
    g, n = _cases("ref_stitch_fake.npz")
    for k in range(5):
        H, W, mh, mw, nc, seed = (int(v) for v in g[f"case{k}_params"])
        lab = det.do_prediction(True, seeded_page(H, W, seed), FakeModel(mh, mw, nc))
        assert lab.dtype == np.uint8 and sha(lab[:, :, 0]) == str(g[f"case{k}_sha"])
    g, n = _cases("ref_nopatch_fake.npz")
    for k in range(n):
        H, W, mh, mw, nc, seed = (int(v) for v in g[f"case{k}_params"])
        page = seeded_page(H, W, seed)
        det.image = page
        lab = det.do_prediction(False, page, FakeModel(mh, mw, nc))
        assert lab.dtype == np.uint8 and sha(lab) == str(g[f"case{k}_sha"])
        ora = odp.do_prediction(False, page, FakeModel(mh, mw, nc), full_shape=page.shape)
        assert sha(ora) == str(g[f"case{k}_sha"])


def test_prepost_helpers_equal_reference(tmp_path):
    from sbb_textline_detection_b200 import detector as D
    g = golden("ref_prepost.npz")
    h, w, seed = (int(v) for v in g["otsu_in_seed"])
    doc = synth.document_page(h, w, seed=seed)
    want = np.unpackbits(g["otsu_packed"])[:h * w].reshape(h, w).astype(bool)
    assert bool(g["otsu_allch_equal"]) and set(g["otsu_values"].tolist()) <= {0.0, 255.0}
    o = odp.otsu_copy(doc)
    assert ((o[:, :, 0] > 0) == want).all() and (o[:, :, 1] == o[:, :, 0]).all()
    det = D.textline_detector(str(tmp_path / "x.png"), str(tmp_path), "x", str(tmp_path))
    o2 = det.otsu_copy(doc)
    assert str(o2.dtype) == str(g["otsu_dtype"]) and ((o2[:, :, 0] > 0) == want).all()
    rnd = seeded_page(333, 211, 32)
    for k in range(4):
        oh, ow = (int(v) for v in g[f"resize{k}_hw"])
        assert sha(odp.resize_nearest(rnd, oh, ow)) == str(g[f"resize{k}_sha"])
        assert sha(det.resize_image(rnd, oh, ow)) == str(g[f"resize{k}_sha"])
    for k in range(4):
        h, w, hi, wi, ho, wo = (int(v) for v in g[f"scale{k}"])
        assert odp.scaled_size(h, w) == (hi, wi)
        png = str(tmp_path / f"s{k}.png")
        cv2.imwrite(png, seeded_page(h, w, 40 + k))
        d = D.textline_detector(png, str(tmp_path), None, str(tmp_path))
        d.get_image_and_scales()
        assert (d.img_hight_int, d.img_width_int, d.height_org, d.width_org) == (hi, wi, ho, wo)
        assert [d.scale_y, d.scale_x] == g[f"scale{k}_f"].tolist()
        assert sha(d.image) == str(g[f"scale{k}_sha"]) and d.f_name == str(g[f"scale{k}_fname"])


def test_oracle_network_through_reference_loop(textline_weights):
    """Reference loop + oracle network (the fixture) == oracle loop + oracle network."""
    g = golden("ref_page96_textline.npz")
    h, w, seed = (int(v) for v in g["page_seed"])
    net = OracleNet(*textline_weights).as_keras_like(96, 96)
    lab = odp.do_prediction(True, synth.document_page(h, w, seed=seed), net, predict_batch=1)[:, :, 0]
    assert np.mean(lab != g["labels"]) <= 1e-4  # same arithmetic; slack only for BLAS thread-count effects


# ----------------------------------------------------------------------------- GPU: the product path
@pytest.mark.gpu
def test_gpu_page_vs_reference_loop(built_lib, textline_weights):
    from sbb_textline_detection_b200.model import SbbModel
    g = golden("ref_page96_textline.npz")
    h, w, seed = (int(v) for v in g["page_seed"])
    wts, nc = textline_weights
    m = SbbModel(wts, 96, 96, nc, max_batch=16)
    lab = m.predict_page(synth.document_page(h, w, seed=seed))
    m.close()
    assert lab.shape == g["labels"].shape and lab.dtype == np.uint8
    assert np.mean(lab != g["labels"]) <= 1e-3


@pytest.mark.gpu
def test_gpu_stage_drivers_vs_reference_stage_drivers(built_lib, monkeypatch, tmp_path):
    """BASELINE config 3 through the drop-in class vs the REFERENCE class' own extract_page /
    extract_text_regions / textline_contours (oracle networks plugged into the reference)."""
    from sbb_textline_detection_b200 import detector as D
    monkeypatch.setenv("SBB_SYNTHETIC_MODELS", "1")
    g = golden("ref_pipeline96.npz")
    h, w, seed = (int(v) for v in g["page_seed"])
    page = synth.document_page(h, w, seed=seed)
    det = D.textline_detector(str(tmp_path / "p.png"), str(tmp_path), "p", str(tmp_path), tile=96,
                              cache_models=False, max_batch=16)
    det.image = page.copy()
    model, sess = det.start_new_session_and_model(det.model_page_dir)
    border = det.do_prediction(False, det.image, model)
    sess.close()
    assert np.mean(border[:, :, 0] != g["border"]) <= 1e-3
    image_page, page_coord = det.extract_page()
    assert list(page_coord) == g["page_coord"].tolist()
    assert (det.cont_page[0] == g["cont_page"]).all()
    regions = det.extract_text_regions(image_page)
    assert regions.dtype == np.uint8 and regions.shape[2] == 3 and bool(g["regions_allch_equal"])
    assert (regions[:, :, 0] == regions[:, :, 1]).all()
    assert np.mean(regions[:, :, 0] != g["regions"]) <= 2e-3
    textline = det.textline_contours(image_page)
    assert textline.shape == g["textline"].shape and np.mean(textline != g["textline"]) <= 2e-3


@pytest.mark.gpu
def test_gpu_prepost_vs_reference(built_lib):
    from sbb_textline_detection_b200 import prepost
    g = golden("ref_prepost.npz")
    h, w, seed = (int(v) for v in g["otsu_in_seed"])
    doc = synth.document_page(h, w, seed=seed)
    want = np.unpackbits(g["otsu_packed"])[:h * w].reshape(h, w).astype(bool)
    o = prepost.otsu_copy(doc)
    o = o[0] if isinstance(o, tuple) else o
    assert ((o[:, :, 0] > 0) == want).all() and (o[:, :, 1] == o[:, :, 0]).all() and (o[:, :, 2] == o[:, :, 0]).all()
    rnd = seeded_page(333, 211, 32)
    for k in range(4):
        oh, ow = (int(v) for v in g[f"resize{k}_hw"])
        assert sha(prepost.resize_nearest(rnd, oh, ow)) == str(g[f"resize{k}_sha"])
