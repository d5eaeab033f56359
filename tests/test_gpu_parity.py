"""GPU parity tests (B200): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full page size
-- through size-independent properties (stitch consistency, idempotence).

Tolerances (BASELINE.json north_star): |logit| error <= 1e-3, label-map IoU >= 0.999."""
import numpy as np
import pytest
import torch

from conftest import golden, iou
from oracle import do_prediction as odp
from oracle.resnet50_unet import OracleNet
from sbb_textline_detection_b200 import synth
from sbb_textline_detection_b200.model import SbbModel

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3
IOU_MIN = 0.999


@pytest.fixture(scope="module")
def tiles448():
    page = synth.document_page(2800, 2000, seed=0)
    return np.stack([page[360:808, 360:808], synth.uniform_page(448, 448, 0)]).astype(np.float32) / np.float32(255)


@pytest.fixture(scope="module")
def oracle448(textline_weights, tiles448):
    w, nc = textline_weights
    net = OracleNet(w, nc, torch.float32)
    net.taps = {}
    with torch.no_grad():
        z = net.logits(tiles448).numpy()
    return z, {k: v.permute(0, 2, 3, 1).numpy() for k, v in net.taps.items()}


@pytest.fixture(scope="module")
def model448(built_lib, textline_weights):
    w, nc = textline_weights
    m = SbbModel(w, 448, 448, nc, max_batch=48)
    yield m
    m.close()


def test_tile_logits_labels_probs_vs_oracle(model448, tiles448, oracle448):
    z_ref, _ = oracle448
    labels, probs, logits = model448.predict_tiles(tiles448, True, True, True)
    assert model448.last_launch_count() == 58  # 3 stem + 53 convs + dec1..dec5 (4 parity variants each)
    assert np.abs(logits - z_ref).max() <= LOGIT_TOL
    p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()
    assert np.abs(probs - p_ref).max() <= 5e-4
    ref_lab = z_ref.argmax(-1)
    assert np.mean(labels != ref_lab) <= 2e-4
    assert iou(labels, ref_lab) >= IOU_MIN
    # keras-compatible entry point returns the same probabilities
    assert np.array_equal(model448.predict(tiles448[:1]), probs[:1])


def test_every_layer_vs_oracle(model448, tiles448, oracle448):
    _, taps = oracle448
    model448.predict_tiles(tiles448, True, False, False)
    for i, (name, h, w, c) in enumerate(model448.activations()):
        for t in range(2):
            a = model448.read_activation(i, t)
            ref = taps[name][t]
            assert a.shape == ref.shape
            # fp32 re-ordering noise alone reaches 1.6e-4 relative at the deep layers (exact-fp32 SIMT
            # backend, profiles/r01_diag1_first_gpu_run.txt); the end-to-end gate is LOGIT_TOL above
            assert np.abs(a - ref).max() <= 3e-4 * max(1.0, np.abs(ref).max()), name


def test_golden_tile448(model448):
    g = golden("tile448_textline.npz")
    x = g["tile"][None].astype(np.float32) / np.float32(255)
    labels, _, logits = model448.predict_tiles(x, True, False, True)
    ref = np.unpackbits(g["labels_packed"])[:448 * 448].reshape(448, 448)
    assert np.abs(logits[0][g["ys"], g["xs"]] - g["logits_sampled"]).max() <= LOGIT_TOL
    assert iou(labels[0], ref) >= IOU_MIN


def test_simt_crosscheck_and_golden_tile96(built_lib, textline_weights):
    w, nc = textline_weights
    g = golden("tile96_textline.npz")
    x = g["tile"][None].astype(np.float32) / np.float32(255)
    out = {}
    for backend in ("tcgen05", "simt"):
        m = SbbModel(w, 96, 96, nc, backend=backend, max_batch=2)
        out[backend] = m.predict_tiles(x, True, False, True)
        m.close()
    for backend in out:
        assert np.abs(out[backend][2][0] - g["logits"]).max() <= LOGIT_TOL, backend
    assert np.abs(out["tcgen05"][2] - out["simt"][2]).max() <= LOGIT_TOL


def test_decoder_plan_variants_agree(built_lib, textline_weights, tiles448, oracle448, monkeypatch):
    """The decoder has two plan-time choices that must not change the result: dec5 as ONE merged-parity
    N = 128 GEMM (default) or as four N = 32 output-parity variants, and the M-tile shapes chosen for the
    kept regions (default) or for the full grid; a third knob adds identity shortcuts in the epilogue instead
    of as a K segment of the MMA.  All alternatives stay within the oracle tolerance and give the same page
    label map as the default plan up to the fp32 summation order."""
    w, nc = textline_weights
    z_ref, _ = oracle448
    page = synth.document_page(1000, 900, seed=5)
    outs = {}
    for name, env in (("default", {}), ("parity_variants", {"SBB_DEC5_MERGED": "0"}),
                      ("full_grid_shapes", {"SBB_DEC5_MERGED": "0", "SBB_DEC_RECT": "0"}),
                      ("epilogue_residual", {"SBB_RES_IN_MMA": "0"})):
        with monkeypatch.context() as mp:
            for k, v in env.items():
                mp.setenv(k, v)
            m = SbbModel(w, 448, 448, nc, max_batch=12)
        logits = m.predict_tiles(tiles448, False, False, True)[2]
        outs[name] = (logits, m.predict_page(page))
        m.close()
        assert np.abs(logits - z_ref).max() <= LOGIT_TOL, name
    for name in ("parity_variants", "full_grid_shapes", "epilogue_residual"):
        assert np.abs(outs[name][0] - outs["default"][0]).max() <= 2e-4, name
        assert np.mean(outs[name][1] != outs["default"][1]) <= 1e-4, name


def test_golden_page96_region_model(built_lib, region_weights):
    """4-class region model, 96x96 tiles, whole do_prediction(patches=True) against the golden map."""
    w, nc = region_weights
    g = golden("page96_region.npz")
    m = SbbModel(w, 96, 96, nc, max_batch=5)  # 4x4 = 16 tiles -> several partial batches
    lab = m.predict_page(g["page"])
    m.close()
    assert lab.shape == g["labels"].shape and lab.dtype == np.uint8
    assert np.mean(lab != g["labels"]) <= 1e-3
    for cls in range(nc):
        if (g["labels"] == cls).sum() > 200:
            assert iou(lab, g["labels"], cls) >= 0.995


def test_page_mode_vs_oracle_do_prediction(model448, textline_weights):
    w, nc = textline_weights
    page = synth.document_page(1000, 900, seed=3)
    lab = model448.predict_page(page)
    ref = odp.do_prediction(True, page, OracleNet(w, nc).as_keras_like(448, 448), predict_batch=3)[:, :, 0]
    assert np.mean(lab != ref) <= 3e-4
    assert iou(lab, ref) >= IOU_MIN


def test_full_page_stitch_property(model448):
    """2800x2000 (BASELINE config 2): the fused page call must equal the reference's loop replay fed
    with the SAME GPU path's per-tile labels -- bit-exact, independent of the oracle's speed."""
    page = synth.document_page(2800, 2000, seed=1)
    lab = model448.predict_page(page)
    m, nxf, nyf, tiles = odp.tile_grid(2800, 2000, 448, 448)
    assert (nxf, nyf) == (6, 8)
    x = np.stack([page[y0:y0 + 448, x0:x0 + 448] for (_, _, x0, y0) in tiles]).astype(np.float32) / np.float32(255)
    tl, _, _ = model448.predict_tiles(x, True, False, False)
    ref = odp.stitch_replay(2800, 2000, 448, 448, m, nxf, nyf, tiles, lambda t, *_: tl[t].astype(np.int64))[:, :, 0]
    assert np.array_equal(lab, ref)
    assert np.array_equal(lab, model448.predict_page(page))  # idempotent / deterministic
    assert 0.02 < lab.mean() < 0.98


def test_device_resident_call_equals_host_call(model448):
    page = synth.document_page(1200, 1000, seed=2)
    host = model448.predict_page(page)
    dpage = torch.from_numpy(page).cuda()
    out = model448.predict_page(dpage, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(host, out.cpu().numpy())


def test_explicit_margin_matches_replay(model448):
    """BASELINE config 5 asks for 50 % overlap, i.e. a margin that is not int(0.1*tile)."""
    page = synth.document_page(900, 1000, seed=4)
    margin = 112
    lab = model448.predict_page(page, margin=margin)
    m, nxf, nyf, tiles = odp.tile_grid(900, 1000, 448, 448, margin)
    x = np.stack([page[y0:y0 + 448, x0:x0 + 448] for (_, _, x0, y0) in tiles]).astype(np.float32) / np.float32(255)
    tl, _, _ = model448.predict_tiles(x, True, False, False)
    ref = odp.stitch_replay(900, 1000, 448, 448, m, nxf, nyf, tiles, lambda t, *_: tl[t].astype(np.int64))[:, :, 0]
    assert np.array_equal(lab, ref)


def test_no_patch_path_vs_oracle(built_lib):
    """do_prediction(patches=False) as extract_page uses it (main.py:368-379), through the drop-in class."""
    from sbb_textline_detection_b200.detector import synthetic_weights, textline_detector
    w, nc = synthetic_weights("page")
    page = synth.document_page(1400, 1000, seed=5)
    det = textline_detector("x.png", "/tmp", "x", "/tmp")
    det.image = page
    m = SbbModel(w, 448, 448, nc, max_batch=1)
    got = det.do_prediction(False, page, m)
    m.close()
    ref = odp.do_prediction(False, page, OracleNet(w, nc).as_keras_like(448, 448), full_shape=page.shape)
    assert got.shape == ref.shape == (1400, 1000, 3) and got.dtype == np.uint8
    assert np.mean(got != ref) <= 3e-4


def test_tile672_model(built_lib, textline_weights):
    """BASELINE config 5's tile size: 672x672 (odd stage-2 size 167, one_side_pad to 168)."""
    w, nc = textline_weights
    x = (synth.document_page(672, 672, seed=6)[None].astype(np.float32)) / np.float32(255)
    z_ref = OracleNet(w, nc).logits(x).numpy()
    m = SbbModel(w, 672, 672, nc, max_batch=1)
    labels, _, logits = m.predict_tiles(x, True, False, True)
    m.close()
    assert np.abs(logits - z_ref).max() <= LOGIT_TOL
    assert iou(labels, z_ref.argmax(-1)) >= IOU_MIN


def test_error_paths(model448):
    with pytest.raises(RuntimeError, match="smaller than"):
        model448.predict_page(np.zeros((400, 2000, 3), np.uint8))
    with pytest.raises(RuntimeError, match="tile size"):
        SbbModel(b"SBBW0001" + b"\0" * 16, 450, 448, 2)
    with pytest.raises(RuntimeError, match="magic"):
        SbbModel(b"garbage-garbage-garbage-garbage", 448, 448, 2)


def test_fp16_fast_mode_runs_but_is_not_the_parity_mode(built_lib, textline_weights, tiles448, oracle448):
    """Single-fp16 operands: documented as outside the reference tolerance; it must run
    and be in the right ballpark, nothing more."""
    w, nc = textline_weights
    z_ref, _ = oracle448
    m = SbbModel(w, 448, 448, nc, precision="fp16", max_batch=2)
    labels, _, logits = m.predict_tiles(tiles448, True, False, True)
    m.close()
    assert np.abs(logits - z_ref).max() < 2.0
    assert iou(labels, z_ref.argmax(-1)) > 0.9


def test_three_model_pipeline_vs_oracle(built_lib, monkeypatch, tmp_path):
    """BASELINE config 3: border (patches=False) + region (Otsu, 4 classes) + textline stages through the
    drop-in class' own stage drivers (main.py:384-503), each against the oracle fed with the same image.
    Small tile (96) so that the CPU oracle finishes in seconds."""
    import cv2
    from sbb_textline_detection_b200 import detector as D
    monkeypatch.setenv("SBB_SYNTHETIC_MODELS", "1")
    page = synth.document_page(420, 330, seed=7)
    png = str(tmp_path / "page.png")
    cv2.imwrite(png, page)
    det = D.textline_detector(png, str(tmp_path), "page", str(tmp_path), tile=96, cache_models=False, max_batch=16)
    det.image = page  # get_image_and_scales would blow the page up to 2800 rows: keep the oracle affordable
    T = 96
    nets = {k: OracleNet(*D.synthetic_weights(k)).as_keras_like(T, T) for k in ("page", "region", "textline")}
    # stage 1: extract_page -> crop box from the border model's label map
    image_page, page_coord = det.extract_page()
    assert not hasattr(det, "image")   # main.py:431 deletes the attribute; do_prediction(False) reads its shape (main.py:378)
    det.image = page
    ref_page = odp.do_prediction(False, page, nets["page"], full_shape=page.shape)
    got_page = det.do_prediction(False, page, det.start_new_session_and_model(det.model_page_dir)[0])
    assert np.mean(got_page != ref_page) <= 1e-3
    assert image_page.shape[2] == 3 and len(page_coord) == 4
    # stage 2: region model on the Otsu-binarised crop (channel-0 threshold in all channels, main.py:187-193)
    regions = det.extract_text_regions(image_page)
    ref_regions = odp.do_prediction(True, odp.otsu_copy(image_page), nets["region"], predict_batch=16)
    assert regions.shape == ref_regions.shape and regions.dtype == np.uint8
    assert np.mean(regions != ref_regions) <= 2e-3
    # stage 3: textline model on the raw crop, channel 0 returned (main.py:503)
    textline = det.textline_contours(image_page)
    ref_textline = odp.do_prediction(True, image_page, nets["textline"], predict_batch=16)[:, :, 0]
    assert textline.shape == ref_textline.shape
    assert np.mean(textline != ref_textline) <= 2e-3
    # and the whole run_segmentation() entry (imread + scale rule + 3 stages) executes end to end
    det2 = D.textline_detector(png, str(tmp_path), "page", str(tmp_path), tile=96, cache_models=False, max_batch=48)
    det2.get_image_and_scales()
    assert det2.image.shape[0] == 2800  # main.py:201-203
    crop, coord = det2.extract_page()
    if min(crop.shape[:2]) >= T:  # the synthetic border model may crop to less than one tile
        reg, tl = det2.extract_text_regions(crop), det2.textline_contours(crop)
        assert reg.shape[:2] == tl.shape == (coord[1] - coord[0], coord[3] - coord[2])


def test_pipelined_batch_call_equals_single_calls(model448):
    """predict_pages (H2D / forward / D2H of neighbouring pages overlapped on three streams) returns
    exactly what one blocking predict_page per page returns; mixed page sizes, odd and even counts."""
    pages = [synth.document_page(700 + 50 * k, 600 + 40 * (k % 2), seed=60 + k) for k in range(5)]
    want = [model448.predict_page(p) for p in pages]
    got = model448.predict_pages(pages)
    assert len(got) == 5
    for g, w in zip(got, want):
        assert g.dtype == np.uint8 and (g == w).all()
    got1 = model448.predict_pages(pages[:1])
    assert (got1[0] == want[0]).all()


def _stitch_of_tiles(model, page, T, margin=None):
    m, nxf, nyf, tiles = odp.tile_grid(page.shape[0], page.shape[1], T, T, margin)
    x = np.stack([page[y0:y0 + T, x0:x0 + T] for (_, _, x0, y0) in tiles]).astype(np.float32) / np.float32(255)
    tl = np.concatenate([model.predict_tiles(x[k:k + model.max_batch], True, False, False)[0]
                         for k in range(0, len(x), model.max_batch)])
    return odp.stitch_replay(page.shape[0], page.shape[1], T, T, m, nxf, nyf, tiles, lambda t, *_: tl[t].astype(np.int64))[:, :, 0]


@pytest.mark.parametrize("H,W", [(96, 96), (97, 96), (96, 131), (193, 77 + 96), (300, 77 * 3), (77 * 2 + 1, 500)])
def test_ragged_and_minimal_pages_stitch_property(built_lib, textline_weights, H, W):
    """Edge geometries of the tiler (main.py:246-281): a page of exactly one tile (2x2 tiles all clamped to the
    origin), one pixel more than a tile, widths/heights just past a multiple of the stride (several clamped
    trailing tiles).  Page call == replay of the reference loop over the same path's per-tile labels, bit-exact."""
    w, nc = textline_weights
    m = SbbModel(w, 96, 96, nc, max_batch=7)
    page = synth.document_page(max(H, 200), max(W, 200), seed=H * 1000 + W)[:H, :W]
    page = np.ascontiguousarray(page)
    got = m.predict_page(page)
    ref = _stitch_of_tiles(m, page, 96)
    m.close()
    assert got.shape == (H, W) and np.array_equal(got, ref)


def test_config5_full_size_stitch_property(built_lib, textline_weights):
    """BASELINE config 5 at full size: 4600x3400 page, 672x672 tiles, the reference's margin rule (7x9 = 63 tiles,
    more than one batch).  Bit-exact against the loop replay fed with the same path's per-tile labels."""
    w, nc = textline_weights
    m = SbbModel(w, 672, 672, nc, max_batch=24)
    page = synth.document_page(4600, 3400, seed=5)
    got = m.predict_page(page)
    ref = _stitch_of_tiles(m, page, 672)
    assert np.array_equal(got, ref)
    assert np.array_equal(got, m.predict_page(page, margin=67))   # -1 == int(0.1 * 672)
    m.close()


def test_region_model_logits_448(built_lib, region_weights):
    """The 4-class region model at the production tile size: logits within tolerance, first-max argmax."""
    w, nc = region_weights
    x = (synth.document_page(448, 448, seed=9)[None].astype(np.float32)) / np.float32(255)
    z_ref = OracleNet(w, nc).logits(x).numpy()
    m = SbbModel(w, 448, 448, nc, max_batch=1)
    labels, probs, logits = m.predict_tiles(x, True, True, True)
    m.close()
    assert nc == 4 and logits.shape == (1, 448, 448, 4)
    assert np.abs(logits - z_ref).max() <= LOGIT_TOL
    assert np.mean(labels != z_ref.argmax(-1)) <= 1e-3
    assert np.abs(probs.sum(-1) - 1).max() < 1e-5


def test_strided_inputs_and_outputs_equal_contiguous(model448):
    """Row strides larger than the row (a crop of a bigger page, as the stage drivers pass after the border crop):
    host and device inputs, strided device output."""
    big = synth.document_page(1100, 1300, seed=8)
    crop = big[37:37 + 900, 101:101 + 1000]                       # numpy view: row stride 3900 bytes
    assert not crop.flags["C_CONTIGUOUS"]
    want = model448.predict_page(np.ascontiguousarray(crop))
    dbig = torch.from_numpy(big).cuda()
    dcrop = dbig[37:37 + 900, 101:101 + 1000]                     # device view with the same strides
    got_dev = model448.predict_page(dcrop)
    assert np.array_equal(got_dev.cpu().numpy(), want)
    canvas = torch.full((1000, 1200), 7, dtype=torch.uint8, device="cuda")
    out_view = canvas[50:950, 100:1100]                            # strided output: only the view is written
    model448.predict_page(dcrop, out=out_view)
    torch.cuda.synchronize()
    assert np.array_equal(out_view.cpu().numpy(), want)
    c = canvas.cpu().numpy()
    assert (c[:50] == 7).all() and (c[950:] == 7).all() and (c[:, :100] == 7).all() and (c[:, 1100:] == 7).all()


def test_stage_drivers_on_a_cropped_page(built_lib, monkeypatch, tmp_path):
    """extract_text_regions / textline_contours on a real crop of the device-resident page (offset view) equal the
    same calls on a contiguous copy of that crop."""
    from sbb_textline_detection_b200 import detector as D
    monkeypatch.setenv("SBB_SYNTHETIC_MODELS", "1")
    page = synth.document_page(520, 430, seed=17)
    det = D.textline_detector(str(tmp_path / "p.png"), str(tmp_path), "p", str(tmp_path), tile=96, cache_models=False, max_batch=24)
    det.image = page
    det._device_page()
    crop, _ = det.crop_image_inside_box([33, 21, 350, 440], page)   # x, y, w, h -> a view into `page`
    assert np.shares_memory(crop, page) and not crop.flags["C_CONTIGUOUS"]
    assert det._device_view(crop).data_ptr() != det._dev_page.data_ptr()  # really the offset twin, not an upload
    reg_v, tl_v = det.extract_text_regions(crop), det.textline_contours(crop)
    copy = np.ascontiguousarray(crop)
    reg_c, tl_c = det.extract_text_regions(copy), det.textline_contours(copy)
    assert np.array_equal(reg_v, reg_c) and np.array_equal(tl_v, tl_c)
    assert reg_v.shape == (440, 350, 3) and tl_v.shape == (440, 350)
