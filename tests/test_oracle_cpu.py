"""CPU tests: the oracle against the committed golden vectors, the numpy do_prediction replay,
and the small helpers around the hot path.  (The network oracle is unpinned -- the reference has no
tests or weights; the host-side oracles are pinned in tests/test_reference_golden.py.)"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import do_prediction as odp
from oracle.resnet50_unet import OracleNet, conv_flops, conv_specs
from sbb_textline_detection_b200 import arch


def test_conv_flops_match_survey():
    tot, enc, dec = conv_flops(448, 448, 2)
    assert abs(tot - 87.85e9) < 0.01e9 and abs(enc - 30.75e9) < 0.01e9 and abs(dec - 57.10e9) < 0.01e9
    assert abs(conv_flops(672, 672, 2)[0] - 197.73e9) < 0.02e9
    assert arch.conv_flops_per_tile(448, 448, 2) == (tot, enc, dec)
    assert len(conv_specs(2)) == 61
    assert sum(kh * kw * ci * co for _, _, kh, kw, ci, co in conv_specs(2)) == pytest.approx(32.85e6, rel=2e-3)


def test_oracle_shapes_and_softmax(textline_weights):
    w, nc = textline_weights
    net = OracleNet(w, nc)
    net.taps = {}
    x = np.random.default_rng(0).random((1, 64, 96, 3), dtype=np.float32)
    p = net.predict(x)
    assert p.shape == (1, 64, 96, 2) and p.dtype == np.float32
    np.testing.assert_allclose(p.sum(-1), 1.0, atol=1e-6)
    g = arch.tile_geometry(64, 96)
    assert net.taps["conv1"].shape[2:] == g[1] and net.taps["pool1"].shape[2:] == g[2]
    assert net.taps["res5c"].shape[1:] == (2048,) + g[5]
    assert net.taps["dec3"].shape[2:] == (g[2][0] + 1, g[2][1] + 1)  # one_side_pad: 2*g[3]


def test_oracle_vs_golden_tile96(textline_weights):
    w, nc = textline_weights
    g = golden("tile96_textline.npz")
    z = OracleNet(w, nc).logits(g["tile"][None].astype(np.float32) / np.float32(255)).numpy()[0]
    assert np.abs(z - g["logits"]).max() < 2e-4          # fp32 reassociation across machines
    assert np.mean(z.argmax(-1) != g["labels"]) < 1e-3


def test_oracle_fp64_agrees_with_fp32(textline_weights):
    w, nc = textline_weights
    g = golden("tile96_textline.npz")
    x = g["tile"][None].astype(np.float64) / 255.0
    z64 = OracleNet(w, nc, torch.float64).logits(x).numpy()[0]
    assert np.abs(z64 - g["logits"]).max() < 5e-4


def test_div255_is_one_fp32_division():
    v = np.arange(256)
    a = (v.astype(np.float64) / 255.0).astype(np.float32)       # main.py:239 then Keras' fp32 cast
    b = v.astype(np.float32) / np.float32(255.0)                # what the kernels compute
    assert (a == b).all()


def _fake_seg(t, i, j, x0, y0, mh, mw):
    yy, xx = np.mgrid[0:mh, 0:mw]
    return ((yy * 7 + xx * 13 + t * 31 + i * 3 + j * 5) % 251).astype(np.int64)


def test_stitch_replay_vs_golden():
    g = golden("stitch_hash.npz")
    k = 0
    while f"c{k}_H" in g.files:
        H, W, mh, mw, margin = (int(g[f"c{k}_{n}"]) for n in ("H", "W", "mh", "mw", "margin"))
        m, nxf, nyf, tiles = odp.tile_grid(H, W, mh, mw, None if margin < 0 else margin)
        assert (nxf, nyf) == (int(g[f"c{k}_nxf"]), int(g[f"c{k}_nyf"]))
        out = odp.stitch_replay(H, W, mh, mw, m, nxf, nyf, tiles,
                                lambda t, i, j, x0, y0: _fake_seg(t, i, j, x0, y0, mh, mw))[:, :, 0]
        assert (out.sum(1, dtype=np.int64) == g[f"c{k}_rowsum"]).all()
        assert (out.sum(0, dtype=np.int64) == g[f"c{k}_colsum"]).all()
        assert (out[::37, ::41] == g[f"c{k}_sample"]).all()
        k += 1
    assert k == 6


def test_grids_of_baseline_configs():
    _, nxf, nyf, tiles = odp.tile_grid(2800, 2000, 448, 448)
    assert (nxf, nyf, len(tiles)) == (6, 8, 48)
    assert sorted({t[2] for t in tiles}) == [0, 360, 720, 1080, 1440, 1552]
    assert sorted({t[3] for t in tiles}) == [0, 360, 720, 1080, 1440, 1800, 2160, 2352]
    assert odp.tile_grid(4600, 3400, 672, 672)[1:3] == (7, 9)
    assert odp.tile_grid(4600, 3400, 672, 672, 168)[1:3] == (11, 14)


def test_do_prediction_with_fake_model_is_position_exact():
    """A model whose output encodes the absolute page position of every tile pixel: the stitched map
    must then equal the same function of the page position wherever a tile wrote."""
    mh = mw = 64

    class Fake:
        layers = [type("L", (), {"output_shape": (None, mh, mw, 4)})()]

        def predict(self, x):
            cls = np.floor(x[..., 0] * 255.0 + 0.5).astype(np.int64) % 4
            return np.eye(4, dtype=np.float32)[cls]

    rng = np.random.default_rng(3)
    page = rng.integers(0, 256, (150, 131, 3), dtype=np.uint8)
    out = odp.do_prediction(True, page, Fake())
    assert out.dtype == np.uint8 and out.shape == (150, 131, 3)
    assert (out[:, :, 0] == page[:, :, 0] % 4).all() and (out[:, :, 0] == out[:, :, 2]).all()


def test_resize_nearest_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for (h, w, oh, ow) in [(2800, 2000, 448, 448), (448, 448, 2800, 2000), (333, 517, 448, 448), (448, 448, 3361, 2417)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert (odp.resize_nearest(img, oh, ow) == cv2.resize(img, (ow, oh), interpolation=cv2.INTER_NEAREST)).all()


def test_otsu_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    from sbb_textline_detection_b200 import synth
    for seed in range(3):
        img = synth.document_page(400, 300, seed)
        thr, ref = cv2.threshold(img[:, :, 0], 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert odp.otsu_threshold_u8(img[:, :, 0]) == int(thr)
        assert (odp.otsu_copy(img)[:, :, 1] == ref).all()


def test_scaled_size():
    assert odp.scaled_size(2000, 1500) == (2800, 2100)
    assert odp.scaled_size(4600, 3400) == (5520, 4080)
