"""CPU tests of the boundary: the library builds for sm_100a, loads without a GPU, exports every
symbol include/sbb_textline.h declares; host-only entry points; weight blob; host glue."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from oracle import do_prediction as odp
from sbb_textline_detection_b200 import _lib, weights
from sbb_textline_detection_b200.model import compute_tile_grid


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "sbb_textline.h")).read()
    declared = set(re.findall(r"\b(sbb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(built_lib)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.sbb_abi_version() == 1


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM (B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_tile_grid_errors(built_lib):
    with pytest.raises(RuntimeError, match="smaller than"):
        compute_tile_grid(400, 2000, 448, 448)
    with pytest.raises(RuntimeError, match="margin"):
        compute_tile_grid(2800, 2000, 448, 448, margin=224)


def _owner_map_by_replay(H, W, mh, mw, margin):
    m, nxf, nyf, tiles = odp.tile_grid(H, W, mh, mw, margin)
    out = np.full((H, W), -1, np.int64)
    for t, (i, j, x0, y0) in enumerate(tiles):
        ax = 0 if i == 0 else m
        bx = mw - m if i == 0 else (mw if i == nxf - 1 else mw - m)
        ay = 0 if j == 0 else m
        by = mh - m if j == 0 else (mh if j == nyf - 1 else mh - m)
        out[y0 + ay:y0 + by, x0 + ax:x0 + bx] = i * nyf + j
    return out, nxf, nyf, tiles


@pytest.mark.parametrize("H,W,mh,mw,margin", [(2800, 2000, 448, 448, None), (4600, 3400, 672, 672, None),
                                              (4600, 3400, 672, 672, 168), (448, 448, 448, 448, None),
                                              (449, 1000, 448, 448, None), (901, 1203, 448, 448, 0)])
def test_owner_tables_equal_loop_replay(built_lib, H, W, mh, mw, margin):
    ref, nxf, nyf, tiles = _owner_map_by_replay(H, W, mh, mw, margin)
    nx, ny, org, ox, oy = compute_tile_grid(H, W, mh, mw, -1 if margin is None else margin)
    assert (nx, ny) == (nxf, nyf)
    assert [tuple(r) for r in org.tolist()] == [(x0, y0, i, j) for (i, j, x0, y0) in tiles]
    own = ox[None, :].astype(np.int64) * ny + oy[:, None].astype(np.int64)
    own[(ox[None, :] < 0) | (oy[:, None] < 0)] = -1
    assert (own == ref).all()
    assert (ref >= 0).all()  # every pixel has an owner whenever the page is >= the tile


def test_owner_tables_random(built_lib):
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.sampled_from([32, 64, 96]), st.sampled_from([32, 64, 96]), st.integers(0, 300), st.integers(0, 300),
               st.integers(-1, 15))
    def check(mh, mw, eh, ew, margin):
        H, W = mh + eh, mw + ew
        if margin >= 0 and (mw - 2 * margin <= 0 or mh - 2 * margin <= 0):
            return
        ref, nxf, nyf, _ = _owner_map_by_replay(H, W, mh, mw, None if margin < 0 else margin)
        nx, ny, _, ox, oy = compute_tile_grid(H, W, mh, mw, margin)
        own = ox[None, :].astype(np.int64) * ny + oy[:, None].astype(np.int64)
        own[(ox[None, :] < 0) | (oy[:, None] < 0)] = -1
        assert (nx, ny) == (nxf, nyf) and (own == ref).all()

    check()


def test_blob_roundtrip_and_bn_fold(textline_weights):
    w, nc = textline_weights
    blob = weights.pack_blob(w, nc)
    n2, recs = weights.unpack_blob(blob)
    assert n2 == nc and len(recs) == 62 and recs[0][0] == "conv1" and recs[1][0] == "bn_conv1"
    # folded conv == conv + BN on a random input (one mid layer)
    name, kh, kw, cin, cout, wt, bias = next(r for r in recs if r[0] == "res3b_branch2b")
    x = torch.randn(1, cin, 9, 9)
    k = torch.from_numpy(w[name + "/kernel"]).permute(3, 2, 0, 1)
    y = torch.nn.functional.conv2d(x, k, torch.from_numpy(w[name + "/bias"]), padding=1)
    bn = "bn3b_branch2b"
    s = torch.from_numpy(w[bn + "/gamma"] / np.sqrt(w[bn + "/var"] + 1e-3))
    y = (y - torch.from_numpy(w[bn + "/mean"])[None, :, None, None]) * s[None, :, None, None] + \
        torch.from_numpy(w[bn + "/beta"])[None, :, None, None]
    kf = torch.from_numpy(wt.reshape(cout, kh, kw, cin).copy()).permute(0, 3, 1, 2)
    y2 = torch.nn.functional.conv2d(x, kf, torch.from_numpy(bias.copy()), padding=1)
    assert (y - y2).abs().max() < 1e-4


def test_weights_are_reproducible():
    a = weights.random_init(1234, 2)
    b = weights.random_init(1234, 4)
    assert (a["res4c_branch2b/kernel"] == b["res4c_branch2b/kernel"]).all()
    assert a["cls/kernel"].shape == (1, 1, 32, 2) and b["cls/kernel"].shape == (1, 1, 32, 4)
    assert abs(float(a["conv1/kernel"].std()) - np.sqrt(2 / 147)) < 0.01


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


def test_detector_generic_path_equals_oracle_replay(tmp_path):
    """The drop-in class, driven with a duck-typed (non-GPU) model, reproduces the oracle replay of
    main.py:231-366 -- the compat path a real Keras model would take."""
    cv2 = pytest.importorskip("cv2")
    from sbb_textline_detection_b200.detector import textline_detector
    mh = mw = 64

    class Fake:
        layers = [type("L", (), {"output_shape": (None, mh, mw, 3)})()]

        def predict(self, x):
            s = x.sum(-1)
            cls = (np.floor(s * 255.0 + 0.5).astype(np.int64) + np.arange(mw)[None, None, :]) % 3
            return np.eye(3, dtype=np.float32)[cls]

    rng = np.random.default_rng(9)
    page = rng.integers(0, 256, (200, 170, 3), dtype=np.uint8)
    det = textline_detector("x.png", str(tmp_path), "x", str(tmp_path))
    got = det.do_prediction(True, page, Fake())
    ref = odp.do_prediction(True, page, Fake())
    assert got.dtype == np.uint8 and (got == ref).all()
    det.image = page
    got = det.do_prediction(False, page, Fake())
    ref = odp.do_prediction(False, page, Fake(), full_shape=page.shape)
    assert (got == ref).all()
