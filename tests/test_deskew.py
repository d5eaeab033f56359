"""Deskew search against fixtures minted by the unmodified reference
(tests/golden/make_golden_deskew.py) and against cv2 itself."""
import numpy as np
import pytest

from conftest import golden
from oracle import deskew as odk
from sbb_textline_detection_b200 import deskew

N_CASES = 9


def _case(g, k):
    h, w, skew, seed = g[f"case{k}_params"]
    h, w = int(h), int(w)
    mask = np.unpackbits(g[f"case{k}_mask"])[:h * w].reshape(h, w)
    return mask, float(g[f"case{k}_slope"])


def test_oracle_profiles_equal_reference_rotations():
    g = golden("ref_deskew.npz")
    mask, _ = _case(g, 1)
    assert (odk.rotation_profiles_cv2(mask, g["prof_angles"]) == g["prof_rowsums"]).all()


def test_host_profile_logic_reproduces_reference_slopes():
    """The product's host-side selection logic (profile statistics, NaN filtering, the reference's
    filtered-index quirk, second search range) fed with cv2-made profiles == the reference's slopes."""
    g = golden("ref_deskew.npz")
    for k in range(N_CASES):
        mask, want = _case(g, k)
        a1 = np.linspace(-25, 25, 80)
        ang = deskew._best_angle(odk.rotation_profiles_cv2(mask, a1), a1, 2)
        if abs(ang) > 15:
            a2 = np.linspace(-90, -50, 30)
            ang = deskew._best_angle(odk.rotation_profiles_cv2(mask, a2), a2, 2)
        assert ang == want, (k, ang, want)


def test_padded_geometry_and_inverse_affine():
    import cv2
    assert deskew.padded_geometry(300, 500) == (700, 350 - 150, 350 - 250)
    assert deskew.padded_geometry(421, 261)[0] == int(421 * 1.4)
    M = cv2.getRotationMatrix2D((350, 350), 7.3, 1.0)
    inv = deskew.inverse_affine(M).reshape(2, 3)
    np.testing.assert_allclose(inv, cv2.invertAffineTransform(M), rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_gpu_profiles_bit_identical_to_cv2(built_lib):
    g = golden("ref_deskew.npz")
    mask, _ = _case(g, 1)
    got = deskew.rotation_profiles(mask.astype(np.uint8), g["prof_angles"])
    assert got.dtype == np.int32 and (got == g["prof_rowsums"]).all()
    angles = np.concatenate([np.linspace(-25, 25, 80), np.linspace(-90, -50, 30)])
    rng = np.random.default_rng(0)
    for k in (0, 2, 3, 6, 7):
        mask, _ = _case(g, k)
        assert (deskew.rotation_profiles(mask.astype(np.uint8), angles) == odk.rotation_profiles_cv2(mask, angles)).all(), k
    noise = (rng.random((211, 173)) < 0.3).astype(np.uint8)  # worst case: every window is mixed
    assert (deskew.rotation_profiles(noise, angles[::5]) == odk.rotation_profiles_cv2(noise, angles[::5])).all()
    full = np.ones((90, 140), np.uint8) * 255                 # any single non-zero value
    assert (deskew.rotation_profiles(full, angles[::9]) == odk.rotation_profiles_cv2(full, angles[::9])).all()


@pytest.mark.gpu
def test_gpu_deskew_slope_equals_reference(built_lib):
    import torch
    g = golden("ref_deskew.npz")
    for k in range(N_CASES):
        mask, want = _case(g, k)
        assert deskew.return_deskew_slope(mask, 2) == want, k
    mask, want = _case(g, 2)
    assert deskew.return_deskew_slope(torch.from_numpy(mask.astype(np.uint8)).cuda(), 2) == want  # device-resident mask


@pytest.mark.gpu
def test_gpu_deskew_on_the_reference_pipeline_crops(built_lib):
    """The region crops and slopes of a full reference run() (make_golden_pipeline_xml.py): page-sized
    masks, the sizes the deskew search sees in production."""
    g = golden("ref_pipeline_deskew.npz")
    for k in range(int(g["n"])):
        h, w = (int(v) for v in g[f"crop{k}_shape"])
        crop = np.unpackbits(g[f"crop{k}_bits"])[:h * w].reshape(h, w).astype(np.uint8)
        assert deskew.return_deskew_slope(crop, 2) == float(g[f"crop{k}_slope"]), (k, h, w)


@pytest.mark.gpu
def test_gpu_deskew_whole_page_of_regions(built_lib, capsys):
    """All 55 text regions of one 2800x2000 page (make_golden_deskew_page.py): identical slope for every region;
    the reference's CPU search took ~70 s for them (recorded in the fixture)."""
    import time
    g = golden("ref_deskew_page55.npz")
    n = int(g["n"])
    crops = []
    for k in range(n):
        h, w = (int(v) for v in g[f"crop{k}_shape"])
        crops.append(np.unpackbits(g[f"crop{k}_bits"])[:h * w].reshape(h, w).astype(np.uint8))
    deskew.return_deskew_slope(crops[0], 2)  # warm-up
    t0 = time.perf_counter()
    got = [deskew.return_deskew_slope(c, 2) for c in crops]
    dt = time.perf_counter() - t0
    want = [float(g[f"crop{k}_slope"]) for k in range(n)]
    assert got == want
    with capsys.disabled():
        print(f"\n[deskew] {n} regions: GPU search {dt:.2f} s (incl. host profile statistics), "
              f"reference CPU search {float(g['reference_cpu_seconds']):.1f} s")
