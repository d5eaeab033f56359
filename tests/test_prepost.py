"""Pre/post byte operations against OpenCV itself -- the reference's own
dependency for these calls (main.py:112-113, 178-194, 397, 2074-2075) -- bit-exact.  The numpy
restatements in oracle/do_prediction.py are pinned against cv2 in tests/test_oracle_cpu.py."""
import numpy as np
import pytest

from sbb_textline_detection_b200 import synth

cv2 = pytest.importorskip("cv2")
KERNEL = np.ones((5, 5), np.uint8)


def test_oracle_morphology_iterations_equal_big_rectangle():
    """n iterations of the 5x5 rectangle == one (4n+1)^2 rectangle over in-bounds pixels (what the kernel computes)."""
    rng = np.random.default_rng(5)
    img = (rng.random((97, 123)) > 0.7).astype(np.uint8) * 255
    for n in (1, 3, 6):
        r = 2 * n
        pad = np.pad(img, r, constant_values=0)
        ref = np.zeros_like(img)
        for y in range(img.shape[0]):
            for x in range(0, img.shape[1], 7):
                ref[y, x] = pad[y:y + 2 * r + 1, x:x + 2 * r + 1].max()
        got = cv2.dilate(img, KERNEL, iterations=n)
        assert (got[:, ::7] == ref[:, ::7]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,out", [((2000, 1500, 3), (2800, 2100)), ((2800, 2000, 3), (448, 448)),
                                       ((448, 448, 3), (2800, 2000)), ((333, 517), (448, 448)), ((96, 96, 3), (301, 77))])
def test_resize_nearest_equals_cv2(built_lib, shape, out):
    import torch
    from sbb_textline_detection_b200 import prepost
    img = np.random.default_rng(1).integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.resize(img, (out[1], out[0]), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(prepost.resize_nearest(img, *out), ref)
    got = prepost.resize_nearest(torch.from_numpy(img).cuda(), *out)
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.gpu
def test_otsu_copy_equals_reference_code(built_lib):
    import torch
    from sbb_textline_detection_b200 import prepost
    for seed, (h, w) in enumerate([(2800, 2000), (400, 300), (97, 1031)]):
        img = synth.document_page(h, w, seed)
        thr, t1 = cv2.threshold(img[:, :, 0], 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        ref = np.stack([t1, t1, t1], axis=2)  # main.py:191-193
        got, got_thr = prepost.otsu_copy(img, return_threshold=True)
        assert got_thr == int(thr) and np.array_equal(got, ref)
        assert np.array_equal(prepost.otsu_copy(torch.from_numpy(img).cuda()).cpu().numpy(), ref)
    # flat image: no threshold separates anything (OpenCV returns 0)
    flat = np.full((64, 64, 3), 200, np.uint8)
    thr, t1 = cv2.threshold(flat[:, :, 0], 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    got, got_thr = prepost.otsu_copy(flat, return_threshold=True)
    assert got_thr == int(thr) and np.array_equal(got[:, :, 0], t1)


@pytest.mark.gpu
def test_erode_dilate_equal_cv2(built_lib):
    import torch
    from sbb_textline_detection_b200 import prepost
    rng = np.random.default_rng(2)
    lab = (rng.random((701, 533)) > 0.6).astype(np.uint8)
    lab3 = np.repeat((rng.integers(0, 4, (420, 330))[:, :, None]).astype(np.uint8), 3, axis=2)  # region label image
    for img in (lab * 255, lab3):
        for n in (1, 2, 3, 4, 6):
            assert np.array_equal(prepost.erode(img, n), cv2.erode(img, KERNEL, iterations=n)), ("erode", n)
            assert np.array_equal(prepost.dilate(img, n), cv2.dilate(img, KERNEL, iterations=n)), ("dilate", n)
    # the reference's region clean-up (main.py:2074-2075) on a device-resident label image
    d = torch.from_numpy(lab3).cuda()
    got = prepost.dilate(prepost.erode(d, 3), 4).cpu().numpy()
    assert np.array_equal(got, cv2.dilate(cv2.erode(lab3, KERNEL, iterations=3), KERNEL, iterations=4))
    # full page size (BASELINE config 2): idempotence-style property, erode <= id <= dilate
    page = (rng.random((2800, 2000)) > 0.5).astype(np.uint8) * 255
    e, dl = prepost.erode(page, 1), prepost.dilate(page, 1)
    assert (e <= page).all() and (page <= dl).all()
    assert np.array_equal(e, cv2.erode(page, KERNEL)) and np.array_equal(dl, cv2.dilate(page, KERNEL))
