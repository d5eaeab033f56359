"""ORACLE (test infrastructure, NOT product code) for the deskew search: the reference's own statements
(main.py:159-163 rotate_image, :1604-1617 padding, :1628-1629 + :1546 binarise and sum along x) executed
with cv2 on the CPU.  Pinned against the reference itself by tests/golden/ref_deskew.npz
(make_golden_deskew.py runs the unmodified reference).  Only tests/ may import this."""
import cv2
import numpy as np


def padded_patch(img_patch: np.ndarray) -> np.ndarray:
    """main.py:1602-1617"""
    h, w = img_patch.shape[:2]
    side = int(max(h, w) * 1.4)
    padded = np.zeros((side, side))
    c = int(side / 2.)
    hy, hx = int(h / 2.), int(w / 2.)
    padded[c - hy:c - hy + h, c - hx:c - hx + w] = img_patch
    return padded


def rotate_image(img: np.ndarray, slope: float) -> np.ndarray:
    """main.py:159-163"""
    h, w = img.shape[:2]
    M = cv2.getRotationMatrix2D((w // 2, h // 2), slope, 1.0)
    return cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)


def rotation_profiles_cv2(img_patch: np.ndarray, angles) -> np.ndarray:
    padded = padded_patch(img_patch)
    out = []
    for a in angles:
        r = rotate_image(padded, float(a))
        r[r != 0] = 1
        out.append(r.sum(axis=1))
    return np.array(out).astype(np.int32)
