"""Golden vectors for the deskew search, minted by EXECUTING THE UNMODIFIED REFERENCE
(``textline_detector.return_deskew_slope`` / ``rotate_image``, main.py:159-163, 1545-1718; pure
cv2/numpy/scipy, imported in place through ref_import.py).

    python tests/golden/make_golden_deskew.py   ->  tests/golden/ref_deskew.npz
"""
import os
import sys
import tempfile
import warnings

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def line_mask(h, w, skew_deg, seed, pitch=28, thick=9):
    """Binary 'textline mask' of a text region: horizontal bars with ragged ends, rotated by skew."""
    rng = np.random.default_rng(seed)
    big = np.zeros((h * 2, w * 2), np.uint8)
    for y in range(pitch, 2 * h - pitch, pitch):
        x0 = rng.integers(w // 3, w // 2)
        x1 = rng.integers(3 * w // 2, 5 * w // 3)
        big[y:y + thick, x0:x1] = 1
    M = cv2.getRotationMatrix2D((w, h), skew_deg, 1.0)
    rot = cv2.warpAffine(big, M, (2 * w, 2 * h), flags=cv2.INTER_NEAREST)
    return np.ascontiguousarray(rot[h // 2:h // 2 + h, w // 2:w // 2 + w])


CASES = [  # h, w, skew, seed
    (300, 500, 0.0, 1), (300, 500, 3.0, 2), (420, 260, -7.0, 3), (150, 700, 12.0, 4),
    (260, 260, -20.0, 5), (333, 411, 24.0, 6), (500, 380, 70.0, 7), (64, 300, 1.5, 8), (200, 320, 0.0, 9),
]


def main():
    warnings.filterwarnings("ignore")
    ref = ref_import.load_reference_main()
    tmp = tempfile.mkdtemp()
    det = ref.textline_detector(os.path.join(tmp, "x.png"), tmp, "x", tmp)
    out = {}
    for k, (h, w, skew, seed) in enumerate(CASES):
        m = line_mask(h, w, skew, seed)
        if k == len(CASES) - 1:
            m[:] = 0  # empty region
        slope = det.return_deskew_slope(m, 2)
        out[f"case{k}_params"] = np.array([h, w, skew, seed], np.float64)
        out[f"case{k}_mask"] = np.packbits(m)
        out[f"case{k}_slope"] = np.float64(slope)
        print((h, w, skew), "->", slope)
    # a few raw rotations for the profile check: row sums of the binarised reference rotate_image output
    m = line_mask(300, 500, 3.0, 2)
    side = int(500 * 1.4)
    pad = np.zeros((side, side))
    c = int(side / 2.)
    pad[c - 150:c - 150 + 300, c - 250:c - 250 + 500] = m
    angs = np.array([-25.0, -3.1645569620253156, 0.0, 7.3, 25.0, -90.0, -61.5])
    prof = []
    for a in angs:
        r = det.rotate_image(pad, a)
        r[r != 0] = 1
        prof.append(r.sum(axis=1))
    out["prof_angles"] = angs
    out["prof_rowsums"] = np.array(prof).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_deskew.npz"), **out)


if __name__ == "__main__":
    main()
