"""All text-region crops of one synthetic 2800x2000 page as the reference's deskew stage sees them
(do_work_of_slopes main.py:1729-1738: crop of the textline mask per region box, eroded twice) with the slope the
UNMODIFIED reference computes for each (return_deskew_slope main.py:1601-1718; about a minute of CPU for the page).

    python tests/golden/make_golden_deskew_page.py  ->  tests/golden/ref_deskew_page55.npz
"""
import os
import sys
import tempfile
import time
import warnings

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_import  # noqa: E402
import semantic_fake  # noqa: E402
from sbb_textline_detection_b200 import synth  # noqa: E402


class FineRegions(semantic_fake.SemanticFake):
    """Smaller structuring element than semantic_fake's region stand-in: paragraph-sized regions (55 on this page)."""

    def predict(self, x):
        if self.kind != "region":
            return super().predict(x)
        x = np.asarray(x)[0]
        ink = (x.mean(axis=2) < 0.5).astype(np.uint8)
        cls = (cv2.dilate(ink, np.ones((25, 45), np.uint8)) > 0).astype(np.int64)
        return np.eye(self.nc, dtype=np.float32)[cls][None]


def main():
    warnings.filterwarnings("ignore")
    ref = ref_import.load_reference_main()
    tmp = tempfile.mkdtemp()
    png = os.path.join(tmp, "p.png")
    cv2.imwrite(png, synth.document_page(1400, 1000, seed=5))
    for f, k, nc in (("model_page_mixed_best.h5", "page", 2), ("model_strukturerkennung.h5", "region", 4),
                     ("model_textline_new.h5", "textline", 2)):
        ref_import.MODEL_FACTORY[tmp + "/" + f] = (lambda k=k, nc=nc: FineRegions(k, n_classes=nc))
    det = ref.textline_detector(png, tmp, "p", tmp)
    det.get_image_and_scales()
    image_page, _ = det.extract_page()
    tr = det.extract_text_regions(image_page)
    tr = cv2.erode(tr, det.kernel, iterations=3)
    tr = cv2.dilate(tr, det.kernel, iterations=4)
    contours = det.get_text_region_contours_and_boxes(tr)
    tl = det.textline_contours(image_page)
    out = {"n": np.array(len(det.boxes))}
    t_total = 0.0
    for k, box in enumerate(det.boxes):
        crop, _ = det.crop_image_inside_box(box, np.repeat(tl[:, :, None], 3, axis=2))
        crop = cv2.erode(crop[:, :, 0], det.kernel, iterations=2)
        t = time.time()
        slope = det.return_deskew_slope(crop, 2)
        t_total += time.time() - t
        out[f"crop{k}_shape"] = np.array(crop.shape)
        out[f"crop{k}_bits"] = np.packbits(crop != 0)
        out[f"crop{k}_slope"] = np.float64(slope)
    out["reference_cpu_seconds"] = np.float64(t_total)
    np.savez_compressed(os.path.join(HERE, "ref_deskew_page55.npz"), **out)
    print(len(det.boxes), "regions; reference deskew search", round(t_total, 1), "s")


if __name__ == "__main__":
    main()
