"""Full ``run()`` of the UNMODIFIED reference (CLI path main.py:2056-2157) on a synthetic document page
with ink-density stand-in models (semantic_fake.py): pins the PAGE-XML content (regions, lines) that the
bound class (sbb_textline_detection_b200.compat.bind_reference) must reproduce.

    python tests/golden/make_golden_pipeline_xml.py -> tests/golden/ref_pipeline_run.xml (+ page png seed)
"""
import os
import sys
import tempfile
import warnings

import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_import  # noqa: E402
import semantic_fake  # noqa: E402
from make_golden_xml import normalise  # noqa: E402
from sbb_textline_detection_b200 import synth  # noqa: E402

PAGE = (1400, 1000, 5)  # h, w, seed -> get_image_and_scales makes it 2800 x 2000


def main():
    warnings.filterwarnings("ignore")
    ref = ref_import.load_reference_main()
    tmp = tempfile.mkdtemp()
    png = os.path.join(tmp, "page.png")
    cv2.imwrite(png, synth.document_page(*PAGE[:2], seed=PAGE[2]))
    for f in ("model_page_mixed_best.h5", "model_strukturerkennung.h5", "model_textline_new.h5"):
        ref_import.MODEL_FACTORY[tmp + "/" + f] = (lambda p=f: semantic_fake.loader(p))
    det = ref.textline_detector(png, tmp, "page", tmp)
    det.run()
    xml = open(os.path.join(tmp, "page.xml")).read().replace(png, "page.png")
    border, regions = semantic_fake.summarise_xml(xml)
    print("regions", len(regions), "lines", sum(len(r[1]) for r in regions))
    open(os.path.join(HERE, "ref_pipeline_run.xml"), "w").write(normalise(xml))

    # the deskew inputs/outputs of that run (do_work_of_slopes main.py:1729-1738: crop of the textline mask
    # per region box, eroded twice, -> return_deskew_slope): pins the GPU deskew on real pipeline crops
    import numpy as np
    det2 = ref.textline_detector(png, tmp, "page2", tmp)
    det2.get_image_and_scales()
    image_page, _ = det2.extract_page()
    mask = det2.textline_contours(image_page)
    out = {}
    for k, (box, slope) in enumerate(zip(det.boxes, det.slopes)):
        crop, _ = det2.crop_image_inside_box(box, np.repeat(mask[:, :, np.newaxis], 3, axis=2))
        crop = cv2.erode(crop[:, :, 0], det2.kernel, iterations=2)
        again = det2.return_deskew_slope(crop, 2)
        assert again == slope, (k, again, slope)
        out[f"crop{k}_shape"] = np.array(crop.shape)
        out[f"crop{k}_bits"] = np.packbits(crop != 0)
        out[f"crop{k}_slope"] = np.float64(slope)
    np.savez_compressed(os.path.join(HERE, "ref_pipeline_deskew.npz"), n=np.array(len(det.boxes)), **out)
    print("deskew crops", len(det.boxes), [tuple(out[f"crop{k}_shape"]) for k in range(len(det.boxes))], det.slopes)


if __name__ == "__main__":
    main()
