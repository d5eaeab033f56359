"""Full ``run()`` of the UNMODIFIED reference (main.py:2056-2157) with the ORACLE networks plugged in -- the real
ResNet50-U-Net arithmetic on the CPU, document-like synthetic weights (sbb_textline_detection_b200/semantic.py) --
on a synthetic scan with a dark scanner border.  Pins, for the GPU path and the bound class:

  * the three label maps the reference's stage drivers obtain (border crop box, region labels, textline mask)
  * what its host glue makes of them: region boxes, deskew slope per region, PAGE-XML (regions and lines)

    python tests/golden/make_golden_semantic_run.py   -> tests/golden/ref_semantic_run.{npz,xml}
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
from make_golden_xml import normalise  # noqa: E402
from oracle.resnet50_unet import OracleNet  # noqa: E402
from sbb_textline_detection_b200 import semantic, synth  # noqa: E402

PAGE = (1400, 1000, 31, 60)  # h, w, seed, scanner border -> get_image_and_scales makes it 2800 x 2000
TILE = 448
FILES = {"model_page_mixed_best.h5": "page", "model_strukturerkennung.h5": "region", "model_textline_new.h5": "textline"}


def oracle_models():
    return {f: OracleNet(*semantic.semantic_weights(k)).as_keras_like(TILE, TILE) for f, k in FILES.items()}


def main():
    warnings.filterwarnings("ignore")
    ref = ref_import.load_reference_main()
    tmp = tempfile.mkdtemp()
    png = os.path.join(tmp, "page.png")
    cv2.imwrite(png, synth.framed_page(*PAGE[:2], seed=PAGE[2], frame=PAGE[3]))
    models = oracle_models()
    for f in FILES:
        ref_import.MODEL_FACTORY[tmp + "/" + f] = (lambda m=models[f]: m)
    # record what the stage drivers return inside run()
    seen = {}
    cls = ref.textline_detector

    class Recording(cls):
        def extract_text_regions(self, img):
            seen["image_page_shape"] = np.array(img.shape)
            out = cls.extract_text_regions(self, img)
            seen["regions"] = out[:, :, 0].copy()
            assert (out[:, :, 0] == out[:, :, 1]).all() and (out[:, :, 0] == out[:, :, 2]).all()
            return out

        def textline_contours(self, img):
            out = cls.textline_contours(self, img)
            seen["textline"] = out.copy()
            return out

        def extract_page(self):
            crop, coord = cls.extract_page(self)
            seen["page_coord"] = np.array(coord)
            return crop, coord

    t0 = time.time()
    det = Recording(png, tmp, "page", tmp)
    det.run()
    print(f"run(): {time.time() - t0:.1f} s")
    xml = open(os.path.join(tmp, "page.xml")).read().replace(png, "page.png")
    border, regions = semantic_fake.summarise_xml(xml)
    print("border", border, "regions", len(regions), "lines", [len(r[1]) for r in regions])
    print("slopes", det.slopes)
    open(os.path.join(HERE, "ref_semantic_run.xml"), "w").write(normalise(xml))
    order = np.lexsort(np.array(det.boxes).T[::-1])           # the reference collects in process-completion order
    np.savez_compressed(os.path.join(HERE, "ref_semantic_run.npz"),
                        page=np.array(PAGE), page_coord=seen["page_coord"], image_page_shape=seen["image_page_shape"],
                        regions_packed=np.packbits(seen["regions"] == 1), regions_other=np.array(int((seen["regions"] > 1).sum())),
                        textline_packed=np.packbits(seen["textline"] != 0),
                        boxes=np.array(det.boxes)[order], slopes=np.array(det.slopes, np.float64)[order],
                        n_regions=np.array(len(regions)), n_lines=np.array(sum(len(r[1]) for r in regions)))


if __name__ == "__main__":
    main()
