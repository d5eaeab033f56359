"""Golden PAGE-XML files written by the UNMODIFIED reference's write_into_page_xml (main.py:1908-2053)
for hand-made region / line polygons (both shapes the reference handles: cv2 contours [N,1,2] and
rotated-rectangle points [N,2]) and for the Border-only case (main.py:2125-2129, 2152-2156).

    python tests/golden/make_golden_xml.py  ->  tests/golden/ref_page_full.xml, ref_page_border_only.xml
"""
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def scene():
    """Deterministic inputs shared with tests/test_page_xml.py."""
    rng = np.random.default_rng(3)
    page_coord = [37, 2700, 21, 1950]
    cont_page = [np.array([[21, 37], [1950, 37], [1950, 2700], [21, 2700]])]
    regions, lines, boxes = [], [], []
    for r in range(4):
        x, y, w, h = 100 + 30 * r, 150 + 600 * r, 1500 - 100 * r, 480
        n = int(rng.integers(5, 12))
        regions.append(rng.integers(0, 1900, (n, 1, 2)).astype(np.int32))     # like cv2.findContours
        boxes.append([y, y + h, x, x + w])                                    # crop_image_inside_box order
        ls = []
        for j in range(int(rng.integers(0, 5)) if r == 1 else int(rng.integers(2, 5))):
            if j % 2:
                ls.append(rng.integers(0, 1400, (4, 2)).astype(np.int64))     # rotated-rectangle points
            else:
                ls.append(rng.integers(0, 1400, (int(rng.integers(4, 9)), 1, 2)).astype(np.int32))
        lines.append(ls)
    order = [2, 0, 3, 1]
    ids = ["r0", "r1", "r2", "r3"]
    return dict(page_coord=page_coord, cont_page=cont_page, regions=regions, lines=lines, boxes=boxes, order=order,
                ids=ids, height_org=1403, width_org=1017, scale_y=2800 / 1403.0, scale_x=2028 / 1017.0,
                image_dir="/data/scans/page_0001.png")


def normalise(xml: str) -> str:
    return re.sub(r"<(Created|LastChange)>[^<]*</", r"<\1>T</", xml)


def main():
    ref = ref_import.load_reference_main()
    tmp = tempfile.mkdtemp()
    s = scene()
    for name, with_regions in (("ref_page_full", True), ("ref_page_border_only", False)):
        det = ref.textline_detector(s["image_dir"], tmp, name, tmp)
        det.height_org, det.width_org, det.scale_x, det.scale_y = s["height_org"], s["width_org"], s["scale_x"], s["scale_y"]
        det.cont_page = s["cont_page"]
        det.all_found_texline_polygons, det.all_box_coord = s["lines"], s["boxes"]
        if with_regions:
            det.write_into_page_xml(s["regions"], s["page_coord"], tmp, s["order"], s["ids"])
        else:
            det.write_into_page_xml([], s["page_coord"], tmp, None, None)
        xml = open(os.path.join(tmp, name + ".xml")).read()
        open(os.path.join(HERE, name + ".xml"), "w").write(normalise(xml))
        print(name, len(xml), "bytes")


if __name__ == "__main__":
    main()
