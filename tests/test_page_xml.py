"""PAGE-XML writer byte-for-byte against files written by the unmodified
reference (tests/golden/make_golden_xml.py), timestamps normalised."""
import os
import sys
import types

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden_xml import normalise, scene  # noqa: E402  (only its pure helpers; no reference import)

from sbb_textline_detection_b200 import page_xml


def _det(tmp_path, name, s):
    d = types.SimpleNamespace(image_dir=s["image_dir"], dir_out=str(tmp_path), f_name=name, height_org=s["height_org"],
                              width_org=s["width_org"], scale_x=s["scale_x"], scale_y=s["scale_y"], cont_page=s["cont_page"],
                              all_found_texline_polygons=s["lines"], all_box_coord=s["boxes"])
    return d


def test_full_page_xml_equals_reference(tmp_path):
    s = scene()
    p = page_xml.write_into_page_xml(_det(tmp_path, "full", s), s["regions"], s["page_coord"], str(tmp_path), s["order"], s["ids"])
    assert normalise(open(p).read()) == open(os.path.join(GOLDEN, "ref_page_full.xml")).read()


def test_border_only_xml_equals_reference(tmp_path):
    s = scene()
    p = page_xml.write_into_page_xml(_det(tmp_path, "border", s), [], s["page_coord"], str(tmp_path), None, None)
    assert normalise(open(p).read()) == open(os.path.join(GOLDEN, "ref_page_border_only.xml")).read()


def test_coordinates_truncate_like_int():
    assert page_xml.points_attr([[5, 9], [[7, 11]]], 1, 2, 2.0, 3.0) == "3,3 4,4"
    assert page_xml.points_attr([[-5, 0]], 0, 0, 2.0, 2.0) == "-2,0"   # int() truncates toward zero
