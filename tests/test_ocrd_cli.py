"""OCR-D processor restatement (sbb_textline_detection_b200/ocrd_cli.py; reference ocrd_cli.py:29-214).  The
OCR-D stack is not installed here, so the merge of the detector's PAGE result into the workspace page is tested
on duck-typed PAGE objects: only the generateDS accessors the reference touches exist on them."""
import json
import os

import pytest

from sbb_textline_detection_b200 import ocrd_cli

REF_TOOL = "/root/reference/qurator/sbb_textline_detector/ocrd-tool.json"


class Coords:
    def __init__(self, points):
        self.points = points

    def get_points(self):
        return self.points


class Seg:
    def __init__(self, points, lines=()):
        self._c, self._lines = Coords(points), list(lines)

    def get_Coords(self):
        return self._c

    def set_Coords(self, c):
        self._c = c

    def get_TextLine(self):
        return self._lines

    def set_TextLine(self, lines):
        self._lines = list(lines)


class Page:
    def __init__(self, h, w, border=None, regions=(), order=None):
        self.h, self.w, self.border, self.regions, self.order = h, w, border, list(regions), order

    def get_imageHeight(self):
        return self.h

    def get_imageWidth(self):
        return self.w

    def get_Border(self):
        return self.border

    def set_Border(self, b):
        self.border = b

    def get_ReadingOrder(self):
        return self.order

    def set_ReadingOrder(self, o):
        self.order = o

    def get_TextRegion(self):
        return self.regions

    def set_TextRegion(self, r):
        self.regions = list(r)


def test_tool_description_matches_the_reference():
    tool = ocrd_cli.OCRD_TOOL["tools"][ocrd_cli.TOOL]
    assert tool["executable"] == "ocrd-sbb-textline-detector" and set(tool["parameters"]) == {"model"}
    assert json.loads(ocrd_cli.dump_tool_json()) == ocrd_cli.OCRD_TOOL
    if os.path.exists(REF_TOOL):
        assert json.load(open(REF_TOOL)) == ocrd_cli.OCRD_TOOL


def test_merge_translates_and_replaces_border_order_regions_lines():
    shift = lambda poly: [[x + 100, y + 50] for x, y in poly]     # page transform: the image was cropped at (100, 50)
    lines = [Seg("10,10 200,10 200,30 10,30"), Seg("10,40 200,40 200,60 10,60")]
    tmp = Page(1000, 800, border=Seg("0,0 700,0 700,900 0,900"),
               regions=[Seg("5,5 300,5 300,100 5,100", lines), Seg("5,200 300,200 300,300 5,300")], order="RO")
    page = Page(1200, 1000, border=Seg("1,1 2,1 2,2"), regions=[Seg("0,0 1,0 1,1")], order="old")
    warnings = []
    log = type("L", (), {"warning": lambda self, *a: warnings.append(a[0])})()
    n_regions, n_lines = ocrd_cli.merge_segmentation(page, tmp, None, log, coords_type=Coords, to_absolute=shift)
    assert (n_regions, n_lines) == (2, 2) and len(warnings) == 3            # border, reading order, regions replaced
    assert page.get_Border().get_Coords().points == "100,50 800,50 800,950 100,950"
    assert page.get_ReadingOrder() == "RO"
    assert page.get_TextRegion()[0].get_Coords().points == "105,55 400,55 400,150 105,150"
    assert [l.get_Coords().points for l in page.get_TextRegion()[0].get_TextLine()] == \
        ["110,60 300,60 300,80 110,80", "110,90 300,90 300,110 110,110"]
    assert page.get_TextRegion()[1].get_TextLine() == []


def test_clipping_without_shapely_is_refused_not_approximated():
    try:
        import shapely.ops  # noqa: F401   (other tests install a two-attribute shapely stand-in without .ops)
        pytest.skip("shapely installed: the reference's clipping path runs")
    except ImportError:
        pass
    page = Page(100, 100)
    inside = [[1, 1], [50, 1], [50, 50]]
    assert ocrd_cli.polygon_for_parent(inside, page) == inside
    with pytest.raises(RuntimeError, match="shapely"):
        ocrd_cli.polygon_for_parent([[1, 1], [150, 1], [50, 50]], page)


def test_processor_needs_the_ocrd_stack():
    try:
        import ocrd  # noqa: F401
        pytest.skip("OCR-D stack installed")
    except ImportError:
        pass
    with pytest.raises(ImportError):
        ocrd_cli.make_processor()
