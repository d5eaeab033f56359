"""OCR-D processor ``ocrd-sbb-textline-detector`` on the B200 hot path (reference: ocrd_cli.py:29-141 the
processor, :144-214 the coordinate helpers; ocrd-tool.json the tool description).

The reference's processor saves the page image to a temporary PNG, runs ``textline_detector(...).run()`` on it,
parses the PAGE-XML that wrote, and merges Border / ReadingOrder / TextRegions / TextLines into the workspace's
PAGE file after mapping the coordinates back through the page transform and clipping every polygon to its parent.
This module is that processor with ONE difference: the detector class is the reference's own class bound to the
GPU methods (compat.bind_reference), its three models stay resident across the pages of a workspace.

``ocrd`` / ``ocrd_models`` / ``ocrd_utils`` / ``shapely`` are imported when the processor is built, not at module
import (they are optional dependencies of this package, as the OCR-D stack is of the reference's CLI):

    from sbb_textline_detection_b200 import ocrd_cli
    ocrd_cli.ocrd_sbb_textline_detector()         # console-script entry point (click), needs the OCR-D stack

The merge itself (``merge_segmentation``) only touches the generateDS accessors of the PAGE objects
(get_/set_Border, get_/set_TextRegion, ...), so it is tested without the OCR-D stack on duck-typed pages.
"""
from __future__ import annotations

import json
import os
import tempfile

TOOL = "ocrd-sbb-textline-detector"

# ocrd-tool.json of the reference, restated (same executable, steps, file groups and the one parameter)
OCRD_TOOL = {
    "version": "0.0.1",
    "git_url": "https://github.com/qurator-spk/sbb_textline_detection",
    "tools": {
        TOOL: {
            "executable": TOOL,
            "categories": ["Layout analysis"],
            "description": "Printspace, region and textline detection",
            "steps": ["layout/segmentation/region", "layout/segmentation/line"],
            "input_file_grp": ["OCR-D-IMG"],
            "output_file_grp": ["OCR-D-SBB-SEG-LINE"],
            "parameters": {
                "model": {"type": "string", "format": "uri", "content-type": "text/directory", "cacheable": True,
                          "description": "Path to directory containing models to be used "
                                         "(See https://qurator-data.de/sbb_textline_detector/)"},
            },
        }
    },
}


# ------------------------------------------------------------------------------------------------ geometry
def make_valid(polygon):
    """ocrd_cli.py:202-214: rotate the start point / simplify with growing tolerance until the ring is valid."""
    from shapely.geometry import Polygon
    for split in range(1, len(polygon.exterior.coords) - 1):
        if polygon.is_valid or polygon.simplify(polygon.area).is_valid:
            break
        polygon = Polygon(polygon.exterior.coords[-split:] + polygon.exterior.coords[:-split])
    for tolerance in range(1, int(polygon.area)):
        if polygon.is_valid:
            break
        polygon = polygon.simplify(tolerance)
    return polygon


def parent_polygon(parent):
    """The polygon a child is clipped to: a page's Border (or its full image), a region's Coords (ocrd_cli.py:158-167)."""
    if hasattr(parent, "get_imageHeight"):      # PageType
        if parent.get_Border():
            return _points(parent.get_Border().get_Coords().points)
        h, w = parent.get_imageHeight(), parent.get_imageWidth()
        return [[0, 0], [0, h], [w, h], [w, 0]]
    return _points(parent.get_Coords().points)


def _points(points: str):
    """ocrd_utils.polygon_from_points: 'x1,y1 x2,y2 ...' -> [[x1, y1], ...]"""
    return [[int(float(v)) for v in pair.split(",")] for pair in points.split()]


def _points_str(polygon) -> str:
    """ocrd_utils.points_from_polygon"""
    return " ".join("%i,%i" % (int(x), int(y)) for x, y in polygon)


def polygon_for_parent(polygon, parent):
    """Clip ``polygon`` to its parent (ocrd_cli.py:155-199): unchanged when inside, None when the intersection is
    empty, otherwise the (convex hull of the) intersection, rounded and made valid.  Needs shapely for anything
    but the contained case against an axis-parallel rectangle, which is decided exactly without it."""
    pp = parent_polygon(parent)
    try:
        from shapely.geometry import Polygon
        from shapely.ops import unary_union
    except ImportError:
        xs, ys = [p[0] for p in pp], [p[1] for p in pp]
        rect = len(pp) == 4 and len(set(xs)) == 2 and len(set(ys)) == 2
        if rect and all(min(xs) <= x <= max(xs) and min(ys) <= y <= max(ys) for x, y in polygon):
            return polygon
        raise RuntimeError("clipping a polygon to its parent needs shapely (an OCR-D dependency)")
    import numpy as np
    childp, parentp = Polygon(polygon), Polygon(pp)
    if childp.within(parentp):
        return polygon
    childp, parentp = make_valid(childp), make_valid(parentp)
    interp = childp.intersection(parentp)
    if interp.is_empty or interp.area == 0.0:
        return None
    if interp.geom_type == "GeometryCollection":
        interp = unary_union([geom for geom in interp.geoms if geom.area > 0])
    if interp.geom_type == "MultiPolygon":
        interp = interp.convex_hull
    if interp.minimum_clearance < 1.0:
        interp = make_valid(Polygon(np.round(interp.exterior.coords)))
    return interp.exterior.coords[:-1]


def adapt_coords(segment, parent, transform, coords_type=None, to_absolute=None):
    """ocrd_cli.py:144-154: segment polygon -> absolute coordinates (undoing the page transform) -> clipped to
    the parent; returns the segment with new Coords, or None when nothing is left."""
    polygon = _points(segment.get_Coords().get_points())
    if to_absolute is None:
        from ocrd_utils import coordinates_for_segment
        to_absolute = lambda poly: coordinates_for_segment(poly, None, transform)  # noqa: E731
    polygon_new = polygon_for_parent(to_absolute(polygon), parent)
    if polygon_new is None:
        return None
    if coords_type is None:
        from ocrd_models.ocrd_page_generateds import CoordsType as coords_type
    segment.set_Coords(coords_type(points=_points_str(polygon_new)))
    return segment


def merge_segmentation(page, tmp_page, transform, log=None, **adapt_kw):
    """Border, ReadingOrder, TextRegions and their TextLines of the detector's PAGE result into the workspace
    page (ocrd_cli.py:86-128), every polygon through ``adapt_coords``.  Returns (regions kept, lines kept)."""
    warn = log.warning if log is not None else (lambda *_a: None)
    if page.get_Border():
        warn("Removing existing page border")
    page.set_Border(None)
    text_border = adapt_coords(tmp_page.get_Border(), page, transform, **adapt_kw)
    if text_border is None:
        warn("new border would be empty, skipping")
    else:
        page.set_Border(text_border)
    if page.get_ReadingOrder():
        warn("Removing existing regions' reading order")
    page.set_ReadingOrder(tmp_page.get_ReadingOrder())
    if page.get_TextRegion():
        warn("Removing existing text regions")
    regions, n_lines = [], 0
    for text_region in tmp_page.get_TextRegion():
        text_region = adapt_coords(text_region, page, transform, **adapt_kw)
        if text_region is None:
            warn("new text region polygon would be empty, skipping")
            continue
        regions.append(text_region)
        lines = []
        for text_line in text_region.get_TextLine():
            text_line = adapt_coords(text_line, text_region, transform, **adapt_kw)
            if text_line is None:
                warn("new text line polygon would be empty, skipping")
                continue
            lines.append(text_line)
        text_region.set_TextLine(lines)
        n_lines += len(lines)
    page.set_TextRegion(regions)
    return len(regions), n_lines


# ------------------------------------------------------------------------------------------------ processor
def detector_class(reference_main: str | None = None, **bind_kw):
    """The reference's textline_detector bound to the GPU hot path (compat.bind_reference); the installed
    ``qurator.sbb_textline_detector`` package is used unless ``reference_main`` names a main.py."""
    from . import cli, compat
    return compat.bind_reference(cli._reference_module(reference_main), **bind_kw)


def make_processor(reference_main: str | None = None, **bind_kw):
    """-> the ocrd.Processor subclass (built lazily: the OCR-D stack is an optional dependency)."""
    import ocrd_models.ocrd_page
    from ocrd import Processor
    from ocrd_modelfactory import page_from_file
    from ocrd_utils import assert_file_grp_cardinality, getLogger, make_file_id

    textline_detector = detector_class(reference_main or os.environ.get("SBB_REFERENCE_MAIN"), **bind_kw)

    class OcrdSbbTextlineDetectorRecognize(Processor):
        def __init__(self, *args, **kwargs):
            kwargs["ocrd_tool"] = OCRD_TOOL["tools"][TOOL]
            kwargs["version"] = OCRD_TOOL["version"]
            super().__init__(*args, **kwargs)

        def process(self):
            log = getLogger("processor.OcrdSbbTextlineDetectorRecognize")
            assert_file_grp_cardinality(self.input_file_grp, 1)
            assert_file_grp_cardinality(self.output_file_grp, 1)
            model = self.resolve_resource(self.parameter["model"])
            for n, input_file in enumerate(self.input_files):
                page_id = input_file.pageId or input_file.ID
                log.info("INPUT FILE %i / %s", n, input_file)
                file_id = make_file_id(input_file, self.output_file_grp)
                os.makedirs(self.output_file_grp, exist_ok=True)
                pcgts = page_from_file(self.workspace.download_file(input_file))
                page = pcgts.get_Page()
                page_image, page_coords, _info = self.workspace.image_from_page(
                    page, page_id, feature_filter="cropped,binarized,grayscale_normalized")
                with tempfile.TemporaryDirectory() as tmp_dirname:
                    image_file = tempfile.mkstemp(dir=tmp_dirname, suffix=".png")[1]
                    page_image.save(image_file)
                    x = textline_detector(image_file, tmp_dirname, file_id, model)   # models are cached per process
                    x.run()
                    tmp_pcgts = ocrd_models.ocrd_page.parse(os.path.join(tmp_dirname, file_id) + ".xml", silence=True)
                    tmp_page = tmp_pcgts.get_Page()
                pcgts.set_pcGtsId(file_id)
                merge_segmentation(page, tmp_page, page_coords, log)
                self.add_metadata(pcgts)
                self.workspace.add_file(ID=file_id, file_grp=self.output_file_grp, pageId=page_id,
                                        mimetype="application/vnd.prima.page+xml",
                                        local_filename=os.path.join(self.output_file_grp, file_id) + ".xml",
                                        content=ocrd_models.ocrd_page.to_xml(pcgts))

    return OcrdSbbTextlineDetectorRecognize


def ocrd_sbb_textline_detector(*args, **kwargs):
    """Console-script entry point (ocrd_cli.py:29-32): the standard OCR-D command line around the processor."""
    import click
    from ocrd.decorators import ocrd_cli_options, ocrd_cli_wrap_processor

    @click.command()
    @ocrd_cli_options
    def _main(*a, **k):
        return ocrd_cli_wrap_processor(make_processor(), *a, **k)
    return _main(*args, **kwargs)


def dump_tool_json() -> str:
    """The ocrd-tool.json content (``ocrd-sbb-textline-detector --dump-json`` prints this in the OCR-D CLI)."""
    return json.dumps(OCRD_TOOL, indent=2)


if __name__ == "__main__":
    ocrd_sbb_textline_detector()
