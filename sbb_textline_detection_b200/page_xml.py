"""PAGE-XML output of the reference (``textline_detector.write_into_page_xml``, main.py:1908-2053),
restated: same element tree, attribute order, id scheme (r<k> / l<k>),
reading-order group, and the reference's coordinate rule -- contour coordinates are shifted by the page
crop (and, for text lines, by the region box), divided by the resize scale and TRUNCATED with int()
(main.py:1951-2043).  Pinned byte-for-byte (timestamps aside) against files written by the unmodified
reference (tests/golden/make_golden_xml.py)."""
from __future__ import annotations

import datetime
import os
import xml.etree.ElementTree as ET

import numpy as np

NS = "http://schema.primaresearch.org/PAGE/gts/pagecontent/2019-07-15"
READING_ORDER_GROUP_ID = "ro357564684568544579089"  # main.py:1969 (a constant in the reference)


def points_attr(poly, off_x, off_y, scale_x, scale_y) -> str:
    """'x,y x,y ...' for a polygon given either as [N,2] points or as a cv2 contour [N,1,2]."""
    parts = []
    for p in poly:
        if len(p) != 2:
            p = p[0]
        parts.append(f"{int((p[0] + off_x) / scale_x)},{int((p[1] + off_y) / scale_y)}")
    return " ".join(parts)


def build_tree(image_filename, height_org, width_org, scale_x, scale_y, cont_page, page_coord, regions,
               textlines_per_region=None, box_coords=None, order_of_texts=None, id_of_texts=None, now=None):
    now = now or datetime.datetime.now().isoformat()
    root = ET.Element("PcGts")
    root.set("xmlns", NS)
    root.set("xmlns:xsi", "http://www.w3.org/2001/XMLSchema-instance")
    root.set("xsi:schemaLocation", NS)
    meta = ET.SubElement(root, "Metadata")
    ET.SubElement(meta, "Creator").text = "SBB_QURATOR"
    ET.SubElement(meta, "Created").text = now
    ET.SubElement(meta, "LastChange").text = now
    page = ET.SubElement(root, "Page")
    page.set("imageFilename", image_filename)
    page.set("imageHeight", str(height_org))
    page.set("imageWidth", str(width_org))
    page.set("type", "content")
    page.set("readingDirection", "left-to-right")
    page.set("textLineOrder", "top-to-bottom")
    border = ET.SubElement(ET.SubElement(page, "Border"), "Coords")
    border.set("points", points_attr(cont_page[0], 0, 0, scale_x, scale_y))
    if len(regions) > 0:
        group = ET.SubElement(ET.SubElement(page, "ReadingOrder"), "OrderedGroup")
        group.set("id", READING_ORDER_GROUP_ID)
        for k in np.argsort(order_of_texts):
            ref = ET.SubElement(group, "RegionRefIndexed")
            ref.set("index", str(order_of_texts[k]))
            ref.set("regionRef", id_of_texts[k])
        line_id = 0
        for k, region in enumerate(regions):
            tr = ET.SubElement(page, "TextRegion")
            tr.set("id", "r" + str(k))
            tr.set("type", "paragraph")
            ET.SubElement(tr, "Coords").set("points", points_attr(region, page_coord[2], page_coord[0], scale_x, scale_y))
            for line in textlines_per_region[k]:
                tl = ET.SubElement(tr, "TextLine")
                tl.set("id", "l" + str(line_id))
                line_id += 1
                ET.SubElement(tl, "Coords").set("points", points_attr(line, page_coord[2] + box_coords[k][2],
                                                                      page_coord[0] + box_coords[k][0], scale_x, scale_y))
    return ET.ElementTree(root)


def write_page_xml(path, *args, **kwargs):
    build_tree(*args, **kwargs).write(path)
    return path


def write_into_page_xml(det, contours, page_coord, dir_of_image, order_of_texts, id_of_texts):
    """Same call as the reference method (``dir_of_image`` is unused there too); ``det``
    carries image_dir, height_org, width_org, scale_x/y, cont_page, all_found_texline_polygons,
    all_box_coord, dir_out, f_name like the reference instance does."""
    has = len(contours) > 0
    return write_page_xml(os.path.join(det.dir_out, det.f_name) + ".xml", det.image_dir, det.height_org, det.width_org,
                          det.scale_x, det.scale_y, det.cont_page, page_coord, contours,
                          det.all_found_texline_polygons if has else None, det.all_box_coord if has else None,
                          order_of_texts, id_of_texts)
