"""Ink-density stand-ins for the three models so that the reference's host glue sees plausible label
maps (text blocks, text lines) on a synthetic document page; shared by make_golden_pipeline_xml.py and
tests/test_compat_binding.py."""
import cv2
import numpy as np


class _L:
    def __init__(self, shape):
        self.output_shape = shape


class SemanticFake:
    def __init__(self, kind, tile=448, n_classes=2):
        self.kind, self.nc = kind, n_classes
        self.layers = [_L((None, tile, tile, n_classes))]

    def predict(self, x):
        x = np.asarray(x)[0]
        ink = (x.mean(axis=2) < 0.5).astype(np.uint8)
        if self.kind == "page":
            cls = np.ones(ink.shape, np.int64)
        elif self.kind == "region":
            cls = (cv2.dilate(ink, np.ones((61, 91), np.uint8)) > 0).astype(np.int64)
        else:
            cls = (cv2.dilate(ink, np.ones((3, 31), np.uint8)) > 0).astype(np.int64)
        return np.eye(self.nc, dtype=np.float32)[cls][None]


def loader(path):
    if "page" in path:
        return SemanticFake("page")
    if "struktur" in path:
        return SemanticFake("region", n_classes=4)
    return SemanticFake("textline")


def summarise_xml(xml: str):
    """Order-independent content of a PAGE-XML file: the reference assigns region ids in process-completion
    order (main.py:1782-1794), so compare sets of (region polygon, its line polygons)."""
    import re
    border = re.search(r"<Border><Coords points=\"([^\"]*)\"", xml).group(1)
    regions = []
    for m in re.finditer(r"<TextRegion [^>]*><Coords points=\"([^\"]*)\" />(.*?)</TextRegion>", xml, re.S):
        lines = tuple(sorted(re.findall(r"<TextLine [^>]*><Coords points=\"([^\"]*)\"", m.group(2))))
        regions.append((m.group(1), lines))
    return border, sorted(regions)
