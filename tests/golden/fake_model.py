"""Closed-form stand-in for ``model.predict`` shared by make_golden_from_reference.py and the tests:
class = f(pixel bytes, position inside the tile) -> one-hot float32.  Position dependence makes any
crop-offset / overwrite-order mistake visible, and nothing depends on float comparisons."""
import numpy as np


class _L:
    def __init__(self, shape):
        self.output_shape = shape


class FakeModel:
    def __init__(self, mh, mw, n_classes):
        self.layers = [_L((None, mh, mw, n_classes))]
        self.nc = n_classes

    def classes(self, x):
        x = np.asarray(x)
        n, h, w, _ = x.shape
        v = np.rint(x * 255.0).astype(np.int64)
        yy, xx = np.mgrid[0:h, 0:w]
        return (v[..., 0] * 7 + v[..., 1] * 13 + v[..., 2] * 31 + yy[None] * 3 + xx[None] * 5) % self.nc

    def predict(self, x):
        return np.eye(self.nc, dtype=np.float32)[self.classes(x)]


def seeded_page(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
