"""Import the UNMODIFIED reference module (/root/reference/qurator/sbb_textline_detector/main.py) in
this container, where tensorflow / keras / shapely / matplotlib / seaborn are absent, by putting inert
stub modules in sys.modules first.  Nothing of the reference is copied: the file is executed where it
lies.  Only the third-party calls that the hot path replaces are stubbed --

  * keras.models.load_model(path, compile=False) -> whatever ``MODEL_FACTORY[path]`` returns (a
    duck-typed model with ``.layers[-1].output_shape`` and ``.predict``, main.py:221,227-229,287)
  * tf.InteractiveSession() -> object with ``.close()`` (main.py:220,428)

everything else in the reference (numpy / cv2 / scipy arithmetic: do_prediction's tiling and stitch,
otsu_copy, resize_image, get_image_and_scales, the stage drivers) runs as written.  Used ONLY by
make_golden_from_reference.py (fixture generation; /root/reference does not exist on the GPU box).
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF_MAIN = "/root/reference/qurator/sbb_textline_detector/main.py"
MODEL_FACTORY = {}   # model path -> zero-arg callable returning the duck-typed model


class _Session:
    def close(self):
        pass


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_main():
    if not os.path.exists(REF_MAIN):
        raise FileNotFoundError(REF_MAIN)

    class _Logger:
        def setLevel(self, *_):
            pass

    class _GpuOpt:
        allow_growth = False

    class _Config:
        gpu_options = _GpuOpt()

    def load_model(path, compile=False):
        return MODEL_FACTORY[path]()

    if "tensorflow" not in sys.modules:
        _stub("tensorflow", get_logger=lambda: _Logger(), ConfigProto=_Config, InteractiveSession=_Session)
    if "keras" not in sys.modules:
        k = _stub("keras")
        k.models = _stub("keras.models", load_model=load_model, model_from_json=None)
        k.backend = _stub("keras.backend", clear_session=lambda: None)
    install_glue_stubs()
    spec = importlib.util.spec_from_file_location("_sbb_reference_main", REF_MAIN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install_glue_stubs():
    """Stand-ins for the host-glue dependencies this image lacks (matplotlib, seaborn: imported, never
    called on the run() path; shapely: Polygon.area / .exterior.coords only) and the cv2.cv2 alias."""
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "shapely"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    if not hasattr(sys.modules["shapely"], "geometry"):
        class Polygon:  # the reference reads .area and .exterior.coords (main.py:69-74, 85-90, 102-107)
            def __init__(self, pts):
                self._p = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
                ring = self._p if (self._p[0] == self._p[-1]).all() else np.vstack([self._p, self._p[:1]])
                self.exterior = types.SimpleNamespace(coords=[tuple(r) for r in ring.tolist()])  # closed ring

            @property
            def area(self):
                x, y = self._p[:, 0], self._p[:, 1]
                return 0.5 * abs(float(x @ np.roll(y, -1) - y @ np.roll(x, -1)))
        sys.modules["shapely"].geometry = _stub("shapely.geometry", Polygon=Polygon)
    # the two OpenCV API drifts since the reference's pinned 4.5.1 (cv2.cv2 alias; numpy ints as the point of
    # pointPolygonTest) -- same shim the product's compat.import_reference applies
    from sbb_textline_detection_b200.compat import modernise_cv2
    modernise_cv2()
    if "matplotlib.pyplot" in sys.modules and isinstance(sys.modules["matplotlib"], types.ModuleType):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
