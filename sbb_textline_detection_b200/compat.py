"""Bind the reference's own host glue to the B200 hot path (executable form).

``bind_reference(ref_module)`` takes the imported reference module
(``qurator.sbb_textline_detector.main``) and returns a subclass of ITS ``textline_detector`` in which only
the hot-path methods are replaced:

    start_new_session_and_model   main.py:216-223   -> GPU-resident SbbModel (cached per process)
    do_prediction                 main.py:225-380   -> fused tiled forward + argmax + stitch on the GPU
    return_deskew_slope           main.py:1601-1718 -> one-launch rotation profiles on the GPU (identical angle)
    get_slopes_and_deskew         main.py:1760-1799 -> the reference's per-chunk worker run IN-PROCESS (its fork
                                                       fan-out cannot use CUDA in the children)

Everything else -- contours, line separation, reading order, PAGE-XML (main.py:456-2053) and ``run()`` --
is the reference's code, untouched, so the CLI / OCR-D wrapper keep working on top of the returned class.
On a modern stack the reference module no longer imports (tensorflow 1.15 / keras 2.3 pins);
``import_reference(path)`` imports it with inert tensorflow / keras stand-ins, since after binding
neither is called any more.
"""
from __future__ import annotations

import importlib.util
import sys
import types

from . import detector as D


def import_reference(main_py: str, name: str = "_sbb_reference_main"):
    """Import the reference's main.py from ``main_py`` when tensorflow / keras are not installed: the
    two packages are only used by start_new_session_and_model / K.clear_session, which the bound class
    replaces / which become no-ops.  Real installs are left alone."""
    def stub(modname, **attrs):
        m = types.ModuleType(modname)
        m.__dict__.update(attrs)
        sys.modules[modname] = m
        return m

    class _Logger:
        def setLevel(self, *_):
            pass

    def _unavailable(*_a, **_k):
        raise RuntimeError("tensorflow/keras are not installed; use bind_reference() so the GPU path serves the models")

    try:
        import tensorflow  # noqa: F401
    except Exception:
        stub("tensorflow", get_logger=lambda: _Logger(), ConfigProto=_unavailable, InteractiveSession=_unavailable)
    try:
        import keras  # noqa: F401
    except Exception:
        k = stub("keras")
        k.models = stub("keras.models", load_model=_unavailable, model_from_json=_unavailable)
        k.backend = stub("keras.backend", clear_session=lambda: None)
    modernise_cv2()
    spec = importlib.util.spec_from_file_location(name, main_py)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def modernise_cv2():
    """The reference pins opencv-python-headless==4.5.1.48 (requirements.txt); two of its call sites do not
    survive a current OpenCV unchanged:
      * main.py:471 spells ``cv2.cv2.RETR_TREE`` (the pre-4.6 module layout) -> alias
      * ``seperate_lines*`` pass numpy integers as the point of ``cv2.pointPolygonTest`` (main.py:780 and
        siblings); OpenCV >= 4.6 only accepts Python numbers there, the resulting cv2.error is swallowed by
        the bare ``except`` of textline_contours_postprocessing (main.py:1521) and every region silently ends
        up with ZERO text lines -> coerce the point, nothing else changes."""
    import cv2
    if not hasattr(cv2, "cv2"):
        cv2.cv2 = cv2
    if not getattr(cv2.pointPolygonTest, "_sbb_coerces", False):
        orig = cv2.pointPolygonTest

        def pointPolygonTest(contour, pt, measureDist):  # noqa: N802  (OpenCV's name)
            return orig(contour, (float(pt[0]), float(pt[1])), measureDist)
        pointPolygonTest._sbb_coerces = True
        cv2.pointPolygonTest = pointPolygonTest


def bind_reference(ref_module, *, device: int = 0, tile: int | None = None, precision: str = "fp16x3", max_batch: int = 48,
                   cache_models: bool = True, gpu_deskew: bool = True, model_loader=None):
    """-> subclass of ``ref_module.textline_detector`` with the hot path on the GPU.
    ``model_loader(path)`` (optional) overrides how a model path becomes a model object (tests plug
    duck-typed models in here; default: detector.load_model_file -> SbbModel)."""
    base = ref_module.textline_detector
    ours = D.textline_detector

    class textline_detector(base):  # noqa: N801  (the reference's class name)
        def __init__(self, image_dir, dir_out, f_name, dir_models):
            super().__init__(image_dir, dir_out, f_name, dir_models)
            self._device, self._tile, self._precision = device, tile, precision
            self._cache, self._max_batch = cache_models, max_batch

        def start_new_session_and_model(self, model_dir):
            if model_loader is not None:
                return model_loader(model_dir), D._NullSession()
            return ours.start_new_session_and_model(self, model_dir)

        def do_prediction(self, patches, img, model):
            from . import _lib
            try:
                return ours.do_prediction(self, patches, img, model)
            except _lib.SbbError as e:
                # run() wraps the region / textline stages in bare ``except:`` blocks (main.py:2069-2157) that turn
                # ANY failure into an empty result: remember a broken hot path so that run() below re-raises it
                self._hot_path_error = e
                raise

        def run(self):
            """The reference's ``run()`` (main.py:2056-2157); its bare ``except:`` blocks write a border-only (or
            region-less) PAGE-XML for ANY failure, so a hot-path error is re-raised once it has returned."""
            self._hot_path_error = None
            base.run(self)
            if self._hot_path_error is not None:
                raise self._hot_path_error

        if gpu_deskew:
            def return_deskew_slope(self, img_patch, sigma_des):
                from . import _lib, deskew
                if not deskew.is_two_valued(img_patch):
                    # several non-zero values: the reference interpolates them before its != 0 test
                    # (main.py:1631-1632), which the binarise-first GPU search does not reproduce -> its own code
                    return base.return_deskew_slope(self, img_patch, sigma_des)
                try:
                    return deskew.return_deskew_slope(img_patch, sigma_des, device=self._device)
                except _lib.SbbError as e:
                    # the caller (main.py:1734-1739) turns EVERY exception into slope 0: remember a broken hot
                    # path so that get_slopes_and_deskew / run() fail loudly instead
                    self._hot_path_error = e
                    raise

            def get_slopes_and_deskew(self, contours, textline_mask_tot):
                """main.py:1760-1799 without the ``multiprocessing.Process`` fan-out: the reference forks
                cpu_count() workers AFTER the parent created its CUDA context, and CUDA cannot be used in such a
                child -- every GPU call in it fails, main.py:1736-1739 swallows that into slope 0.  The search
                is one kernel launch per region here, so the reference's own per-chunk worker
                (``do_work_of_slopes``, untouched) runs in this process over all boxes as ONE chunk; the
                reference collects its chunks in completion order, of which this is the in-order case."""
                class _Collect:
                    def __init__(self):
                        self.items = []

                    def put(self, item):
                        self.items.append(item)

                q = _Collect()
                self.do_work_of_slopes(q, self.boxes, textline_mask_tot, contours)
                if getattr(self, "_hot_path_error", None) is not None:
                    raise self._hot_path_error
                slopes, polys, boxes, regions = q.items[0]
                self.slopes = list(slopes)
                self.all_found_texline_polygons = list(polys)
                self.boxes = list(boxes)
                return list(regions)

    textline_detector.__doc__ = "reference textline_detector bound to the sbb_textline_detection_b200 hot path"
    return textline_detector
