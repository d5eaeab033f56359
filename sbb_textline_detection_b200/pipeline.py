"""Page dispatcher for the three-model pipeline of ``run()`` (main.py:2056-2107): border model -> crop ->
region model (Otsu) -> textline model, then the host glue.

The reference processes ONE page per process invocation, stage after stage, re-creating each model for every
stage (main.py:386, 442, 492).  Served from a GPU that leaves the device idle whenever the host is busy: the
border stage ends in a host contour pass (cv2.findContours + the crop decision, main.py:398-426) before the two
tiled models can start, and everything after ``textline_contours`` (contours, deskew, line separation, XML) is
host work.  The dispatcher keeps the three models resident (``detector._MODEL_CACHE``) and runs several pages
in flight, one worker thread per page: while one page is in its host sections -- OpenCV and the ctypes calls
release the GIL -- the other workers' GPU stages fill the device.  Calls on one model handle are serialised by its
lock on the host and ordered on the device by the library's event chain, so workers share the handles safely;
every page has its own crop geometry, which the handle's geometry cache absorbs.

    with PageDispatcher(dir_models, workers=4) as d:
        for result in d.map(image_paths):            # results in input order
            page_coord, regions, textline_mask = result

``stage`` picks what a worker runs per page: "segmentation" (the three model stages -> label maps, BASELINE
config 3) or "run" (the bound reference's full ``run()`` -> PAGE-XML on disk, configs 4/5; needs the reference
module, see compat.bind_reference).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import threading

from . import detector as D


class PageDispatcher:
    def __init__(self, dir_models: str, dir_out: str | None = None, *, workers: int = 4, device: int = 0,
                 tile: int | None = None, precision: str = "fp16x3", max_batch: int = 48, stage: str = "segmentation",
                 detector_cls=None):
        assert stage in ("segmentation", "run")
        self.dir_models, self.dir_out = dir_models, dir_out or "."
        self.kw = dict(device=device, tile=tile, precision=precision, max_batch=max_batch)
        self.stage = stage
        self.cls = detector_cls          # stage "run": the class compat.bind_reference returned
        if stage == "run" and detector_cls is None:
            raise ValueError('stage="run" needs detector_cls=compat.bind_reference(reference_module, ...)')
        self.workers = max(1, int(workers))
        self._pool = cf.ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="sbb-page")
        self._first = threading.Lock()
        self._warm = False
        self._tls = threading.local()

    # ------------------------------------------------------------------ one page
    def _one(self, item):
        if self.stage == "run":
            det = self.cls(item, self.dir_out, None, self.dir_models)
            det.run()
            return os.path.join(self.dir_out, det.f_name + ".xml")
        if isinstance(item, str):
            det = D.textline_detector(item, self.dir_out, None, self.dir_models, **self.kw)
            det.get_image_and_scales()
        else:                            # an already scaled page image (uint8 BGR array): benchmarks, services
            det = D.textline_detector("<array>", self.dir_out, "page", self.dir_models, **self.kw)
            det.image = item
        image_page, page_coord = det.extract_page()
        regions = det.extract_text_regions(image_page)
        textline = det.textline_contours(image_page)
        return page_coord, regions, textline

    def _on_own_stream(self, item):
        # one CUDA stream per worker: a worker's blocking copies (.cpu()) then wait for ITS page only, not for
        # whatever the other workers have queued on a shared stream
        import torch
        if not torch.cuda.is_available():   # stage="run" with a detector class that brings its own (test) models
            return self._one(item)
        st = getattr(self._tls, "stream", None)
        if st is None:
            st = self._tls.stream = torch.cuda.Stream(torch.device("cuda", self.kw["device"]))
        with torch.cuda.stream(st):
            return self._one(item)

    def _guarded(self, item):
        # the first page loads (and caches) the three models: let exactly one worker do that
        if not self._warm:
            with self._first:
                if not self._warm:
                    out = self._on_own_stream(item)
                    self._warm = True
                    return out
        return self._on_own_stream(item)

    # ------------------------------------------------------------------ many pages
    def map(self, items):
        """Results in input order; at most ``workers`` pages in flight."""
        futures = [self._pool.submit(self._guarded, it) for it in items]
        for f in futures:
            yield f.result()

    def close(self):
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
