"""Drop-in for the segmentation hot path of ``qurator.sbb_textline_detector.main.textline_detector``.

Same constructor and method names as the reference class for everything on the hot path
(SURVEY.md section 8a):

    start_new_session_and_model   main.py:216-223
    do_prediction                 main.py:225-380
    resize_image / otsu_copy      main.py:112-113 / 178-194
    get_image_and_scales          main.py:196-214
    extract_page                  main.py:384-437   (model inference + border crop)
    extract_text_regions          main.py:439-454
    textline_contours             main.py:490-503

The contour / deskew / reading-order / PAGE-XML glue after these calls (main.py:456-481, 516-2053) is
host code that stays with the reference (SURVEY.md section 8f marks it "next"); INTEGRATION.md shows
how the reference class binds to this module by overriding exactly the two methods
``start_new_session_and_model`` and ``do_prediction``.

Models are GPU-resident and cached per process (the reference reloads each .h5 for every page and
stage, main.py:386,442,492); ``session.close()`` is therefore a no-op unless ``cache_models=False``.
"""
from __future__ import annotations

import os

import cv2
import numpy as np

from . import weights as W
from .model import SbbModel, SbbSession

# file names the reference hard-codes (main.py:58-60) -> (synthetic seed, n_classes, stats file)
_SYNTHETIC = {
    "model_page_mixed_best.h5": (1236, 2, "page"),
    "model_strukturerkennung.h5": (1235, 4, "region"),
    "model_textline_new.h5": (1234, 2, "textline"),
}
_MODEL_CACHE: dict = {}
_MODEL_CACHE_LOCK = __import__("threading").Lock()   # detectors on several threads (pipeline.PageDispatcher) share the cache


class _NullSession:
    def close(self):
        pass


def synthetic_weights(kind: str):
    """Seeded random-init weights with calibrated BatchNorm statistics for 'textline' | 'region' |
    'page' (there are no .h5 files or network access in this environment)."""
    seed, nc, name = {v[2]: v for v in _SYNTHETIC.values()}[kind]
    stats = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", f"bn_stats_{name}.npz"))
    return W.apply_bn_stats(W.random_init(seed, nc), stats), nc


def load_model_file(path: str, tile: int | None = None, device: int = 0, precision: str = "fp16x3", max_batch: int = 48):
    """Resolve a model path the way the reference's ``load_model(model_dir, compile=False)`` call
    site expects (main.py:221).  Order: an ``.sbbw`` blob (weights.pack_blob) next to the ``.h5``;
    the Keras ``.h5`` itself (keras_h5.read_keras_h5 -- bundled HDF5 reader, BatchNorm folded on the
    host); with neither present and SBB_SYNTHETIC_MODELS=1 the seeded synthetic weights of the same role.

    The model's input size decides the tile grid, the margins and the ``patches=False`` resize
    (main.py:227-233, 371), so it is never guessed: it comes from the blob header / the file's
    ``model_config``; ``tile`` is only used when the file records none (and must agree when it does);
    with neither this raises instead of falling back to 448."""
    base = os.path.basename(path)
    blob_path = os.path.splitext(path)[0] + ".sbbw"

    def resolve(recorded, what):
        if recorded is not None:
            if tile is not None and tuple(recorded) != (tile, tile):
                raise ValueError(f"{what} was built for {recorded[0]}x{recorded[1]} inputs, tile={tile} requested")
            return tuple(recorded)
        if tile is None:
            raise ValueError(f"{what} does not record the model's input size: pass tile= explicitly "
                             "(a wrong tile size silently changes the tile grid and margins)")
        return tile, tile

    if os.path.exists(blob_path):
        blob = open(blob_path, "rb").read()
        nc, _ = W.unpack_blob(blob)
        th, tw = resolve(W.blob_tile(blob), blob_path)
        return SbbModel(blob, th, tw, nc, device=device, precision=precision, max_batch=max_batch)
    if os.path.exists(path):
        from .keras_h5 import read_keras_h5
        w, nc, tile_hw = read_keras_h5(path)
        th, tw = resolve(tile_hw, path)
        return SbbModel(w, th, tw, nc, device=device, precision=precision, max_batch=max_batch)
    if base in _SYNTHETIC and os.environ.get("SBB_SYNTHETIC_MODELS") in ("1", "semantic"):
        if os.environ["SBB_SYNTHETIC_MODELS"] == "semantic":   # document-like stand-ins (semantic.py)
            from .semantic import semantic_weights
            w, nc = semantic_weights(_SYNTHETIC[base][2])
        else:
            w, nc = synthetic_weights(_SYNTHETIC[base][2])
        t = tile or 448
        return SbbModel(w, t, t, nc, device=device, precision=precision, max_batch=max_batch)
    raise FileNotFoundError(
        f"neither {path} nor {blob_path} found (SBB_SYNTHETIC_MODELS=1: seeded random-init weights, "
        "=semantic: document-like synthetic weights)")


class textline_detector:
    def __init__(self, image_dir, dir_out, f_name, dir_models, *, device: int = 0, tile: int | None = None,
                 precision: str = "fp16x3", cache_models: bool = True, max_batch: int = 48):
        self.image_dir = image_dir  # path of the page image file (main.py:46-47 keeps the historical name)
        self.dir_out = dir_out
        self.f_name = f_name
        if self.f_name is None:
            try:
                self.f_name = image_dir.split('/')[len(image_dir.split('/')) - 1]
                self.f_name = self.f_name.split('.')[0]
            except Exception:
                self.f_name = self.f_name.split('.')[0]
        self.dir_models = dir_models
        self.kernel = np.ones((5, 5), np.uint8)
        self.model_page_dir = dir_models + '/model_page_mixed_best.h5'
        self.model_region_dir = dir_models + '/model_strukturerkennung.h5'
        self.model_textline_dir = dir_models + '/model_textline_new.h5'
        self._device, self._tile, self._precision = device, tile, precision
        self._cache, self._max_batch = cache_models, max_batch

    # ------------------------------------------------------------------ helpers (main.py:112, 174, 178)
    def resize_image(self, img_in, input_height, input_width):
        return cv2.resize(img_in, (input_width, input_height), interpolation=cv2.INTER_NEAREST)

    def crop_image_inside_box(self, box, img_org_copy):
        image_box = img_org_copy[box[1]:box[1] + box[3], box[0]:box[0] + box[2]]
        return image_box, [box[1], box[1] + box[3], box[0], box[0] + box[2]]

    def otsu_copy(self, img):
        # main.py:178-194: three thresholds are computed but channel 0's result fills all channels
        img_r = np.zeros(img.shape)
        _, threshold1 = cv2.threshold(img[:, :, 0], 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        img_r[:, :, 0] = threshold1
        img_r[:, :, 1] = threshold1
        img_r[:, :, 2] = threshold1
        return img_r

    def get_image_and_scales(self):
        self.image = cv2.imread(self.image_dir)
        if self.image is None:
            raise FileNotFoundError(self.image_dir)
        self.height_org = self.image.shape[0]
        self.width_org = self.image.shape[1]
        if self.image.shape[0] < 2500:
            self.img_hight_int = 2800
        else:
            self.img_hight_int = int(self.image.shape[0] * 1.2)
        self.img_width_int = int(self.img_hight_int * self.image.shape[1] / float(self.image.shape[0]))
        self.scale_y = self.img_hight_int / float(self.image.shape[0])
        self.scale_x = self.img_width_int / float(self.image.shape[1])
        self.image = self.resize_image(self.image, self.img_hight_int, self.img_width_int)

    # ------------------------------------------------------------------ model lifecycle (main.py:216-223)
    def start_new_session_and_model(self, model_dir):
        key = (os.path.abspath(model_dir), self._device, self._tile, self._precision,
               os.environ.get("SBB_SYNTHETIC_MODELS"))
        if self._cache:
            with _MODEL_CACHE_LOCK:      # one thread loads a model (tens of GB of workspace), the others wait for it
                if key not in _MODEL_CACHE:
                    _MODEL_CACHE[key] = load_model_file(model_dir, self._tile, self._device, self._precision, self._max_batch)
                return _MODEL_CACHE[key], _NullSession()
        model = load_model_file(model_dir, self._tile, self._device, self._precision, self._max_batch)
        return model, SbbSession(model)

    # ------------------------------------------------------------------ the hot path (main.py:225-380)
    def do_prediction(self, patches, img, model):
        """uint8 BGR [H,W,3] -> uint8 [H,W,3] label image (class id in all 3 channels).
        An SbbModel runs the fused GPU call; any other object with ``layers``/``predict`` (e.g. a real
        Keras model) goes through the reference's own loop, restated below for that case."""
        img_height_model = model.layers[len(model.layers) - 1].output_shape[1]
        img_width_model = model.layers[len(model.layers) - 1].output_shape[2]
        if isinstance(model, SbbModel):
            if patches:
                if img.shape[0] < img_height_model or img.shape[1] < img_width_model:
                    raise ValueError("image smaller than the model tile: the reference's negative-origin "
                                     "slicing (main.py:276-281) is undefined here")
                seg = model.predict_page(np.ascontiguousarray(img, dtype=np.uint8))
                return np.repeat(seg[:, :, np.newaxis], 3, axis=2)
            # main.py:368-379.  /255 and INTER_NEAREST commute (both are per-element), so resize the
            # uint8 image first and let the GPU path normalise.
            small = self.resize_image(np.ascontiguousarray(img, dtype=np.uint8), img_height_model, img_width_model)
            seg = model.predict_full(small)
            seg_color = np.repeat(seg[:, :, np.newaxis], 3, axis=2)
            return self.resize_image(seg_color, self.image.shape[0], self.image.shape[1]).astype(np.uint8)
        return _do_prediction_generic(self, patches, img, model)

    # ------------------------------------------------------------------ device-resident page
    # the page is uploaded ONCE; the three stages read crops of the device copy and
    # the byte operations between them (nearest resize, Otsu, dilate) run on the GPU (prepost.py), so the
    # 134 MB float64 temporaries of the reference (main.py:180-193, :239) never exist and only label maps
    # travel back.  Results are bit-identical to the host statements they replace (tests/test_prepost.py).
    def _device_page(self):
        import torch
        host = self.image
        if getattr(self, "_host_page", None) is not host:
            self._host_page = host
            self._dev_page = torch.from_numpy(np.ascontiguousarray(host, dtype=np.uint8)).to(f"cuda:{self._device}")
        return self._dev_page

    def _device_view(self, img):
        """The device twin of ``img`` when it is a crop (numpy view) of the cached host page, else an upload."""
        import torch
        host = getattr(self, "_host_page", None)
        if host is not None and isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.ndim == 3 \
                and img.strides == host.strides and np.may_share_memory(img, host):
            off = img.__array_interface__["data"][0] - host.__array_interface__["data"][0]
            y0, rem = divmod(off, host.strides[0])
            x0, c = divmod(rem, host.strides[1])
            if off >= 0 and c == 0 and img.shape[2] == 3 and y0 + img.shape[0] <= host.shape[0] and x0 + img.shape[1] <= host.shape[1]:
                return self._dev_page[y0:y0 + img.shape[0], x0:x0 + img.shape[1]]
        return torch.from_numpy(np.ascontiguousarray(img, dtype=np.uint8)).to(f"cuda:{self._device}")

    # ------------------------------------------------------------------ stage drivers
    def extract_page(self):
        patches = False
        model_page, session_page = self.start_new_session_and_model(self.model_page_dir)
        img = self.image
        if isinstance(model_page, SbbModel):
            from . import prepost
            mh, mw = model_page.layers[-1].output_shape[1:3]
            d_page = self._device_page()
            seg = model_page.predict_full(prepost.resize_nearest(d_page, mh, mw))          # main.py:371-376
            full = prepost.resize_nearest(seg, img.shape[0], img.shape[1])                 # main.py:378
            # main.py:394-397: gray of (v,v,v) is v; dilation (a max filter) commutes with the >0 threshold
            grown = prepost.dilate(full, iterations=6).cpu().numpy()
            _, thresh = cv2.threshold(grown, 0, 255, 0)
        else:
            img_page_prediction = self.do_prediction(patches, img, model_page)
            imgray = cv2.cvtColor(img_page_prediction, cv2.COLOR_BGR2GRAY)
            _, thresh = cv2.threshold(imgray, 0, 255, 0)
            thresh = cv2.dilate(thresh, self.kernel, iterations=6)
        contours, _ = cv2.findContours(thresh, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        # main.py:400-404 sit OUTSIDE the reference's try: an all-background border map has no contour and
        # np.argmax of the empty size list raises ValueError out of extract_page (and out of run()), as there
        cnt_size = np.array([cv2.contourArea(contours[j]) for j in range(len(contours))])
        cnt = contours[np.argmax(cnt_size)]
        x, y, w, h = cv2.boundingRect(cnt)
        try:
            box = [x, y, w, h]
            croped_page, page_coord = self.crop_image_inside_box(box, self.image)
        except Exception:  # main.py:417-419: whole image
            box = [0, 0, self.image.shape[1] - 1, self.image.shape[0] - 1]
            croped_page, page_coord = self.crop_image_inside_box(box, self.image)
        self.cont_page = [np.array([[page_coord[2], page_coord[0]], [page_coord[3], page_coord[0]],
                                    [page_coord[3], page_coord[1]], [page_coord[2], page_coord[1]]])]
        session_page.close()
        # main.py:431: the page attribute does not survive this stage (the crop returned above is a view of
        # it and keeps the pixels alive; the device twin stays cached for the next two stages)
        del self.image
        return croped_page, page_coord

    def extract_text_regions(self, img):
        patches = True
        model_region, session_region = self.start_new_session_and_model(self.model_region_dir)
        if isinstance(model_region, SbbModel) and self._fits(img, model_region):
            from . import prepost
            binar = prepost.otsu_copy(self._device_view(img))                              # main.py:443-444 on the GPU
            seg = model_region.predict_page(binar)
            # the x3 channel repeat of main.py:292 is done before the copy back (one pass instead of a 9 ms np.repeat)
            prediction_regions = seg[:, :, None].expand(-1, -1, 3).contiguous().cpu().numpy()
        else:
            img = self.otsu_copy(img)
            img = img.astype(np.uint8)
            prediction_regions = self.do_prediction(patches, img, model_region)
        session_region.close()
        return prediction_regions

    def textline_contours(self, img):
        patches = True
        model_textline, session_textline = self.start_new_session_and_model(self.model_textline_dir)
        if isinstance(model_textline, SbbModel) and self._fits(img, model_textline):
            seg = model_textline.predict_page(self._device_view(img)).cpu().numpy()       # channel 0 is all :503 returns
            session_textline.close()
            return seg
        img = img.astype(np.uint8)
        prediction_textline = self.do_prediction(patches, img, model_textline)
        session_textline.close()
        return prediction_textline[:, :, 0]

    @staticmethod
    def _fits(img, model):
        """Pages smaller than the tile take the generic route, whose error message names the reference's
        undefined negative-origin slicing (do_prediction)."""
        return img.shape[0] >= model.tile_h and img.shape[1] >= model.tile_w

    # ------------------------------------------------------------------ host-glue entry points restated here
    def rotate_image(self, img_patch, slope):
        (h, w) = img_patch.shape[:2]
        M = cv2.getRotationMatrix2D((w // 2, h // 2), slope, 1.0)
        return cv2.warpAffine(img_patch, M, (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)

    def return_deskew_slope(self, img_patch, sigma_des):
        """main.py:1601-1718 with the rotation search on the GPU (deskew.py); identical angle."""
        from . import deskew
        return deskew.return_deskew_slope(img_patch, sigma_des, device=self._device)

    def write_into_page_xml(self, contours, page_coord, dir_of_image, order_of_texts, id_of_texts):
        """main.py:1908-2053 (page_xml.py); byte-identical output."""
        from . import page_xml
        return page_xml.write_into_page_xml(self, contours, page_coord, dir_of_image, order_of_texts, id_of_texts)

    def run_segmentation(self):
        """The three model stages of ``run()`` (main.py:2056-2107) without the contour / deskew / XML
        glue: returns (page_coord, region label image, textline mask) on the cropped page."""
        self.get_image_and_scales()
        image_page, page_coord = self.extract_page()
        text_regions = self.extract_text_regions(image_page)
        textline_mask = self.textline_contours(image_page)
        return page_coord, text_regions, textline_mask


def _do_prediction_generic(self, patches, img, model):
    """The reference loop (main.py:231-380) for duck-typed models; kept so a real Keras model (or a
    test double) produces the reference's result through the same class."""
    mh = model.layers[len(model.layers) - 1].output_shape[1]
    mw = model.layers[len(model.layers) - 1].output_shape[2]
    if not patches:
        imgf = self.resize_image(img / float(255.0), mh, mw)
        seg = np.argmax(model.predict(imgf.reshape(1, mh, mw, 3)), axis=3)[0]
        seg_color = np.repeat(seg[:, :, np.newaxis], 3, axis=2)
        return self.resize_image(seg_color.astype(np.uint8), self.image.shape[0], self.image.shape[1])
    margin = int(0.1 * mw)
    wm, hm = mw - 2 * margin, mh - 2 * margin
    imgf = img / float(255.0)
    H, Wd = imgf.shape[:2]
    out = np.zeros((H, Wd), np.uint8)
    nxf, nyf = -(-Wd // wm), -(-H // hm)
    for i in range(nxf):
        for j in range(nyf):
            x0, y0 = min(i * wm, Wd - mw), min(j * hm, H - mh)
            seg = np.argmax(model.predict(imgf[y0:y0 + mh, x0:x0 + mw][None]), axis=3)[0]
            ax = 0 if i == 0 else margin
            bx = mw - margin if i == 0 else (mw if i == nxf - 1 else mw - margin)
            ay = 0 if j == 0 else margin
            by = mh - margin if j == 0 else (mh if j == nyf - 1 else mh - margin)
            out[y0 + ay:y0 + by, x0 + ax:x0 + bx] = seg[ay:by, ax:bx]
    return np.repeat(out[:, :, np.newaxis], 3, axis=2)
