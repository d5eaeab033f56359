"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of
``textline_detector.do_prediction`` (qurator/sbb_textline_detector/main.py:225-380) and of the small
helpers its three callers apply around it (``otsu_copy`` main.py:178-194, ``resize_image``
main.py:112-113, the scale rule of ``get_image_and_scales`` main.py:196-214).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.

PINNED against the reference's own execution: tests/golden/ref_stitch_fake.npz, ref_nopatch_fake.npz and
ref_prepost.npz were minted by running the UNMODIFIED reference class from /root/reference (tensorflow / keras
stubbed, tests/golden/make_golden_from_reference.py) and tests/test_reference_golden.py holds this module
to them bit for bit, incl. the BASELINE config-2 and config-5 grids.  ``do_prediction`` below
is a literal replay of the reference's nested loop and its 9-case if/elif chain (the statement order
is kept so the last-writer-wins overlap behaves identically); ``model`` is anything with the two
attributes the reference touches (``layers[-1].output_shape`` and ``predict``), e.g.
``oracle.resnet50_unet.KerasLikeModel``.
"""
from __future__ import annotations

import numpy as np


def resize_nearest(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(..., interpolation=cv2.INTER_NEAREST) (main.py:112-113): source index =
    min(floor(dst * src/dst_size), src-1) per axis, computed in double precision like OpenCV."""
    h, w = img.shape[:2]
    ys = np.minimum(np.floor(np.arange(out_h) * (h / float(out_h))).astype(np.int64), h - 1)
    xs = np.minimum(np.floor(np.arange(out_w) * (w / float(out_w))).astype(np.int64), w - 1)
    return img[ys][:, xs]


def scaled_size(h: int, w: int):
    """main.py:201-207: pages shorter than 2500 px go to height 2800, others are scaled by 1.2."""
    nh = 2800 if h < 2500 else int(h * 1.2)
    nw = int(nh * w / float(h))
    return nh, nw


def otsu_threshold_u8(ch: np.ndarray) -> int:
    """Otsu's threshold as cv2.threshold(..., THRESH_OTSU) computes it for uint8 (main.py:187):
    maximise between-class variance over the 256-bin histogram; first maximum wins."""
    hist = np.bincount(ch.ravel(), minlength=256).astype(np.float64)
    total = hist.sum()
    scale = 1.0 / total
    mu = float((np.arange(256) * hist).sum()) * scale
    q1, mu1, best, thr = 0.0, 0.0, 0.0, 0
    for i in range(256):
        p_i = hist[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < np.finfo(np.float32).eps or max(q1, q2) > 1.0 - np.finfo(np.float32).eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > best:
            best, thr = sigma, i
    return thr


def otsu_copy(img: np.ndarray) -> np.ndarray:
    """main.py:178-194 incl. its quirk: the threshold of channel 0 is written to all 3 channels."""
    thr = otsu_threshold_u8(img[:, :, 0])
    b = np.where(img[:, :, 0] > thr, 255, 0).astype(np.uint8)
    return np.stack([b, b, b], axis=2)


def tile_grid(img_h: int, img_w: int, mh: int, mw: int, margin: int | None = None):
    """main.py:233-281: returns (margin, nxf, nyf, list of (i, j, x0, y0) in loop order)."""
    if margin is None:
        margin = int(0.1 * mw)
    width_mid = mw - 2 * margin
    height_mid = mh - 2 * margin
    nxf = img_w / float(width_mid)
    nyf = img_h / float(height_mid)
    nxf = int(nxf) + 1 if nxf > int(nxf) else int(nxf)
    nyf = int(nyf) + 1 if nyf > int(nyf) else int(nyf)
    tiles = []
    for i in range(nxf):
        for j in range(nyf):
            x0 = i * width_mid
            x1 = x0 + mw
            y0 = j * height_mid
            y1 = y0 + mh
            if x1 > img_w:
                x1 = img_w
                x0 = img_w - mw
            if y1 > img_h:
                y1 = img_h
                y0 = img_h - mh
            tiles.append((i, j, x0, y0))
    return margin, nxf, nyf, tiles


def stitch_replay(img_h, img_w, mh, mw, margin, nxf, nyf, tiles, seg_of_tile):
    """main.py:294-364: the 9-case crop + overwrite, replayed literally.  ``seg_of_tile(t, i, j, x0, y0)``
    returns the [mh, mw] class map of tile t."""
    prediction_true = np.zeros((img_h, img_w, 3))
    for t, (i, j, x0, y0) in enumerate(tiles):
        seg = seg_of_tile(t, i, j, x0, y0)
        seg_color = np.repeat(seg[:, :, np.newaxis], 3, axis=2)
        x1, y1 = x0 + mw, y0 + mh
        if i == 0 and j == 0:
            prediction_true[y0 + 0:y1 - margin, x0 + 0:x1 - margin, :] = seg_color[0:mh - margin, 0:mw - margin, :]
        elif i == nxf - 1 and j == nyf - 1:
            prediction_true[y0 + margin:y1 - 0, x0 + margin:x1 - 0, :] = seg_color[margin:mh, margin:mw, :]
        elif i == 0 and j == nyf - 1:
            prediction_true[y0 + margin:y1 - 0, x0 + 0:x1 - margin, :] = seg_color[margin:mh, 0:mw - margin, :]
        elif i == nxf - 1 and j == 0:
            prediction_true[y0 + 0:y1 - margin, x0 + margin:x1 - 0, :] = seg_color[0:mh - margin, margin:mw, :]
        elif i == 0 and j != 0 and j != nyf - 1:
            prediction_true[y0 + margin:y1 - margin, x0 + 0:x1 - margin, :] = seg_color[margin:mh - margin, 0:mw - margin, :]
        elif i == nxf - 1 and j != 0 and j != nyf - 1:
            prediction_true[y0 + margin:y1 - margin, x0 + margin:x1 - 0, :] = seg_color[margin:mh - margin, margin:mw, :]
        elif i != 0 and i != nxf - 1 and j == 0:
            prediction_true[y0 + 0:y1 - margin, x0 + margin:x1 - margin, :] = seg_color[0:mh - margin, margin:mw - margin, :]
        elif i != 0 and i != nxf - 1 and j == nyf - 1:
            prediction_true[y0 + margin:y1 - 0, x0 + margin:x1 - margin, :] = seg_color[margin:mh, margin:mw - margin, :]
        else:
            prediction_true[y0 + margin:y1 - margin, x0 + margin:x1 - margin, :] = seg_color[margin:mh - margin, margin:mw - margin, :]
    return prediction_true.astype(np.uint8)


def do_prediction(patches: bool, img: np.ndarray, model, full_shape=None, margin: int | None = None,
                  predict_batch: int = 1):
    """main.py:225-380.  ``img`` uint8 [H,W,3] BGR.  Returns uint8 [H,W,3] (class id x3).
    ``full_shape`` is ``self.image.shape`` for the patches=False branch (main.py:378).
    ``predict_batch`` > 1 batches the model.predict calls (results are per-tile independent); the
    reference itself is batch-1 (main.py:287-288)."""
    mh, mw, _ = model.layers[-1].output_shape[1:4]
    if patches:
        imgf = img / float(255.0)
        img_h, img_w = imgf.shape[0], imgf.shape[1]
        margin, nxf, nyf, tiles = tile_grid(img_h, img_w, mh, mw, margin)
        segs = {}

        def flush(pending):
            batch = np.stack([imgf[y0:y0 + mh, x0:x0 + mw, :] for (_, _, _, x0, y0) in pending])
            pred = model.predict(batch)
            for (t, _, _, _, _), p in zip(pending, pred):
                segs[t] = np.argmax(p, axis=2)

        pending = []
        for t, (i, j, x0, y0) in enumerate(tiles):
            pending.append((t, i, j, x0, y0))
            if len(pending) == predict_batch:
                flush(pending)
                pending = []
        if pending:
            flush(pending)
        return stitch_replay(img_h, img_w, mh, mw, margin, nxf, nyf, tiles, lambda t, *_: segs[t])
    imgf = img / float(255.0)
    imgf = resize_nearest(imgf, mh, mw)
    label_p_pred = model.predict(imgf.reshape(1, mh, mw, 3))
    seg = np.argmax(label_p_pred, axis=3)[0]
    seg_color = np.repeat(seg[:, :, np.newaxis], 3, axis=2)
    fh, fw = (full_shape or img.shape)[:2]
    return resize_nearest(seg_color, fh, fw).astype(np.uint8)
