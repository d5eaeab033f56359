"""Synthetic page generators: there is no network for real scans or weights,
so tests and bench.py use these.  Pages are uint8 [H,W,3] in BGR order, exactly what the reference's
``do_prediction`` receives after ``cv2.imread`` + ``get_image_and_scales`` (main.py:196-214)."""
from __future__ import annotations

import numpy as np


def uniform_page(h: int, w: int, seed: int = 0) -> np.ndarray:
    """(i) adversarial: i.i.d. uniform uint8."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def document_page(h: int, w: int, seed: int = 0) -> np.ndarray:
    """(ii) document-like: bright noisy background, 1-3 columns of dark 'text line' bars made of
    random-width glyph blobs, mild per-page skew."""
    rng = np.random.default_rng(1000 + seed)
    page = 235.0 + rng.normal(0.0, 6.0, size=(h, w)).astype(np.float32)
    ncol = int(rng.integers(1, 4))
    mx, my = int(0.06 * w), int(0.05 * h)
    gap = int(0.03 * w)
    colw = (w - 2 * mx - (ncol - 1) * gap) // ncol
    skew = float(rng.uniform(-0.02, 0.02))  # rise per pixel in x (about +-1 degree)
    for c in range(ncol):
        x_lo = mx + c * (colw + gap)
        nlines = int(rng.integers(30, 61))
        pitch = (h - 2 * my) / nlines
        lh = max(4, int(pitch * rng.uniform(0.35, 0.55)))
        for ln in range(nlines):
            if rng.random() < 0.08:  # paragraph gap
                continue
            y_base = my + int(ln * pitch)
            x = x_lo + (int(rng.integers(0, colw // 8)) if rng.random() < 0.15 else 0)
            x_end = x_lo + colw - (int(rng.integers(0, colw // 2)) if rng.random() < 0.2 else 0)
            while x < x_end:
                gw = int(rng.integers(3, 14))
                if rng.random() < 0.18:  # inter-word space
                    x += gw
                    continue
                gh = int(lh * rng.uniform(0.6, 1.0))
                y0 = y_base + int(skew * (x - x_lo)) + (lh - gh)
                y0 = max(0, min(h - gh - 1, y0))
                x1 = min(x + gw, w)
                page[y0:y0 + gh, x:x1] = rng.uniform(20, 70)
                x += gw + int(rng.integers(1, 3))
    img = np.clip(page, 0, 255).astype(np.uint8)
    out = np.stack([img, img, img], axis=2)
    tint = rng.integers(-6, 7, size=3)
    out = np.clip(out.astype(np.int16) + tint[None, None, :], 0, 255).astype(np.uint8)
    return np.ascontiguousarray(out)


def framed_page(h: int, w: int, seed: int = 0, frame: int = 120) -> np.ndarray:
    """(iii) a document page inside a dark, noisy scanner border of ``frame`` pixels -- what the border model
    of the pipeline crops away (main.py:384-437), so that every page gets its own crop geometry."""
    inner = document_page(h - 2 * frame, w - 2 * frame, seed=seed)
    rng = np.random.default_rng(2000 + seed)
    page = rng.integers(8, 40, size=(h, w, 3), dtype=np.uint8)
    page[frame:h - frame, frame:w - frame] = inner
    return np.ascontiguousarray(page)
