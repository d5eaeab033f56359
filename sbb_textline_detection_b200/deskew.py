"""Deskew search of the reference on the GPU.

    return_deskew_slope(img_patch, sigma_des)      main.py:1601-1718
        rotate_image                                main.py:159-163   (cv2.warpAffine, INTER_CUBIC, BORDER_REPLICATE)
        get_standard_deviation_of_summed_textline_patch_along_width   main.py:1545-1599

The reference rotates a zero-padded float64 copy of every text region's textline mask 80 (+30) times on
the CPU -- its largest remaining cost, hidden behind ``cpu_count()`` forked workers (main.py:1760-1799).
Here ONE kernel launch (``sbb_rotate_rowsum_u8``) produces the binarised row profiles of all candidate
angles, bit-identical to OpenCV, and only the 1-D profile statistics (scipy gaussian_filter1d /
find_peaks, as in the reference) stay on the host.  The chosen angle is therefore identical to the
reference's, including its quirk of indexing ``angles`` with a position in the NaN-filtered list
(main.py:1665-1667).  No CPU fallback: raises if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import warnings

import cv2
import numpy as np
from scipy.ndimage import gaussian_filter1d
from scipy.signal import find_peaks

from . import _lib

_HUGE = 1000000000000000000000  # the reference's "no usable peaks" marker (main.py:1640)


def inverse_affine(M) -> np.ndarray:
    """The dst->src map cv2.warpAffine derives from a 2x3 matrix (no WARP_INVERSE_MAP), same
    statement order in double precision."""
    m = np.asarray(M, np.float64).ravel().copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def padded_geometry(h: int, w: int):
    """main.py:1611-1617: side of the padded square and where the patch sits inside it."""
    side = int(max(h, w) * 1.4)
    c = int(side / 2.0)
    return side, c - int(h / 2.0), c - int(w / 2.0)


def rotation_profiles(mask, angles, device: int = 0) -> np.ndarray:
    """int32 [len(angles), S]: row sums of the binarised rotations of the padded mask (see module doc).
    ``mask``: uint8 [h, w] numpy array (staged through CUDA device ``device``) or CUDA torch tensor (runs where it
    lives), two-valued (0 / non-zero)."""
    h, w = int(mask.shape[0]), int(mask.shape[1])
    side, oy, ox = padded_geometry(h, w)
    center = (side // 2, side // 2)
    inv = np.ascontiguousarray(np.stack([inverse_affine(cv2.getRotationMatrix2D(center, float(a), 1.0))
                                         for a in angles]))
    n = inv.shape[0]
    lib = _lib.lib()
    if isinstance(mask, np.ndarray):
        src = np.ascontiguousarray(mask, dtype=np.uint8)
        out = np.empty((n, side), np.int32)
        _lib.check(lib.sbb_rotate_rowsum_u8(src.ctypes.data_as(C.c_void_p), h, w, src.strides[0], side, oy, ox,
                                            inv.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p),
                                            _lib.SBB_MEM_HOST, int(device), None))
        return out
    import torch
    assert mask.is_cuda and mask.dtype == torch.uint8
    src = mask.contiguous()
    out = torch.empty((n, side), dtype=torch.int32, device=mask.device)
    stream = torch.cuda.current_stream(mask.device).cuda_stream or 1
    _lib.check(lib.sbb_rotate_rowsum_u8(C.c_void_p(src.data_ptr()), h, w, src.stride(0), side, oy, ox,
                                        inv.ctypes.data_as(C.c_void_p), n, C.c_void_p(out.data_ptr()),
                                        _lib.SBB_MEM_DEVICE, mask.device.index or 0, C.c_void_p(stream)))
    return out.cpu().numpy()


def _best_angle(profiles: np.ndarray, angles: np.ndarray, sigma: float) -> float:
    """Per-angle statistics of main.py:1545-1599 for all candidate angles at once: the two Gaussian filters run
    over the whole [angles, rows] array (scipy filters line by line, so every row is bit-identical to the
    reference's per-profile call); peak finding and the acceptance rule stay per profile."""
    y = profiles.astype(np.float64)
    n_ang, n = y.shape
    framed = np.zeros((n_ang, n + 20))
    framed[:, 10:n + 10] = y
    inverted = -framed + np.max(framed, axis=1, keepdims=True)
    inv_framed = np.zeros((n_ang, n + 40))
    inv_framed[:, 10:n + 30] = inverted
    z_all = gaussian_filter1d(y, sigma, axis=1)
    z_inv_all = gaussian_filter1d(inv_framed, sigma, axis=1)
    spread = []
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for z, z_inv in zip(z_all, z_inv_all):
            try:
                valleys, _ = find_peaks(z_inv, height=0)
                crests, _ = find_peaks(z, height=0)
                valleys = valleys - 10 - 10
                crest_vals = z[crests]
                crest_vals = crest_vals[crest_vals > 10]
                valley_vals = z[valleys]            # IndexError exactly where the reference's fancy indexing raises
                mean_crest = np.mean(crest_vals)
                limit = mean_crest - (mean_crest - 0) / 20.3
                score = np.mean(valley_vals[valley_vals < limit])
                sd = np.std(z)
                if score == 0:
                    score = _HUGE
            except Exception:
                score, sd = _HUGE, 0
            if score != score:      # NaN: the reference drops this angle from the list (main.py:1650-1652)
                continue
            spread.append(sd)
    try:
        return angles[np.argmax(np.array(spread))]  # position in the FILTERED list, as in the reference
    except Exception:
        return 0


def is_two_valued(img_patch) -> bool:
    """The GPU search rotates the BINARISED patch; the reference cubic-interpolates the patch's values and tests
    ``!= 0`` afterwards (main.py:1631-1632).  The two agree bit for bit when the patch holds zero and ONE other
    value (what the textline mask crops of main.py:1729-1733 are); with several non-zero values the negative
    lobes of the cubic kernel could cancel differently, so callers keep the reference's CPU search for those."""
    vals = np.unique(np.asarray(img_patch))
    return vals.size <= 1 or (vals.size == 2 and vals[0] == 0)


def return_deskew_slope(img_patch, sigma_des, device: int = 0):
    """Same result as the reference's ``textline_detector.return_deskew_slope`` (main.py:1601-1718) for a
    two-valued patch (``is_two_valued``; anything else raises ValueError -- there is no silent approximation)."""
    if isinstance(img_patch, np.ndarray):
        if not is_two_valued(img_patch):
            raise ValueError("GPU deskew search needs a two-valued mask (0 / one non-zero value)")
        mask = (np.asarray(img_patch) != 0).astype(np.uint8)
    else:
        mask = img_patch
    angles = np.linspace(-25, 25, 80)
    ang = _best_angle(rotation_profiles(mask, angles, device), angles, sigma_des)
    if abs(ang) > 15:
        angles = np.linspace(-90, -50, 30)
        ang = _best_angle(rotation_profiles(mask, angles, device), angles, sigma_des)
    return ang
