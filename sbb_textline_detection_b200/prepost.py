"""The cv2 calls the reference applies around its three models, on the GPU:

    resize_nearest   cv2.resize(..., INTER_NEAREST)      main.py:112-113 (:214, :371, :378)
    otsu_copy        textline_detector.otsu_copy         main.py:178-194
    erode / dilate   cv2.erode / cv2.dilate, 5x5 ones    main.py:397, 2074-2075

Each accepts a numpy uint8 array (host: staged through the device, returns numpy) or a torch CUDA
uint8 tensor (device-resident, asynchronous on the current torch stream, returns a tensor).  Results
are bit-identical to OpenCV (tests/test_prepost.py).  No CPU fallback: raises if the CUDA library
is missing."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _as3(img):
    return img if img.ndim == 3 else img[:, :, None]


def _io(img, out_shape, device: int = 0):
    """-> (src ptr, src row stride, dst array, dst ptr, dst row stride, memkind, device, stream); ``device`` is
    the CUDA device a HOST array is staged through (a CUDA tensor runs where it lives)"""
    if isinstance(img, np.ndarray):
        src = np.ascontiguousarray(img, dtype=np.uint8)
        dst = np.empty(out_shape, np.uint8)
        return (src, src.ctypes.data_as(C.c_void_p), src.strides[0], dst, dst.ctypes.data_as(C.c_void_p), dst.strides[0],
                _lib.SBB_MEM_HOST, int(device), None)
    import torch
    assert img.is_cuda and img.dtype == torch.uint8
    src = img.contiguous()
    dst = torch.empty(out_shape, dtype=torch.uint8, device=img.device)
    stream = torch.cuda.current_stream(img.device).cuda_stream or 1  # 0 -> cudaStreamLegacy
    return (src, C.c_void_p(src.data_ptr()), src.stride(0), dst, C.c_void_p(dst.data_ptr()), dst.stride(0),
            _lib.SBB_MEM_DEVICE, img.device.index or 0, C.c_void_p(stream))


def resize_nearest(img, out_h: int, out_w: int, device: int = 0):
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_NEAREST) for uint8 [H,W] or [H,W,C]."""
    H, W = int(img.shape[0]), int(img.shape[1])
    Cn = 1 if img.ndim == 2 else int(img.shape[2])
    shape = (out_h, out_w) if img.ndim == 2 else (out_h, out_w, Cn)
    keep, ps, ss, dst, pd, ds, kind, dev, st = _io(img, shape, device)
    _lib.check(_lib.lib().sbb_resize_nearest_u8(ps, H, W, Cn, ss, pd, out_h, out_w, ds, kind, dev, st))
    return dst


def otsu_copy(img, return_threshold: bool = False, device: int = 0):
    """main.py:178-194: Otsu of channel 0, written to all 3 channels as 0/255 (uint8 [H,W,3])."""
    H, W = int(img.shape[0]), int(img.shape[1])
    Cn = 1 if img.ndim == 2 else int(img.shape[2])
    keep, ps, ss, dst, pd, ds, kind, dev, st = _io(img, (H, W, 3), device)
    thr = C.c_int32(-1)
    _lib.check(_lib.lib().sbb_otsu_copy_u8(ps, H, W, Cn, ss, pd, ds, C.byref(thr) if return_threshold else None,
                                           kind, dev, st))
    return (dst, thr.value) if return_threshold else dst


def _morph(img, op: int, iterations: int, device: int = 0):
    H, W = int(img.shape[0]), int(img.shape[1])
    Cn = 1 if img.ndim == 2 else int(img.shape[2])
    keep, ps, ss, dst, pd, ds, kind, dev, st = _io(img, tuple(img.shape), device)
    _lib.check(_lib.lib().sbb_morph5x5_u8(ps, H, W, Cn, ss, pd, ds, op, iterations, kind, dev, st))
    return dst


def erode(img, iterations: int = 1, device: int = 0):
    """cv2.erode(img, np.ones((5, 5), np.uint8), iterations=iterations)"""
    return _morph(img, 0, iterations, device)


def dilate(img, iterations: int = 1, device: int = 0):
    """cv2.dilate(img, np.ones((5, 5), np.uint8), iterations=iterations)"""
    return _morph(img, 1, iterations, device)
