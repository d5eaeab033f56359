"""Model handle: the replacement for the (keras model, tf session) pair that
``textline_detector.start_new_session_and_model`` returns (main.py:216-223).

``SbbModel`` duck-types the three things ``do_prediction`` touches on a Keras model
(``.layers[-1].output_shape`` main.py:227-229, ``.predict`` main.py:287-288/373-374) so that even the
UNMODIFIED reference loop runs on it, and adds the fused page call the drop-in detector uses."""
from __future__ import annotations

import ctypes as C
import functools
import threading

import numpy as np

from . import _lib
from . import weights as W


class _Layer:
    def __init__(self, shape):
        self.output_shape = shape


class SbbSession:
    """Stands in for the tf.InteractiveSession of main.py:220; ``close`` frees the GPU model."""

    def __init__(self, model: "SbbModel"):
        self._model = model

    def close(self):
        self._model.close()


def _ptr(a):
    """numpy array -> (void*, SBB_MEM_HOST); torch CUDA tensor -> (void*, SBB_MEM_DEVICE)."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p), _lib.SBB_MEM_HOST
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr()), (_lib.SBB_MEM_DEVICE if a.is_cuda else _lib.SBB_MEM_HOST)
    raise TypeError(type(a))


def _serialised(fn):
    """One native call per handle at a time: the handle owns mutable per-call state (geometry cache, activation
    workspace, staging buffers), and detector instances on several threads share cached models.  On the DEVICE the
    C library orders the calls of one handle itself (an event chain across streams, sbb_net.cu: chain_begin)."""
    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        with self._lock:
            return fn(self, *a, **k)
    return wrapper


class SbbModel:
    def __init__(self, weights: dict | bytes, tile_h: int, tile_w: int, n_classes: int, *, device: int = 0,
                 precision: str = "fp16x3", backend: str = "tcgen05", max_batch: int = 48):
        self._h = None
        self._lock = threading.RLock()
        blob = weights if isinstance(weights, (bytes, bytearray)) else W.pack_blob(weights, n_classes, (tile_h, tile_w))
        self._blob = np.frombuffer(blob, dtype=np.uint8)
        desc = _lib.ModelDesc()
        desc.tile_h, desc.tile_w, desc.n_classes = tile_h, tile_w, n_classes
        desc.precision = {"fp16x3": _lib.SBB_PREC_FP16X3, "fp16": _lib.SBB_PREC_FP16}[precision]
        desc.backend = {"tcgen05": _lib.SBB_BACKEND_TCGEN05, "simt": _lib.SBB_BACKEND_SIMT}[backend]
        desc.device, desc.max_batch = device, max_batch
        desc.weights = self._blob.ctypes.data_as(C.c_void_p)
        desc.weights_nbytes = self._blob.size
        h = C.c_void_p()
        _lib.check(_lib.lib().sbb_model_create(C.byref(desc), C.byref(h)))
        self._h = h
        self._blob = None
        self.tile_h, self.tile_w, self.n_classes = tile_h, tile_w, n_classes
        self.precision, self.backend, self.device, self.max_batch = precision, backend, device, max_batch
        # what do_prediction reads at main.py:227-229
        self.layers = [_Layer((None, tile_h, tile_w, n_classes))]

    # -- lifecycle ----------------------------------------------------------------------------
    @_serialised
    def close(self):
        if self._h is not None:
            _lib.lib().sbb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if self._h is None:
            raise RuntimeError("model is closed")
        return self._h

    # -- keras-compatible ---------------------------------------------------------------------
    def predict(self, x):
        """model.predict: float [n,th,tw,3] in [0,1] -> softmax probabilities float32 [n,th,tw,C]."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 4 and x.shape[1:] == (self.tile_h, self.tile_w, 3), x.shape
        return self.predict_tiles(x, want_labels=False, want_probs=True)[1]

    # -- native entry points -------------------------------------------------------------------
    @_serialised
    def predict_tiles(self, x, want_labels=True, want_probs=False, want_logits=False):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        labels = np.empty((n, self.tile_h, self.tile_w), np.uint8) if want_labels else None
        probs = np.empty((n, self.tile_h, self.tile_w, self.n_classes), np.float32) if want_probs else None
        logits = np.empty((n, self.tile_h, self.tile_w, self.n_classes), np.float32) if want_logits else None
        _lib.check(_lib.lib().sbb_predict_tiles(self._handle(), _ptr(x)[0], n, _ptr(labels)[0], _ptr(probs)[0],
                                                _ptr(logits)[0], _lib.SBB_MEM_HOST, None))
        return labels, probs, logits

    @_serialised
    def predict_page(self, img, margin: int = -1, out=None, stream=None):
        """do_prediction(patches=True) core: uint8 BGR [H,W,3] -> uint8 label map [H,W].
        Accepts numpy arrays (host) or torch CUDA tensors (device-resident, asynchronous)."""
        H, Wd = int(img.shape[0]), int(img.shape[1])
        if isinstance(img, np.ndarray):
            img = np.ascontiguousarray(img, dtype=np.uint8)
            if out is None:
                out = np.empty((H, Wd), np.uint8)
            in_stride, out_stride = img.strides[0], out.strides[0]
        else:
            import torch
            # a crop of a larger device page is fine: pixels must be packed, rows may be strided
            assert img.dtype == torch.uint8 and img.stride(2) == 1 and img.stride(1) == 3, "need packed BGR pixels"
            if out is None:
                out = torch.empty((H, Wd), dtype=torch.uint8, device=img.device)
            in_stride, out_stride = img.stride(0), out.stride(0)
            # device-resident call: asynchronous on the caller's stream (default: torch's current stream).
            # torch's default stream has handle 0, which the C ABI reads as "the model's own stream":
            # name it explicitly (cudaStreamLegacy == 0x1) so that stream order with torch work holds.
            if stream is None:
                stream = torch.cuda.current_stream(img.device).cuda_stream
            if stream == 0:
                stream = 1
        pin, kind = _ptr(img)
        pout, kind2 = _ptr(out)
        assert kind == kind2, "input and output must live on the same side"
        _lib.check(_lib.lib().sbb_predict_page_tiled(self._handle(), pin, H, Wd, in_stride, margin, pout, out_stride,
                                                     kind, C.c_void_p(stream) if stream else None))
        return out

    @_serialised
    def predict_page_tile_range(self, img, labels, tile_first: int, tile_count: int, keep_labels: bool = True,
                                margin: int = -1, stream=None):
        """Tiles [tile_first, tile_first+tile_count) of the page grid (reference loop order) stitched into
        ``labels`` -- a CUDA uint8 [H,W] tensor or a raw device pointer (int), possibly ANOTHER GPU's memory opened
        with parallel.PeerBuffer: the head epilogue's stores then travel over NVLink.  Device buffers only."""
        import torch
        assert img.is_cuda and img.dtype == torch.uint8 and img.stride(2) == 1 and img.stride(1) == 3
        H, Wd = int(img.shape[0]), int(img.shape[1])
        if isinstance(labels, int):
            pout, out_stride = C.c_void_p(labels), Wd
        else:
            assert labels.is_cuda and labels.dtype == torch.uint8 and tuple(labels.shape) == (H, Wd)
            pout, out_stride = C.c_void_p(labels.data_ptr()), labels.stride(0)
        if stream is None:
            stream = torch.cuda.current_stream(img.device).cuda_stream
        _lib.check(_lib.lib().sbb_predict_page_tile_range(self._handle(), C.c_void_p(img.data_ptr()), H, Wd, img.stride(0),
                                                          margin, pout, out_stride, tile_first, tile_count,
                                                          1 if keep_labels else 0, C.c_void_p(stream or 1)))

    @_serialised
    def predict_pages_stacked(self, stack, n_pages: int, margin: int = -1, out=None, stream=None):
        """``n_pages`` same-size pages stacked vertically (uint8 [n_pages*H, W, 3], numpy or CUDA tensor) through the
        network as ONE batch (C ABI: sbb_predict_pages_stacked) -> stacked label maps uint8 [n_pages*H, W].  Each
        page is tiled and stitched exactly as by ``predict_page``; the per-launch costs are paid once for all pages
        when ``max_batch`` covers their tiles."""
        HS, Wd = int(stack.shape[0]), int(stack.shape[1])
        assert HS % n_pages == 0
        H = HS // n_pages
        if isinstance(stack, np.ndarray):
            stack = np.ascontiguousarray(stack, dtype=np.uint8)
            if out is None:
                out = np.empty((HS, Wd), np.uint8)
            in_stride, out_stride = stack.strides[0], out.strides[0]
        else:
            import torch
            assert stack.dtype == torch.uint8 and stack.stride(2) == 1 and stack.stride(1) == 3, "need packed BGR pixels"
            if out is None:
                out = torch.empty((HS, Wd), dtype=torch.uint8, device=stack.device)
            in_stride, out_stride = stack.stride(0), out.stride(0)
            if stream is None:
                stream = torch.cuda.current_stream(stack.device).cuda_stream
            if stream == 0:
                stream = 1
        pin, kind = _ptr(stack)
        pout, kind2 = _ptr(out)
        assert kind == kind2, "input and output must live on the same side"
        _lib.check(_lib.lib().sbb_predict_pages_stacked(self._handle(), pin, n_pages, H, Wd, in_stride, margin, pout,
                                                        out_stride, kind, C.c_void_p(stream) if stream else None))
        return out

    def pages_per_forward(self, H: int, Wd: int, margin: int = -1) -> int:
        """How many H x W pages one forward of this handle covers (max_batch // tiles per page, at least 1)."""
        nx, ny, _, _, _ = compute_tile_grid(H, Wd, self.tile_h, self.tile_w, margin)
        return max(1, self.max_batch // (nx * ny))

    def predict_pages(self, pages, outs=None, margin: int = -1):
        """Throughput form of ``predict_page`` for a batch of HOST pages (numpy uint8 [H,W,3] or CPU torch
        tensors; pinned memory makes the copies asynchronous): the H2D copies of the next pages and the D2H copies
        of the previous label maps run on their own streams while the current pages are in the network, so the PCIe
        time disappears behind the forward.  Consecutive pages of the same size share ONE forward (stacked call) as far
        as ``max_batch`` covers their tiles.  Returns the list of uint8 [H,W] label maps (``outs``: optional
        pre-allocated, ideally pinned, destinations -- numpy arrays or CPU tensors)."""
        import torch
        dev = torch.device("cuda", self.device)
        if getattr(self, "_streams", None) is None:
            self._streams = [torch.cuda.Stream(dev) for _ in range(3)]
        s_in, s_run, s_out = self._streams
        n = len(pages)
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(a)
        srcs = []
        for k in range(n):
            src = as_t(pages[k])
            assert src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3, tuple(src.shape)
            srcs.append(src.contiguous())
        # groups of consecutive same-size pages, as many as one forward covers
        groups, k = [], 0
        while k < n:
            H, Wd = int(srcs[k].shape[0]), int(srcs[k].shape[1])
            cap = self.pages_per_forward(H, Wd, margin)
            j = k + 1
            while j < n and j - k < cap and tuple(srcs[j].shape) == tuple(srcs[k].shape):
                j += 1
            groups.append((k, j))
            k = j
        res = [None] * n
        d_in, d_out = [None, None], [None, None]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_run = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        for gi, (k0, k1) in enumerate(groups):
            b = gi & 1
            cnt = k1 - k0
            H, Wd = int(srcs[k0].shape[0]), int(srcs[k0].shape[1])
            if d_in[b] is None or tuple(d_in[b].shape) != (cnt * H, Wd, 3):
                torch.cuda.synchronize(dev)
                d_in[b] = torch.empty((cnt * H, Wd, 3), dtype=torch.uint8, device=dev)
                d_out[b] = torch.empty((cnt * H, Wd), dtype=torch.uint8, device=dev)
            with torch.cuda.stream(s_in):
                if gi >= 2:
                    s_in.wait_event(ev_run[b])      # the forward that read this input buffer is done
                for i in range(cnt):
                    d_in[b][i * H:(i + 1) * H].copy_(srcs[k0 + i], non_blocking=True)
                ev_in[b].record(s_in)
            s_run.wait_event(ev_in[b])
            if gi >= 2:
                s_run.wait_event(ev_out[b])         # the D2H copies that read this output buffer are done
            if cnt == 1:
                self.predict_page(d_in[b], margin=margin, out=d_out[b], stream=s_run.cuda_stream)
            else:
                self.predict_pages_stacked(d_in[b], cnt, margin=margin, out=d_out[b], stream=s_run.cuda_stream)
            ev_run[b].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run[b])
                for i in range(cnt):
                    dst = as_t(outs[k0 + i]) if outs is not None else torch.empty((H, Wd), dtype=torch.uint8).pin_memory()
                    assert tuple(dst.shape) == (H, Wd) and dst.dtype == torch.uint8 and dst.is_contiguous()
                    dst.copy_(d_out[b][i * H:(i + 1) * H], non_blocking=True)
                    res[k0 + i] = dst
                ev_out[b].record(s_out)
        s_out.synchronize()
        s_run.synchronize()
        return [r.numpy() if (outs is None or not isinstance(outs[i], torch.Tensor)) else r for i, r in enumerate(res)]

    @_serialised
    def predict_full(self, img_tile, stream=None):
        """do_prediction(patches=False) core on an image already at tile size (numpy -> numpy, or a
        contiguous CUDA uint8 tensor -> CUDA tensor, asynchronous on ``stream`` / torch's current stream)."""
        assert tuple(img_tile.shape) == (self.tile_h, self.tile_w, 3), tuple(img_tile.shape)
        if isinstance(img_tile, np.ndarray):
            img_tile = np.ascontiguousarray(img_tile, dtype=np.uint8)
            out = np.empty((self.tile_h, self.tile_w), np.uint8)
            _lib.check(_lib.lib().sbb_predict_full(self._handle(), _ptr(img_tile)[0], _ptr(out)[0], _lib.SBB_MEM_HOST, None))
            return out
        import torch
        assert img_tile.is_cuda and img_tile.dtype == torch.uint8 and img_tile.is_contiguous()
        out = torch.empty((self.tile_h, self.tile_w), dtype=torch.uint8, device=img_tile.device)
        if stream is None:
            stream = torch.cuda.current_stream(img_tile.device).cuda_stream
        _lib.check(_lib.lib().sbb_predict_full(self._handle(), _ptr(img_tile)[0], _ptr(out)[0], _lib.SBB_MEM_DEVICE,
                                               C.c_void_p(stream or 1)))
        return out

    # -- introspection -------------------------------------------------------------------------
    def activations(self):
        l, out = _lib.lib(), []
        for i in range(l.sbb_model_num_activations(self._handle())):
            name = C.c_char_p()
            h, w, c = C.c_int32(), C.c_int32(), C.c_int32()
            _lib.check(l.sbb_model_activation_info(self._handle(), i, C.byref(name), C.byref(h), C.byref(w), C.byref(c)))
            out.append((name.value.decode(), h.value, w.value, c.value))
        return out

    @_serialised
    def read_activation(self, index: int, tile: int = 0):
        name, h, w, c = self.activations()[index]
        out = np.empty((h, w, c), np.float32)
        _lib.check(_lib.lib().sbb_model_read_activation(self._handle(), index, tile, _ptr(out)[0]))
        return out

    def set_profiling(self, on):
        """True / 1: event pair around every launch (``layer_times``); 2: encoder / decoder split of the undisturbed
        forward from three events (``part_times``); False / 0: off."""
        _lib.check(_lib.lib().sbb_model_set_profiling(self._handle(), int(on)))

    def part_times(self, reset: bool = True):
        """(encoder ms, decoder ms, forwards) summed over the forwards since the last reset (profiling mode 2)."""
        e, d, n = C.c_float(), C.c_float(), C.c_int32()
        _lib.check(_lib.lib().sbb_model_part_times(self._handle(), C.byref(e), C.byref(d), C.byref(n), 1 if reset else 0))
        return e.value, d.value, n.value

    def layer_times(self):
        l, out = _lib.lib(), []
        for i in range(l.sbb_model_num_layers(self._handle())):
            name, ms, fl = C.c_char_p(), C.c_float(), C.c_double()
            _lib.check(l.sbb_model_layer_time(self._handle(), i, C.byref(name), C.byref(ms), C.byref(fl)))
            out.append((name.value.decode(), ms.value, fl.value))
        return out

    @_serialised
    def set_precision_plan(self, layers=()):
        """Layers (names as in ``layer_times``) that read only the hi plane of their activations; () resets."""
        _lib.check(_lib.lib().sbb_model_set_precision_plan(self._handle(), ",".join(layers).encode()))
        self.precision_plan = tuple(layers)

    def geom_cache_stats(self):
        """(hits, misses) of the handle's page-geometry cache."""
        h, ms = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().sbb_model_geom_cache_stats(self._handle(), C.byref(h), C.byref(ms)))
        return h.value, ms.value

    def last_launch_count(self) -> int:
        return int(_lib.lib().sbb_model_last_launch_count(self._handle()))


def compute_tile_grid(H: int, Wd: int, tile_h: int, tile_w: int, margin: int = -1):
    """Host-only: (nxf, nyf, tile_org[n,4] = x0,y0,i,j in reference loop order, owner_x[W], owner_y[H])."""
    l = _lib.lib()
    nx, ny = C.c_int32(), C.c_int32()
    _lib.check(l.sbb_compute_tile_grid(H, Wd, tile_h, tile_w, margin, C.byref(nx), C.byref(ny), None, 0, None, None))
    org = np.zeros((nx.value * ny.value, 4), np.int32)
    ox, oy = np.zeros(Wd, np.int16), np.zeros(H, np.int16)
    _lib.check(l.sbb_compute_tile_grid(H, Wd, tile_h, tile_w, margin, C.byref(nx), C.byref(ny), _ptr(org)[0],
                                       org.shape[0], _ptr(ox)[0], _ptr(oy)[0]))
    return nx.value, ny.value, org, ox, oy
