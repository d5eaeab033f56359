"""Weights: seeded random init (there are no .h5 files and no network here), BatchNorm folding and
the packed blob that ``sbb_model_create`` (include/sbb_textline.h) consumes.

The dict layout mirrors the Keras .h5 groups the reference loads (main.py:221):
``<conv>/kernel`` HWIO float32, ``<conv>/bias``, ``<bn>/gamma|beta|mean|var``.

Blob layout (little endian), version SBBW0001:
    char  magic[8] = "SBBW0001"
    u32   n_classes, n_records, tile_h, tile_w   # model input size the weights were trained for (0, 0 = not
                                                 # recorded); sbb_model_create refuses a different tile size
    per record:
        char name[32]; u32 kh, kw, cin, cout; u64 n_weights
        f32  weights[n_weights]   # OHWI: [cout][kh][kw][cin], BatchNorm scale already folded in
        f32  bias[cout]           # (conv_bias - mean) * scale + beta
    Record "bn_conv1" (kh = kw = 0, cin = 0) carries the stem BatchNorm as weights = scale[64],
    bias = shift[64]: conv1's output must stay raw because the decoder's f1 skip taps it before BN.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from .arch import BN_EPS, conv_specs

MAGIC = b"SBBW0001"


def _layer_rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def random_init(seed: int, n_classes: int) -> dict:
    """He-normal kernels, N(0, 0.01) bias, gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1), BN moving stats
    at the Keras defaults (mean 0, var 1) until ``apply_bn_stats`` installs calibrated ones.
    Per-layer generators are keyed by (seed, layer name), so models that differ only in n_classes
    share every layer except ``cls``."""
    w = {}
    for s in conv_specs(n_classes):
        rng = _layer_rng(seed, s.name)
        fan_in = s.kh * s.kw * s.cin
        w[s.name + "/kernel"] = (rng.standard_normal((s.kh, s.kw, s.cin, s.cout)) *
                                 np.sqrt(2.0 / fan_in)).astype(np.float32)
        w[s.name + "/bias"] = (rng.standard_normal(s.cout) * 0.01).astype(np.float32)
        w[s.bn + "/gamma"] = rng.uniform(0.5, 1.5, s.cout).astype(np.float32)
        w[s.bn + "/beta"] = (rng.standard_normal(s.cout) * 0.1).astype(np.float32)
        w[s.bn + "/mean"] = np.zeros(s.cout, np.float32)
        w[s.bn + "/var"] = np.ones(s.cout, np.float32)
    return w


def apply_bn_stats(w: dict, stats) -> dict:
    """Install calibrated BatchNorm moving statistics (keys ``<bn>/mean``, ``<bn>/var``)."""
    out = dict(w)
    for k in stats.keys() if hasattr(stats, "keys") else stats.files:
        if k.endswith("/mean") or k.endswith("/var"):
            assert k in out and out[k].shape == stats[k].shape, k
            out[k] = np.asarray(stats[k], np.float32)
    return out


def fold_bn(w: dict, n_classes: int):
    """Returns list of (name, kh, kw, cin, cout, weights OHWI f32, bias f32) with BN folded
    (fp64 arithmetic, one rounding to fp32), plus the bn_conv1 (scale, shift) record."""
    recs = []
    for s in conv_specs(n_classes):
        k = w[s.name + "/kernel"].astype(np.float64)
        b = w[s.name + "/bias"].astype(np.float64)
        scale = w[s.bn + "/gamma"].astype(np.float64) / np.sqrt(w[s.bn + "/var"].astype(np.float64) + BN_EPS)
        shift = w[s.bn + "/beta"].astype(np.float64) - w[s.bn + "/mean"].astype(np.float64) * scale
        ohwi = np.transpose(k, (3, 0, 1, 2))  # HWIO -> OHWI
        if s.name == "conv1":
            recs.append((s.name, s.kh, s.kw, s.cin, s.cout,
                         np.ascontiguousarray(ohwi, np.float32), b.astype(np.float32)))
            recs.append(("bn_conv1", 0, 0, 0, s.cout, scale.astype(np.float32), shift.astype(np.float32)))
        else:
            wf = ohwi * scale[:, None, None, None]
            bf = b * scale + shift
            recs.append((s.name, s.kh, s.kw, s.cin, s.cout,
                         np.ascontiguousarray(wf, np.float32), bf.astype(np.float32)))
    return recs


def pack_blob(w: dict, n_classes: int, tile=None) -> bytes:
    """``tile``: (tile_h, tile_w) or one int -- the model's input size (main.py:227-228), recorded in the header
    so that a converted ``.sbbw`` cannot be run with another tile grid than the reference would use."""
    recs = fold_bn(w, n_classes)
    th, tw = (0, 0) if tile is None else ((int(tile), int(tile)) if np.isscalar(tile) else (int(tile[0]), int(tile[1])))
    parts = [MAGIC, struct.pack("<IIII", n_classes, len(recs), th, tw)]
    for name, kh, kw, cin, cout, wt, bias in recs:
        nm = name.encode()
        assert len(nm) < 32
        parts.append(nm.ljust(32, b"\0"))
        parts.append(struct.pack("<IIIIQ", kh, kw, cin, cout, wt.size))
        parts.append(wt.tobytes())
        parts.append(bias.tobytes())
    return b"".join(parts)


def blob_tile(blob: bytes):
    """(tile_h, tile_w) recorded in a blob header, or None when it does not record one."""
    assert blob[:8] == MAGIC, "bad magic"
    _, _, th, tw = struct.unpack_from("<IIII", blob, 8)
    return (th, tw) if th and tw else None


def unpack_blob(blob: bytes):
    """Inverse of pack_blob (host-side check / tooling)."""
    assert blob[:8] == MAGIC, "bad magic"
    n_classes, n_rec, _, _ = struct.unpack_from("<IIII", blob, 8)
    off = 24
    recs = []
    for _ in range(n_rec):
        name = blob[off:off + 32].split(b"\0", 1)[0].decode()
        kh, kw, cin, cout, nw = struct.unpack_from("<IIIIQ", blob, off + 32)
        off += 32 + 24
        wt = np.frombuffer(blob, np.float32, nw, off)
        off += 4 * nw
        bias = np.frombuffer(blob, np.float32, cout, off)
        off += 4 * cout
        recs.append((name, kh, kw, cin, cout, wt, bias))
    assert off == len(blob)
    return n_classes, recs
