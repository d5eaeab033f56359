"""Import of the Keras 2.3 ``.h5`` model files the reference hard-codes (``model_page_mixed_best.h5``,
``model_strukturerkennung.h5``, ``model_textline_new.h5``; main.py:58-60, loaded by
``load_model(model_dir, compile=False)`` main.py:221) into the flat weight dict of ``weights.py``
(SURVEY.md section 8(f) rank 2).  Reads the file with the bundled pure-Python HDF5 parser (h5lite).

Keras layout:  [/model_weights]/<layer>/<layer>/{kernel:0,bias:0,gamma:0,beta:0,moving_mean:0,
moving_variance:0}; group attrs ``layer_names`` / ``weight_names``; root attr ``model_config`` (JSON).
The encoder layers carry the Keras-ResNet50 names (conv1, bn_conv1, res2a_branch2a, bn2a_branch2a, ...);
the decoder / classifier layers are un-named in the training code, so they appear as ``conv2d_<k>`` /
``batch_normalization_<k>`` with a process-global counter: convs are mapped by their (unique) kernel
shapes, BatchNorms by creation order (numeric suffix), both validated against the expected channels.
``model_config`` is only read for the input size; the marshalled ``Lambda`` in it is never executed.
"""
from __future__ import annotations

import json
import re

import numpy as np

from . import h5lite
from .arch import conv_specs

_DEC = ("dec_v5", "dec_v4", "dec1", "dec2", "dec3", "dec4", "dec5", "cls")


def _layer_weights(g):
    """-> {layer_name: {short weight name: array}} for every layer group that holds weights."""
    names = h5lite.attr_strings(g.attrs.get("layer_names"))
    k = 0
    while f"layer_names{k}" in g.attrs:   # Keras splits attributes that outgrow the 64 KB object-header limit
        names += h5lite.attr_strings(g.attrs[f"layer_names{k}"])
        k += 1
    names = names or g.keys()
    out = {}
    for ln in names:
        if ln not in g:
            continue
        lg = g[ln]
        if not isinstance(lg, h5lite.Group):
            continue
        ws = {}
        for path, ds in lg.visit_datasets():
            short = path.split("/")[-1].split(":")[0]
            ws[short] = np.asarray(ds.read())
        if ws:
            out[ln] = ws
    return out


def _suffix(name):
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def read_keras_h5(path):
    """-> (weights dict, n_classes, (tile_h, tile_w) or None).  Raises ValueError with the offending
    layer when the file does not hold the ResNet50-U-Net of SURVEY.md Appendix A."""
    f = h5lite.File(path)
    root = f.root
    g = root["model_weights"] if "model_weights" in root else root
    layers = _layer_weights(g)
    convs = {k: v for k, v in layers.items() if "kernel" in v and v["kernel"].ndim == 4}
    bns = {k: v for k, v in layers.items() if "gamma" in v and "moving_mean" in v}

    anon_convs = sorted((k for k in convs if re.fullmatch(r"conv2d(_\d+)?", k)), key=_suffix)
    anon_bns = sorted((k for k in bns if re.fullmatch(r"batch_normalization(_\d+)?", k)), key=_suffix)
    if not anon_convs:
        raise ValueError("no un-named conv2d_<k> layers: not a resnet50_unet .h5")
    n_classes = int(convs[anon_convs[-1]]["kernel"].shape[3])
    specs = conv_specs(n_classes)
    by_shape = {}
    for k in anon_convs:
        by_shape.setdefault(tuple(convs[k]["kernel"].shape), []).append(k)
    if len(anon_bns) != len(_DEC):
        raise ValueError(f"expected {len(_DEC)} un-named BatchNormalization layers, found {len(anon_bns)}")
    dec_bn = dict(zip(_DEC, anon_bns))

    w = {}
    for s in specs:
        shape = (s.kh, s.kw, s.cin, s.cout)
        if s.name in _DEC:
            cands = by_shape.get(shape, [])
            if len(cands) != 1:
                raise ValueError(f"{s.name}: expected exactly one conv2d_<k> with kernel {shape}, found {cands}")
            ck, bk = cands[0], dec_bn[s.name]
        else:
            ck, bk = s.name, s.bn
        if ck not in convs:
            raise ValueError(f"layer {ck} missing from {path}")
        if bk not in bns:
            raise ValueError(f"layer {bk} missing from {path}")
        c, b = convs[ck], bns[bk]
        if tuple(c["kernel"].shape) != shape:
            raise ValueError(f"{ck}: kernel {c['kernel'].shape}, expected {shape}")
        if b["gamma"].shape != (s.cout,):
            raise ValueError(f"{bk}: {b['gamma'].shape[0]} channels, expected {s.cout} (after {ck})")
        w[s.name + "/kernel"] = np.ascontiguousarray(c["kernel"], np.float32)
        w[s.name + "/bias"] = np.ascontiguousarray(c.get("bias", np.zeros(s.cout)), np.float32)
        w[s.bn + "/gamma"] = np.ascontiguousarray(b["gamma"], np.float32)
        w[s.bn + "/beta"] = np.ascontiguousarray(b.get("beta", np.zeros(s.cout)), np.float32)
        w[s.bn + "/mean"] = np.ascontiguousarray(b["moving_mean"], np.float32)
        w[s.bn + "/var"] = np.ascontiguousarray(b["moving_variance"], np.float32)

    tile = None
    cfg = root.attrs.get("model_config")
    if cfg is not None:
        try:
            cfg = json.loads(h5lite._to_str(cfg))
            for layer in cfg["config"]["layers"]:
                shp = layer.get("config", {}).get("batch_input_shape")
                if shp:
                    tile = (int(shp[1]), int(shp[2])) if shp[1] and shp[2] else None
                    break
        except Exception:
            tile = None
    elif getattr(root, "has_dense_attrs", False):
        import warnings
        warnings.warn(f"{path}: the root group keeps attributes in dense storage, which the bundled HDF5 reader does "
                      "not parse -- model_config (the input size) is unavailable; pass the tile size explicitly")
    return w, n_classes, tile


def main(argv=None):
    """python -m sbb_textline_detection_b200.keras_h5 model.h5 [model.sbbw [tile]]: convert once, load fast.
    The blob records the model's input size (from ``model_config``, or ``tile`` when the file has none)."""
    import sys
    from . import weights
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        print("usage: python -m sbb_textline_detection_b200.keras_h5 model.h5 [model.sbbw [tile]]")
        return 2
    src = argv[0]
    dst = argv[1] if len(argv) > 1 else src.rsplit(".", 1)[0] + ".sbbw"
    w, nc, tile = read_keras_h5(src)
    if tile is None and len(argv) > 2:
        tile = (int(argv[2]), int(argv[2]))
    if tile is None:
        print(f"{src}: no usable model_config (input size); re-run with the tile size as third argument")
        return 2
    with open(dst, "wb") as f:
        f.write(weights.pack_blob(w, nc, tile))
    print(f"{src}: {nc} classes, input {tile} -> {dst}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
