"""Keras .h5 import: the bundled HDF5 reader against a REAL HDF5-library file
(scipy ships one MATLAB v7.3 fixture), and the Keras layout -> weight dict mapping through a test-side
writer that emits the structures h5py 2.x produces (symbol-table groups, contiguous float32 datasets,
fixed-length string array attributes)."""
import json
import os

import numpy as np
import pytest

from h5write_min import Writer, write_keras_model
from sbb_textline_detection_b200 import h5lite, keras_h5, weights
from sbb_textline_detection_b200.arch import conv_specs


def test_reader_on_a_file_written_by_the_hdf5_library():
    import scipy.io
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy test data not installed")
    f = h5lite.File(p)            # 512-byte user block, superblock v0, B-tree/heap/SNOD group
    assert f.base == 512 and f.root.keys() == ["testdouble"]
    d = f.root["testdouble"]
    assert h5lite._to_str(d.attrs["MATLAB_class"]) == "double"
    assert d.shape == (9, 1)
    np.testing.assert_allclose(d.read()[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)


def test_writer_reader_roundtrip_types_and_many_links():
    w = Writer(user_block=512)
    kids = {}
    arrays = {}
    rng = np.random.default_rng(0)
    for i in range(150):  # > one SNOD, exercises the B-tree walk
        a = rng.standard_normal((3, i % 5 + 1)).astype(np.float32 if i % 2 else np.float64)
        arrays[f"d{i:03d}"] = a
        kids[f"d{i:03d}"] = w.dataset(a, {"idx": np.int32(i)})
    kids["ints"] = w.dataset(np.arange(12, dtype=np.int64).reshape(3, 4))
    sub = w.group(kids, {"names": np.array([b"alpha", b"be"], dtype="S8"), "scalar": np.bytes_(b"hello")})
    blob = w.finish(w.group({"sub": sub[0]}))
    f = h5lite.File(blob)
    g = f.root["sub"]
    assert sorted(g.keys()) == sorted(kids)
    assert h5lite.attr_strings(g.attrs["names"]) == ["alpha", "be"]
    assert h5lite._to_str(g.attrs["scalar"]) == "hello"
    for k, a in arrays.items():
        got = f.root["sub/" + k]
        assert got.read().dtype == a.dtype and (got.read() == a).all()
        assert int(got.attrs["idx"]) == int(k[1:])
    assert (f.root["sub/ints"].read() == np.arange(12).reshape(3, 4)).all()
    with pytest.raises(KeyError):
        f.root["sub/nope"]
    with pytest.raises(h5lite.H5Error):
        h5lite.File(b"not an hdf5 file at all" * 100)


@pytest.mark.parametrize("gzip,shuffle", [(True, True), (True, False), (False, False)])
def test_chunked_compressed_datasets(gzip, shuffle):
    """Models re-saved with h5py compression: chunked layout, shuffle + deflate filters, ragged edge chunks."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((3, 3, 37, 50)).astype(np.float32)
    b = rng.integers(-1000, 1000, (45, 7)).astype(np.int64)
    w = Writer()
    root = w.group({"kernel": w.chunked_dataset(a, (3, 3, 16, 32), gzip, shuffle),
                    "ints": w.chunked_dataset(b, (10, 4), gzip, shuffle)})
    f = h5lite.File(w.finish(root))
    assert f.root["kernel"].shape == a.shape and (f.root["kernel"].read() == a).all()
    assert (f.root["ints"].read() == b).all()


def _keras_layers(w, n_classes, offset=0):
    """Arrange a weight dict the way Keras saves the model: named encoder layers, un-named decoder
    layers with a process-global counter (offset simulates a second model built in the same process)."""
    by_layer, order = {}, ["input_1", "zero_padding2d_1"]
    dec = ("dec_v5", "dec_v4", "dec1", "dec2", "dec3", "dec4", "dec5", "cls")
    for s in conv_specs(n_classes):
        if s.name in dec:
            k = dec.index(s.name) + 1 + offset
            cn, bn = f"conv2d_{k}", f"batch_normalization_{k}"
        else:
            cn, bn = s.name, s.bn
        by_layer[cn] = [("kernel:0", w[s.name + "/kernel"]), ("bias:0", w[s.name + "/bias"])]
        by_layer[bn] = [("gamma:0", w[s.bn + "/gamma"]), ("beta:0", w[s.bn + "/beta"]),
                        ("moving_mean:0", w[s.bn + "/mean"]), ("moving_variance:0", w[s.bn + "/var"])]
        order += [cn, bn, f"activation_{len(order)}"]
    return by_layer, order


@pytest.mark.parametrize("n_classes,offset,wrap", [(2, 0, True), (4, 8, True), (2, 3, False)])
def test_keras_layout_import(tmp_path, n_classes, offset, wrap):
    w = weights.random_init(77, n_classes)
    rng = np.random.default_rng(1)
    for k in w:
        if k.endswith("/mean"):
            w[k] = rng.standard_normal(w[k].shape).astype(np.float32)
        if k.endswith("/var"):
            w[k] = rng.uniform(0.5, 2, w[k].shape).astype(np.float32)
    by_layer, order = _keras_layers(w, n_classes, offset)
    cfg = json.dumps({"class_name": "Model", "config": {"layers": [
        {"class_name": "InputLayer", "config": {"batch_input_shape": [None, 448, 448, 3], "name": "input_1"}}]}})
    p = tmp_path / "model_textline_new.h5"
    p.write_bytes(write_keras_model(by_layer, order, cfg, wrap=wrap))
    got, nc, tile = keras_h5.read_keras_h5(str(p))
    assert nc == n_classes and tile == ((448, 448) if wrap else None)
    assert set(got) == set(w)
    for k in w:
        assert got[k].dtype == np.float32 and (got[k] == w[k]).all(), k
    # and the packed blob (what sbb_model_create consumes) is identical to packing the original dict
    assert weights.pack_blob(got, nc) == weights.pack_blob(w, n_classes)


def test_keras_import_rejects_other_architectures(tmp_path):
    w = weights.random_init(5, 2)
    by_layer, order = _keras_layers(w, 2)
    by_layer["conv2d_3"][0] = ("kernel:0", np.zeros((3, 3, 1024, 256), np.float32))  # dec1 with dec2's shape
    p = tmp_path / "m.h5"
    p.write_bytes(write_keras_model(by_layer, order))
    with pytest.raises(ValueError, match="dec1"):
        keras_h5.read_keras_h5(str(p))


def test_converter_cli(tmp_path):
    w = weights.random_init(9, 2)
    by_layer, order = _keras_layers(w, 2)
    p = tmp_path / "m.h5"
    p.write_bytes(write_keras_model(by_layer, order))
    # no model_config in the file: the converter refuses to guess the input size ...
    assert keras_h5.main([str(p)]) == 2 and not (tmp_path / "m.sbbw").exists()
    # ... and records the one it is given in the blob header
    assert keras_h5.main([str(p), str(tmp_path / "m.sbbw"), "672"]) == 0
    blob = (tmp_path / "m.sbbw").read_bytes()
    assert blob == weights.pack_blob(w, 2, 672) and weights.blob_tile(blob) == (672, 672)
    cfg = json.dumps({"class_name": "Model", "config": {"layers": [
        {"class_name": "InputLayer", "config": {"batch_input_shape": [None, 96, 128, 3], "name": "input_1"}}]}})
    q = tmp_path / "n.h5"
    q.write_bytes(write_keras_model(by_layer, order, cfg))
    assert keras_h5.main([str(q)]) == 0
    assert weights.blob_tile((tmp_path / "n.sbbw").read_bytes()) == (96, 128)


def test_model_input_size_is_never_guessed(tmp_path, monkeypatch):
    """a model whose input is not 448x448 must not silently run on a 448 tile grid
    (main.py:227-233 reads the size from the model).  load_model_file takes it from the blob header / the
    file's model_config, checks an explicit ``tile`` against it, and raises when there is neither."""
    from sbb_textline_detection_b200 import detector as D
    seen = {}

    class FakeModel:
        def __init__(self, weights_, th, tw, nc, **kw):
            seen["tile"] = (th, tw)
    monkeypatch.setattr(D, "SbbModel", FakeModel)
    w = weights.random_init(9, 2)
    h5 = tmp_path / "model_textline_new.h5"
    (tmp_path / "model_textline_new.sbbw").write_bytes(weights.pack_blob(w, 2, 672))
    D.load_model_file(str(h5))
    assert seen["tile"] == (672, 672)
    D.load_model_file(str(h5), tile=672)
    with pytest.raises(ValueError, match="672x672"):
        D.load_model_file(str(h5), tile=448)
    (tmp_path / "model_textline_new.sbbw").write_bytes(weights.pack_blob(w, 2))     # header without a size
    with pytest.raises(ValueError, match="does not record"):
        D.load_model_file(str(h5))
    D.load_model_file(str(h5), tile=96)
    assert seen["tile"] == (96, 96)
    (tmp_path / "model_textline_new.sbbw").unlink()
    by_layer, order = _keras_layers(w, 2)
    h5.write_bytes(write_keras_model(by_layer, order))                              # .h5 without model_config
    with pytest.raises(ValueError, match="does not record"):
        D.load_model_file(str(h5))


@pytest.mark.gpu
def test_h5_model_file_drops_into_the_detector(built_lib, tmp_path):
    """A Keras-layout .h5 under the reference's hard-coded file name (main.py:60) loads through
    start_new_session_and_model and predicts exactly like the same weights handed over directly."""
    from sbb_textline_detection_b200 import detector as D, synth
    from sbb_textline_detection_b200.model import SbbModel
    w, nc = D.synthetic_weights("textline")
    by_layer, order = _keras_layers(w, nc, offset=16)
    cfg = json.dumps({"class_name": "Model", "config": {"layers": [
        {"class_name": "InputLayer", "config": {"batch_input_shape": [None, 96, 96, 3], "name": "input_1"}}]}})
    (tmp_path / "model_textline_new.h5").write_bytes(write_keras_model(by_layer, order, cfg))
    det = D.textline_detector(str(tmp_path / "x.png"), str(tmp_path), "x", str(tmp_path), cache_models=False, max_batch=16)
    model, session = det.start_new_session_and_model(det.model_textline_dir)
    assert model.layers[-1].output_shape == (None, 96, 96, nc)
    page = synth.document_page(300, 260, seed=21)
    got = det.do_prediction(True, page, model)
    session.close()
    m = SbbModel(w, 96, 96, nc, max_batch=16)
    want = m.predict_page(page)
    m.close()
    assert (got[:, :, 0] == want).all()
