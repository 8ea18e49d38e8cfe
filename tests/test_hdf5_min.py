"""Dependency-free HDF5 reader/writer for Keras weight files: round trips and format details (CPU only).
No libhdf5 exists in this environment, so the reader is pinned against the writer and against hand-checked
structure facts of the HDF5 File Format Specification, not against h5py output."""
import struct

import numpy as np
import pytest

from chess_alpha_zero_b200 import hdf5_min, keras_graph
from oracle import weights as oweights


def test_weight_file_roundtrip_full_network(tmp_path):
    cfg = keras_graph.build_config()                                  # 7 x 256: 68 layers in the root group
    w = oweights.synth_weights(cfg, seed=3, recipe="randomized")
    path = str(tmp_path / "model_best_weight.h5")
    keras_graph.save_weight_file(path, w, cfg)
    raw = open(path, "rb").read()
    assert raw[:8] == hdf5_min.SIGNATURE and struct.unpack_from("<Q", raw, 40)[0] == len(raw)   # superblock EOF address
    back = keras_graph.load_weight_file(path, cfg)
    assert list(back) == list(keras_graph.expected_shapes(cfg))
    for k, v in w.items():
        assert back[k].dtype == np.float32 and np.array_equal(back[k], v), k
    f = hdf5_min.H5File(path)
    attrs = f.attrs(f.root)
    assert attrs["layer_names"] == [l["name"] for l in cfg["layers"]] and attrs["keras_version"] == ["2.0.8"]
    kids = f.children(f.root)
    assert len(kids) == len(cfg["layers"]) > 2 * hdf5_min._Writer.LEAF_K          # several symbol nodes under the B-tree
    grp = f.resolve("res3_batchnorm2")
    assert f.attrs(grp)["weight_names"] == [f"res3_batchnorm2/{p}" for p in keras_graph.BN_PARTS]
    assert f.attrs(f.resolve("res3_relu1")).get("weight_names") == []               # weightless layers keep a group


def test_model_save_layout_and_errors(tmp_path):
    # 'model.save' files keep the same tree under /model_weights
    w = hdf5_min._Writer()
    d = w.dataset(np.arange(6, dtype=np.float32).reshape(2, 3))
    inner = w.group({"kernel:0": d})[0]
    layer = w.group({"dense_1": inner}, {"weight_names": ["dense_1/kernel:0"]})[0]
    mw = w.group({"dense_1": layer}, {"layer_names": ["dense_1"]})[0]
    root, t, hp = w.group({"model_weights": mw})
    path = str(tmp_path / "full.h5")
    open(path, "wb").write(w.finish(root, t, hp))
    got = hdf5_min.read_keras_weights(path)
    assert list(got) == ["dense_1/kernel:0"] and np.array_equal(got["dense_1/kernel:0"], np.arange(6).reshape(2, 3))
    with pytest.raises(KeyError):
        hdf5_min.H5File(path).resolve("model_weights/nope")
    bad = str(tmp_path / "bad.h5")
    open(bad, "wb").write(b"not hdf5 at all" * 10)
    with pytest.raises(hdf5_min.Hdf5Error):
        hdf5_min.H5File(bad)
    with pytest.raises(keras_graph.GraphError):
        keras_graph.load_weight_file(bad, keras_graph.build_config(blocks=1, filters=64))


def test_reader_handles_variable_length_strings_and_v2_dataspace(tmp_path):
    """Newer Keras/h5py write attribute strings as variable-length (global heap) and version-2 dataspaces."""
    w = hdf5_min._Writer()
    names = ["dense_1", "a_much_longer_layer_name"]
    # global heap collection holding the two strings
    objs = b""
    for i, s in enumerate(names, 1):
        e = s.encode()
        objs += struct.pack("<HHIQ", i, 1, 0, len(e)) + e + b"\0" * (hdf5_min._pad8(len(e)) - len(e))
    objs += struct.pack("<HHIQ", 0, 0, 0, 0)
    coll = w._alloc(b"GCOL" + struct.pack("<B3xQ", 1, 16 + len(objs)) + objs)
    vlen_dt = struct.pack("<BBBBI", 0x19, 0x01, 0x00, 0x00, 16) + hdf5_min._Writer._dtype_str(1)
    ds_v2 = struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 2)
    data = b"".join(struct.pack("<IQI", len(s), coll, i) for i, s in enumerate(names, 1))
    nm = b"layer_names\0"
    attr = struct.pack("<BxHHHB", 3, len(nm), len(vlen_dt), len(ds_v2), 0) + nm + vlen_dt + ds_v2 + data
    root_hdr, t, hp = w.group({})
    # append the attribute by rewriting the root header with it
    msgs = [w._msg(0x0011, struct.pack("<QQ", t, hp)), w._msg(0x000C, attr)]
    root = w._object_header(msgs)
    path = str(tmp_path / "vlen.h5")
    open(path, "wb").write(w.finish(root, t, hp))
    f = hdf5_min.H5File(path)
    assert f.attrs(f.root)["layer_names"] == names
