"""Minimal reader for Keras ``save_weights`` HDF5 files (a "next" row).  Placeholder:
the reference's trained weight file is not shipped (.MISSING_LARGE_BLOBS) and h5py is not installed, so
until the reader lands HDF5 input is accepted only when h5py is importable."""


def read_keras_weights(path, names_in_layer_order):
    try:
        import h5py  # noqa: PLC0415
    except ImportError as exc:
        raise ImportError("reading Keras HDF5 weights needs h5py here; convert to .npz keyed by Keras "
                          "weight names (ChessModel.save writes that)") from exc
    out = {}
    with h5py.File(path, "r") as f:
        root = f["model_weights"] if "model_weights" in f else f
        for lname in [n.decode() if isinstance(n, bytes) else n for n in root.attrs["layer_names"]]:
            grp = root[lname]
            for wname in [n.decode() if isinstance(n, bytes) else n for n in grp.attrs["weight_names"]]:
                out[wname] = grp[wname][()]
    return out
