"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), the Keras-JSON parser, the label space, the pipe server."""
import ctypes
import json
import os
import re
import threading

import numpy as np
import pytest

import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import _lib, keras_graph
from oracle import labels as olabels
from oracle import weights as oweights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIPPED = "/root/reference/data/model/model_best_config.json"


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "caz_b200.h")).read()
    declared = set(re.findall(r"\b(caz_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.library_path())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert _lib.load().caz_abi_version() == 1


def test_arch_struct_matches_header_layout():
    assert ctypes.sizeof(_lib.Arch) == 40 and ctypes.sizeof(_lib.Tensor) == 16


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = keras_graph.build_config(blocks=1, filters=64)
    with pytest.raises(caz.CazError) as e:
        caz.Evaluator(cfg, keras_graph.init_weights(cfg, 0))
    assert e.value.code == 3 and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chess-alpha-zero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_labels_match_oracle_and_reference_hashes():
    import hashlib
    assert caz.labels.LABELS == olabels.LABELS and caz.labels.N_LABELS == 1968
    assert np.array_equal(caz.labels.UNFLIPPED_INDEX, olabels.UNFLIPPED_INDEX)
    assert hashlib.sha256(",".join(caz.labels.LABELS).encode()).hexdigest().startswith("409054c3")
    p = np.random.RandomState(0).rand(1968).astype(np.float32)
    assert np.array_equal(caz.Config.flip_policy(p), olabels.flip_policy(p))
    assert caz.labels.legal_mask(["e2e4", "a7a8q"]).nonzero()[0].tolist() == [930, 1793]


@pytest.mark.parametrize("variant", [dict(), dict(stem_k=3, value_filters=1), dict(blocks=19), dict(filters=64, blocks=2)])
def test_config_roundtrip_and_parse(variant):
    cfg = keras_graph.build_config(**variant)
    assert cfg == oweights.keras_config(**variant)
    arch, names = keras_graph.parse_config(cfg)
    assert (arch.stem_kernel, arch.filters, arch.res_blocks, arch.value_filters) == (
        variant.get("stem_k", 5), variant.get("filters", 256), variant.get("blocks", 7), variant.get("value_filters", 4))
    assert arch.n_labels == 1968 and abs(arch.bn_epsilon - 1e-3) < 1e-9 and arch.policy_filters == 2
    assert len(names) == 5 + 10 * arch.res_blocks + 16
    ow = oweights.synth_weights(cfg, seed=1)
    assert set(names) == set(ow) == set(keras_graph.expected_shapes(cfg))
    for k, shape in keras_graph.expected_shapes(cfg).items():
        assert tuple(ow[k].shape) == tuple(shape), k
    assert keras_graph.keras_layer_order(cfg) == list(ow)          # Keras layer-list order
    init = keras_graph.init_weights(cfg, 0)
    assert sum(v.size for v in init.values()) == sum(v.size for v in ow.values())


@pytest.mark.skipif(not os.path.exists(SHIPPED), reason="reference checkout absent (GPU box)")
def test_parses_the_json_the_reference_ships():
    cfg = json.load(open(SHIPPED))
    arch, names = keras_graph.parse_config(cfg)
    assert (arch.stem_kernel, arch.filters, arch.res_blocks, arch.value_filters, arch.value_fc) == (3, 256, 7, 1, 256)
    assert names[0] == "input_conv-3-256/kernel:0" and names[-1] == "value_out/bias:0"


def test_parser_rejects_foreign_graphs():
    cfg = keras_graph.build_config(blocks=1, filters=64)
    bad = json.loads(json.dumps(cfg))
    bad["layers"][1]["config"]["use_bias"] = True
    with pytest.raises(keras_graph.GraphError):
        keras_graph.parse_config(bad)
    bad = json.loads(json.dumps(cfg))
    bad["layers"][1]["config"]["data_format"] = "channels_last"
    with pytest.raises(keras_graph.GraphError):
        keras_graph.parse_config(bad)


def test_weight_file_roundtrip(tmp_path):
    cfg = keras_graph.build_config(blocks=1, filters=64)
    w = oweights.synth_weights(cfg, seed=2, recipe="randomized")
    path = str(tmp_path / "w.npz")
    keras_graph.save_weight_file(path, w)
    back = keras_graph.load_weight_file(path, cfg)
    assert list(back) == list(keras_graph.expected_shapes(cfg))
    for k in w:
        assert np.array_equal(back[k], w[k])
    del w["value_out/bias:0"]
    keras_graph.save_weight_file(path, w)
    with pytest.raises(keras_graph.GraphError):
        keras_graph.load_weight_file(path, cfg)


def test_fen_records_match_oracle(golden_dir):
    from oracle import encoder as oenc
    fens = json.load(open(os.path.join(golden_dir, "encoder_fens.json")))["fens"]
    recs = caz.encoder.fens_to_records(fens)
    for fen, r in zip(fens, recs):
        assert np.array_equal(r, oenc.fen_to_record(fen)), fen


class _FakeKeras:
    """predict_on_batch with a checkable answer: policy row = one-hot of the plane sum, value = mean."""

    def predict_on_batch(self, data):
        assert data.dtype == np.float32 and data.shape[1:] == (18, 8, 8)
        pol = np.zeros((len(data), 1968), np.float32)
        idx = data.reshape(len(data), -1).sum(axis=1).astype(int) % 1968
        pol[np.arange(len(data)), idx] = 1
        return [pol, data.reshape(len(data), -1).mean(axis=1, keepdims=True)]


class _FakeModel:
    model = _FakeKeras()


def _client(pipe, k, out):
    for j in range(20):
        x = np.full((18, 8, 8), k * 20 + j, np.float32)
        pipe.send(x)
        p, v = pipe.recv()
        out.append((int(p.argmax()) == int(x.sum()) % 1968, abs(v - (k * 20 + j)) < 1e-6, isinstance(v, float)))


def _process_client(pipe, k, queue):
    out = []
    _client(pipe, k, out)
    queue.put(out)


@pytest.mark.parametrize("transport", ["pipe", "shm"])
def test_pipe_protocol_across_processes(transport):
    """The endpoint objects travel to spawned workers the way the reference hands Connections to its pool
    (worker/self_play.py:47,122-129)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    api = caz.ChessModelAPI(_FakeModel(), transport=transport)
    pipes = [api.create_pipe() for _ in range(3)]
    api.start()
    queue = ctx.Queue()
    procs = [ctx.Process(target=_process_client, args=(p, k, queue)) for k, p in enumerate(pipes)]
    [p.start() for p in procs]
    out = [r for _ in procs for r in queue.get(timeout=120)]
    [p.join(timeout=30) for p in procs]
    api.stop()
    assert len(out) == 60 and all(all(r) for r in out)


def test_shm_slab_capacity_is_enforced():
    api = caz.ChessModelAPI(_FakeModel(), transport="shm", shm_capacity=2)
    api.create_pipe(), api.create_pipe()
    with pytest.raises(RuntimeError, match="shm_capacity"):
        api.create_pipe()
    api.stop()


@pytest.mark.parametrize("which", ["ours", "ours-shm", "reference"])
def test_pipe_protocol(which):
    if which == "ours-shm":
        ChessModelAPI = lambda m: caz.ChessModelAPI(m, transport="shm")  # noqa: E731
    elif which == "reference":
        if not os.path.exists("/root/reference/src"):
            pytest.skip("reference checkout absent (GPU box)")
        import sys
        sys.path.insert(0, "/root/reference/src")
        from chess_zero.agent.api_chess import ChessModelAPI
    else:
        ChessModelAPI = caz.ChessModelAPI
    api = ChessModelAPI(_FakeModel())
    pipes = [api.create_pipe() for _ in range(8)]
    api.start()
    out, threads = [], []
    for k, p in enumerate(pipes):
        t = threading.Thread(target=_client, args=(p, k, out))
        t.start()
        threads.append(t)
    for t in threads:
        t.join(timeout=30)
    assert len(out) == 160 and all(all(r) for r in out)


@pytest.mark.parametrize("transport", ["pipe", "shm"])
def test_pipe_server_delivers_failures_instead_of_hanging(transport):
    class Broken:
        class model:
            @staticmethod
            def predict_on_batch(data):
                raise RuntimeError("device lost")
    api = caz.ChessModelAPI(Broken(), transport=transport)
    pipe = api.create_pipe()
    api.start()
    pipe.send(np.zeros((18, 8, 8), np.float32))
    assert pipe.poll(5.0)
    assert isinstance(pipe.recv(), RuntimeError)


def test_pointer_cache_only_for_arrays_that_cannot_move():
    """_lib.ptr remembers the data pointer of arrays that do not own their memory (pinned views), never of ones that do
    (an owning array may be resized in place), and a dead array's id cannot alias a new one."""
    import ctypes
    from chess_alpha_zero_b200 import _lib
    buf = (ctypes.c_uint8 * 64)()
    view = np.frombuffer(buf, dtype=np.uint8)
    p1, p2 = _lib.ptr(view), _lib.ptr(view)
    assert p1 is p2 and p1.value == ctypes.addressof(buf)
    own = np.zeros(16, np.float32)
    q1, q2 = _lib.ptr(own), _lib.ptr(own)
    assert q1 is not q2 and q1.value == q2.value == own.ctypes.data
    key = id(view)
    del view
    other = np.frombuffer((ctypes.c_uint8 * 64)(), dtype=np.uint8)
    assert _lib.ptr(other).value == other.ctypes.data          # whatever id it got, the weak reference check holds
    assert _lib.ptr(None) is None and key is not None


def test_tools_compile():
    """Every measurement tool at least parses (they run on the GPU box only)."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "tools", "*.py"))) + [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    assert len(files) > 15
    import ast
    for f in files:
        ast.parse(open(f).read(), filename=f)
