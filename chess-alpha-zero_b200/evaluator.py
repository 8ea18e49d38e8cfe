"""The object that stands where ``ChessModel.model`` (a Keras Model) stands in the reference.

``Evaluator.predict_on_batch(data)`` has the contract of the one Keras call on the hot path,
    policy_ary, value_ary = self.agent_model.model.predict_on_batch(data)
(/root/reference/src/chess_zero/agent/api_chess.py:66-67): float32 (B,18,8,8) in, [float32 (B,1968),
float32 (B,1)] out, rows in request order.  Everything between is the sm_100a library.
"""
import ctypes
import threading

import numpy as np

from . import _lib
from .keras_graph import parse_config

_PRECISIONS = {"bf16": _lib.PRECISION_BF16, "fp16": _lib.PRECISION_FP16, "fp32": _lib.PRECISION_FP32,
               "bf16_pure": _lib.PRECISION_BF16_PURE}


class Evaluator:
    def __init__(self, keras_config, weights, device=0, precision="bf16", max_batch=4096):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self._lib = _lib.load()
        self.keras_config = keras_config
        self.arch, self.weight_order = parse_config(keras_config)
        self.precision = precision
        self.device = int(device)
        self._lock = threading.Lock()          # one in-flight call per handle (ABI contract)
        self._handle = ctypes.c_void_p()
        tensors, keep = self._pack(weights)
        rc = self._lib.caz_create(ctypes.byref(self.arch), tensors, len(keep), self.device,
                                  _PRECISIONS[precision], int(max_batch), ctypes.byref(self._handle))
        _lib.check(None, rc)
        self.weights = weights                 # {Keras name: array}: what fit() starts from and ChessModel.save writes
        self._trainer = None
        self._loss_weights = None
        self.max_batch = int(max_batch)
        self.n_labels = int(self.arch.n_labels)
        self.input_planes = int(self.arch.input_planes)

    # -- weights ---------------------------------------------------------------------------------
    def _pack(self, weights):
        keep = []
        for name in self.weight_order:
            if name not in weights:
                raise KeyError(f"weight {name} missing")
            keep.append(np.ascontiguousarray(weights[name], dtype=np.float32))
        arr = (_lib.Tensor * len(keep))()
        for i, a in enumerate(keep):
            arr[i].data = a.ctypes.data
            arr[i].numel = a.size
        return arr, keep

    def reload(self, weights):
        """Swap the weight set in place; pipes and handles stay open (lib/model_helper.py:26-41)."""
        tensors, keep = self._pack(weights)
        with self._lock:
            _lib.check(self._handle, self._lib.caz_reload(self._handle, tensors, len(keep)))
        self.weights = weights

    # -- training: the two Keras calls OptimizeWorker makes on model.model (worker/optimize.py:88-103) ------------
    def compile(self, optimizer=None, loss=None, loss_weights=None, **_):
        """model.compile(optimizer=Adam(), loss=['categorical_crossentropy', 'mean_squared_error'], loss_weights=[1.25, 1.0])
        (optimize.py:97-103).  The optimizer is Adam with the Keras 2.0.8 defaults whatever is passed (the reference passes
        Adam()); the two losses are the ones the reference names; only loss_weights is a free parameter."""
        if loss is not None and list(loss) != ["categorical_crossentropy", "mean_squared_error"]:
            raise ValueError("only loss=['categorical_crossentropy', 'mean_squared_error'] is implemented")
        self._loss_weights = [float(v) for v in loss_weights] if loss_weights is not None else [1.0, 1.0]
        self._trainer = None

    def fit(self, x, y, batch_size=32, epochs=1, shuffle=True, validation_split=0.0, callbacks=None, precision="bf16", seed=None, **_):
        """model.fit(state_ary, [policy_ary, value_ary], batch_size=..., epochs=..., shuffle=True, validation_split=0.02,
        callbacks=[...]) (optimize.py:88-93) on the GPU (trainer.py); afterwards this evaluator predicts with the trained weights
        (inference-mode BatchNormalization with the updated moving statistics).  Returns an object with a Keras-like
        ``.history`` dict.  Callbacks are ignored."""
        from .trainer import TrainConfig, Trainer
        if self._loss_weights is None:
            raise RuntimeError("compile() first, as Keras demands")
        if self._trainer is None or self._trainer.max_batch < batch_size or self._trainer.precision != precision:
            if self._trainer is not None:
                self._trainer.close()
            cfg = TrainConfig.normal(policy_loss_weight=self._loss_weights[0], value_loss_weight=self._loss_weights[1])
            self._trainer = Trainer(self.keras_config, self.weights, device=self.device, precision=precision, max_batch=batch_size, config=cfg)
        hist = self._trainer.fit(x, y, batch_size=batch_size, epochs=epochs, shuffle=shuffle, validation_split=validation_split, seed=seed)
        self.reload(self._trainer.weights())

        class History:
            history = hist
        return History()

    # -- the operator ----------------------------------------------------------------------------
    def predict_on_batch(self, data, legal_mask=None):
        x = np.ascontiguousarray(data, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (self.input_planes, 8, 8):
            raise ValueError(f"expected (B,{self.input_planes},8,8) float32 planes, got {x.shape}")
        b = x.shape[0]
        policy = np.empty((b, self.n_labels), np.float32)
        value = np.empty((b, 1), np.float32)
        flags, mask = 0, None
        if legal_mask is not None:
            mask = _lib.as_array(legal_mask, np.uint8, (b, self.n_labels))
            flags = _lib.EVAL_MASKED_SOFTMAX
        with self._lock:
            rc = self._lib.caz_eval(self._handle, _lib.ptr(x), b, _lib.ptr(policy), _lib.ptr(value),
                                    _lib.ptr(mask), flags)
            _lib.check(self._handle, rc)
        return [policy, value]

    def predict_into(self, planes, policy, value, legal_mask=None):
        """Same call on caller-owned (ideally pinned) buffers: no allocation, no copies on the host."""
        b = planes.shape[0]
        flags = _lib.EVAL_MASKED_SOFTMAX if legal_mask is not None else 0
        with self._lock:
            rc = self._lib.caz_eval(self._handle, _lib.ptr(planes), b, _lib.ptr(policy), _lib.ptr(value),
                                    _lib.ptr(legal_mask), flags)
            _lib.check(self._handle, rc)

    def submit(self, planes, policy, value, legal_mask=None):
        """Non-blocking predict_into: returns a ticket for collect() (tickets count up; up to 16 may be outstanding); the host
        buffers (ideally pinned) must stay untouched until collected.  Copies of one batch then overlap the forward
        pass of the other."""
        ticket = ctypes.c_int32(-1)
        flags = _lib.EVAL_MASKED_SOFTMAX if legal_mask is not None else 0
        with self._lock:
            rc = self._lib.caz_submit(self._handle, _lib.ptr(planes), planes.shape[0], _lib.ptr(policy),
                                      _lib.ptr(value), _lib.ptr(legal_mask), flags, ctypes.byref(ticket))
            _lib.check(self._handle, rc)
        return ticket.value

    def submit_records(self, records, policy, value):
        """submit() from board records (uint8 (B, 68)); policy (>= B, n_labels) and value (>= B,) are caller-owned,
        ideally pinned.  A search driver with several groups of games submits under its own lock and collects outside
        it: one group's read-back then overlaps another group's forward pass."""
        r = _lib.as_array(records, np.uint8)
        b = r.shape[0]
        if r.ndim != 2 or r.shape[1] != _lib.RECORD_BYTES or policy.shape[0] < b or value.shape[0] < b:
            raise ValueError("records must be uint8 (B, 68) and the outputs must hold B rows")
        ticket = ctypes.c_int32(-1)
        with self._lock:
            rc = self._lib.caz_submit_records(self._handle, _lib.ptr(r), b, _lib.ptr(policy), _lib.ptr(value), None, 0,
                                              ctypes.byref(ticket))
            _lib.check(self._handle, rc)
        return ticket.value

    def submit_records_priors(self, records, offsets, labels, priors, value):
        """submit_records() that returns one float64 prior per legal move (CSR `offsets` int32 (B + 1,), `labels` int32
        (>= offsets[B],) as SelfPlaySearch.collect_csr hands them out) instead of the policy rows: the prior push of
        select_action_q_and_u runs on the device, 8 bytes per legal move come back instead of 7 872 per position."""
        r = _lib.as_array(records, np.uint8)
        b = r.shape[0]
        if r.ndim != 2 or r.shape[1] != _lib.RECORD_BYTES or offsets.dtype != np.int32 or labels.dtype != np.int32 \
                or priors.dtype != np.float64 or offsets.shape[0] < b + 1 or value.shape[0] < b:
            raise ValueError("records uint8 (B, 68); offsets / labels int32; priors float64; value float32 (>= B,)")
        if priors.shape[0] < int(offsets[b]) or labels.shape[0] < int(offsets[b]):
            raise ValueError("labels / priors are shorter than offsets[B]")
        ticket = ctypes.c_int32(-1)
        with self._lock:
            rc = self._lib.caz_submit_records_priors(self._handle, _lib.ptr(r), b, _lib.ptr(offsets), _lib.ptr(labels),
                                                     _lib.ptr(priors), _lib.ptr(value), ctypes.byref(ticket))
            _lib.check(self._handle, rc)
        return ticket.value

    def collect(self, ticket):
        _lib.check(self._handle, self._lib.caz_collect(self._handle, int(ticket)))

    def predict_records(self, records, legal_mask=None):
        """Board records (68 B each, see encoder.fen_to_record) in; the encoder runs on the device."""
        r = _lib.as_array(records, np.uint8)
        if r.ndim != 2 or r.shape[1] != _lib.RECORD_BYTES:
            raise ValueError("records must be uint8 (B, 68)")
        b = r.shape[0]
        policy = np.empty((b, self.n_labels), np.float32)
        value = np.empty((b, 1), np.float32)
        flags, mask = 0, None
        if legal_mask is not None:
            mask = _lib.as_array(legal_mask, np.uint8, (b, self.n_labels))
            flags = _lib.EVAL_MASKED_SOFTMAX
        with self._lock:
            rc = self._lib.caz_eval_records(self._handle, _lib.ptr(r), b, _lib.ptr(policy), _lib.ptr(value),
                                            _lib.ptr(mask), flags)
            _lib.check(self._handle, rc)
        return [policy, value]

    def predict_records_into(self, records, policy, value):
        """predict_records with caller-owned (ideally pinned, _lib.PinnedArray) outputs: policy (B, n_labels), value (B,)."""
        r = _lib.as_array(records, np.uint8)
        b = r.shape[0]
        if r.ndim != 2 or r.shape[1] != _lib.RECORD_BYTES or policy.shape[0] < b or value.shape[0] < b:
            raise ValueError("records must be uint8 (B, 68) and the outputs must hold B rows")
        with self._lock:
            rc = self._lib.caz_eval_records(self._handle, _lib.ptr(r), b, _lib.ptr(policy), _lib.ptr(value), None, 0)
            _lib.check(self._handle, rc)

    def eval_device(self, d_planes_ptr, batch, d_policy_ptr, d_value_ptr, d_mask_ptr=None):
        """Raw device pointers (ints); enqueues on the handle's stream, no synchronisation."""
        flags = _lib.EVAL_MASKED_SOFTMAX if d_mask_ptr else 0
        with self._lock:
            rc = self._lib.caz_eval_device(self._handle, ctypes.c_void_p(d_planes_ptr), int(batch),
                                           ctypes.c_void_p(d_policy_ptr), ctypes.c_void_p(d_value_ptr),
                                           ctypes.c_void_p(d_mask_ptr) if d_mask_ptr else None, flags)
            _lib.check(self._handle, rc)

    def set_small_batch_limit(self, limit):
        """Batches <= limit use the persistent small-batch tower kernel; -1 = automatic, 0 = never."""
        _lib.check(self._handle, self._lib.caz_set_small_batch_limit(self._handle, int(limit)))

    def set_fused_tower(self, mode):
        """-1: one persistent launch for stem + tower, form by batch size (default); 0: one launch per layer; 1 / 2: forced
        forms of the one-launch tower (include/caz_b200.h)."""
        _lib.check(self._handle, self._lib.caz_set_fused_tower(self._handle, int(mode)))

    def sync(self):
        _lib.check(self._handle, self._lib.caz_sync(self._handle))

    # -- encoder / PUCT (same device, same stream) -----------------------------------------------
    def encode(self, records):
        r = _lib.as_array(records, np.uint8)
        if r.ndim != 2 or r.shape[1] != _lib.RECORD_BYTES:
            raise ValueError("records must be uint8 (B, 68)")
        planes = np.empty((r.shape[0], 18, 8, 8), np.float32)
        with self._lock:
            _lib.check(self._handle, self._lib.caz_encode(self._handle, _lib.ptr(r), r.shape[0], _lib.ptr(planes)))
        return planes

    def push_priors(self, policy, offsets, label_index):
        pol = _lib.as_array(policy, np.float32)
        off = _lib.as_array(offsets, np.int32)
        idx = _lib.as_array(label_index, np.int32)
        n_nodes = off.size - 1
        if pol.shape != (n_nodes, pol.shape[-1]):
            raise ValueError("policy must be (n_nodes, n_labels)")
        out = np.empty(idx.size, np.float64)
        with self._lock:
            rc = self._lib.caz_push_priors(self._handle, _lib.ptr(pol), pol.shape[1], _lib.ptr(off), n_nodes,
                                           _lib.ptr(idx), _lib.ptr(out))
            _lib.check(self._handle, rc)
        return out

    def puct_select(self, offsets, n, q, p, sum_n, c_puct, noise=None, is_root=None, noise_eps=0.0):
        off = _lib.as_array(offsets, np.int32)
        n_nodes = off.size - 1
        n_ = _lib.as_array(n, np.int32)
        q_ = _lib.as_array(q, np.float64)
        p_ = _lib.as_array(p, np.float64)
        sn = _lib.as_array(sum_n, np.int32, (n_nodes,))
        nz = None if noise is None else _lib.as_array(noise, np.float64)
        rt = None if is_root is None else _lib.as_array(is_root, np.uint8, (n_nodes,))
        best = np.empty(n_nodes, np.int32)
        with self._lock:
            rc = self._lib.caz_puct_select(self._handle, _lib.ptr(off), n_nodes, _lib.ptr(n_), _lib.ptr(q_),
                                           _lib.ptr(p_), _lib.ptr(sn), _lib.ptr(nz), _lib.ptr(rt), float(c_puct),
                                           float(noise_eps), _lib.ptr(best))
            _lib.check(self._handle, rc)
        return best

    def debug_conv_layer(self, layer, x_nhwc, residual=None):
        """Test hook: one conv layer (0 = stem, then conv1/conv2 per block) through the build's real kernel."""
        x = _lib.as_array(x_nhwc, np.float32)
        b = x.shape[0]
        res = None if residual is None else _lib.as_array(residual, np.float32, (b, 64, self.arch.filters))
        out = np.empty((b, 64, self.arch.filters), np.float32)
        with self._lock:
            rc = self._lib.caz_debug_conv_layer(self._handle, int(layer), _lib.ptr(x), _lib.ptr(res), b, _lib.ptr(out))
            _lib.check(self._handle, rc)
        return out

    # -- measurement -----------------------------------------------------------------------------
    @property
    def stream(self):
        return self._lib.caz_stream(self._handle)

    @property
    def kernel_launches(self):
        return int(self._lib.caz_kernel_launches(self._handle))

    @property
    def weight_bytes(self):
        return int(self._lib.caz_weight_bytes(self._handle))

    @property
    def inexact_weights(self):
        """bf16 build: folded bf16 weights that the fp16 encoding read by the tensor core does not hold exactly."""
        return int(self._lib.caz_inexact_weights(self._handle))

    def profile(self, enable=True):
        _lib.check(self._handle, self._lib.caz_profile(self._handle, 1 if enable else 0))

    def profile_read(self, reset=True):
        ms = (ctypes.c_float * _lib.STAGES)()
        n = (ctypes.c_int64 * _lib.STAGES)()
        _lib.check(self._handle, self._lib.caz_profile_read(self._handle, ms, n, 1 if reset else 0))
        return {name: {"ms": float(ms[i]), "launches": int(n[i])} for i, name in enumerate(_lib.STAGE_NAMES)}

    def _close_trainer(self):
        if getattr(self, "_trainer", None) is not None:
            self._trainer.close()
            self._trainer = None

    def close(self):
        self._close_trainer()
        if getattr(self, "_handle", None) and self._handle.value:
            self._lib.caz_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
