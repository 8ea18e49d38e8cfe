"""The training step of the reference's `opt` worker on the B200.

``Trainer`` stands where ``OptimizeWorker`` uses the Keras model for training
(/root/reference/src/chess_zero/worker/optimize.py:79-103):

    opt = Adam(); model.compile(optimizer=opt, loss=['categorical_crossentropy', 'mean_squared_error'],
                                loss_weights=[1.25, 1.0])                                  # compile_model
    model.fit(state_ary, [policy_ary, value_ary], batch_size=384, epochs=1, shuffle=True,
              validation_split=0.02)                                                       # train_epoch

``Trainer.fit`` has that signature and meaning; ``train_on_batch`` is one step.  Under torch.distributed (one process per
GPU) every rank computes the gradients of its shard of the batch, the flat gradient vector is summed with an NCCL
all-reduce, and every rank applies the same Adam update: the only collective anywhere in this repository.
"""
import ctypes

import numpy as np

from . import _lib
from .keras_graph import parse_config


class TrainConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("learning_rate", "beta_1", "beta_2", "epsilon", "l2", "policy_loss_weight",
                                              "value_loss_weight", "bn_momentum")]

    @classmethod
    def normal(cls, **over):
        """keras.optimizers.Adam() defaults (keras 2.0.8), ModelConfig.l2_reg and TrainerConfig.loss_weights of
        configs/normal.py:46-68, BatchNormalization's default momentum."""
        c = cls(1e-3, 0.9, 0.999, 1e-8, 1e-4, 1.25, 1.0, 0.99)
        for k, v in over.items():
            setattr(c, k, v)
        return c


class Trainer:
    def __init__(self, keras_config, weights, device=0, precision="fp32", max_batch=384, config=None, distributed=None):
        lib = self._lib = _lib.load()
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        lib.caz_train_create.argtypes = [ctypes.POINTER(_lib.Arch), ctypes.POINTER(_lib.Tensor), i32, i32, i32, i32,
                                         ctypes.POINTER(TrainConfig), vp, ctypes.POINTER(vp)]
        lib.caz_train_destroy.argtypes = [vp]
        lib.caz_train_destroy.restype = None
        lib.caz_train_last_error.argtypes = [vp]
        lib.caz_train_last_error.restype = ctypes.c_char_p
        lib.caz_train_param_count.argtypes = [ctypes.POINTER(_lib.Arch)]
        lib.caz_train_param_count.restype = i64
        lib.caz_train_num_params.argtypes = [vp]
        lib.caz_train_num_params.restype = i64
        lib.caz_train_steps.argtypes = [vp]
        lib.caz_train_steps.restype = i64
        lib.caz_train_forward_backward.argtypes = [vp, vp, vp, vp, i32, vp]
        lib.caz_train_forward_backward_begin.argtypes = [vp, vp, vp, vp, i32]
        lib.caz_train_forward_backward_end.argtypes = [vp, vp]
        lib.caz_train_wait_gradients.argtypes = [vp, i32, vp]
        lib.caz_train_gradient_split.argtypes = [vp]
        lib.caz_train_gradient_split.restype = i64
        lib.caz_train_apply.argtypes = [vp, ctypes.c_float]
        lib.caz_train_read.argtypes = [vp, i32, vp, i64]
        lib.caz_train_read_outputs.argtypes = [vp, i32, vp, vp]
        lib.caz_train_grad_device_ptr.argtypes = [vp]
        lib.caz_train_grad_device_ptr.restype = vp
        self.keras_config = keras_config
        self.arch, self.weight_order = parse_config(keras_config)
        self.config = config or TrainConfig.normal()
        self.device, self.max_batch = int(device), int(max_batch)
        keep = [np.ascontiguousarray(weights[n], dtype=np.float32) for n in self.weight_order]
        self._shapes = [tuple(weights[n].shape) for n in self.weight_order]
        arr = (_lib.Tensor * len(keep))()
        for i, a in enumerate(keep):
            arr[i].data, arr[i].numel = a.ctypes.data, a.size
        self.n_params = int(lib.caz_train_param_count(ctypes.byref(self.arch)))
        # the gradient vector lives in a torch tensor when an all-reduce has to run on it
        self._dist = None
        self._grad_tensor = None
        grad_ptr = None
        if distributed is None:
            try:
                import torch.distributed as dist
                distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            except ImportError:
                distributed = False
        if distributed:
            import torch
            import torch.distributed as dist
            self._dist = dist
            self._grad_tensor = torch.zeros(self.n_params, dtype=torch.float32, device=torch.device("cuda", self.device))
            grad_ptr = ctypes.c_void_p(self._grad_tensor.data_ptr())
        self._h = ctypes.c_void_p()
        codes = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}
        if precision not in codes:
            raise ValueError(f"precision must be one of {sorted(codes)}")
        rc = lib.caz_train_create(ctypes.byref(self.arch), arr, len(keep), self.device, codes[precision], self.max_batch,
                                  ctypes.byref(self.config), grad_ptr, ctypes.byref(self._h))
        if rc:
            raise _lib.CazError(rc, (lib.caz_last_error(None) or b"unknown").decode())
        assert int(lib.caz_train_num_params(self._h)) == self.n_params
        self.precision = precision

    # -- plumbing ----------------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise _lib.CazError(rc, (self._lib.caz_train_last_error(self._h) or b"unknown").decode())

    def close(self):
        if self._h:
            self._lib.caz_train_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _split(self, flat):
        out, off = {}, 0
        for name, shape in zip(self.weight_order, self._shapes):
            n = int(np.prod(shape))
            out[name] = flat[off:off + n].reshape(shape).copy()
            off += n
        return out

    def _read(self, what):
        flat = np.empty(self.n_params, np.float32)
        self._check(self._lib.caz_train_read(self._h, what, _lib.ptr(flat), self.n_params))
        return self._split(flat)

    def weights(self):
        """{Keras weight name: array} of the current parameters (what ChessModel.save / Evaluator.reload take)."""
        return self._read(0)

    def gradients(self):
        """Gradients of the data loss left by the last forward_backward (the l2 term is added inside apply); the entries of
        the moving statistics hold the batch mean / variance of that pass."""
        return self._read(1)

    @property
    def steps(self):
        return int(self._lib.caz_train_steps(self._h))

    # -- the step ----------------------------------------------------------------------------------
    def forward_backward(self, planes, policy_target, value_target):
        x = _lib.as_array(planes, np.float32)
        pt = _lib.as_array(policy_target, np.float32, (x.shape[0], self.arch.n_labels))
        vt = _lib.as_array(value_target, np.float32).reshape(-1)
        if vt.shape[0] != x.shape[0]:
            raise ValueError("value_target must have one entry per position")
        losses = np.zeros(4, np.float32)
        self._check(self._lib.caz_train_forward_backward(self._h, _lib.ptr(x), _lib.ptr(pt), _lib.ptr(vt), x.shape[0], _lib.ptr(losses)))
        return {"loss": float(losses[0]), "policy_out_loss": float(losses[1]), "value_out_loss": float(losses[2]), "l2": float(losses[3])}

    def outputs(self, batch):
        p = np.empty((batch, self.arch.n_labels), np.float32)
        v = np.empty((batch, 1), np.float32)
        self._check(self._lib.caz_train_read_outputs(self._h, batch, _lib.ptr(p), _lib.ptr(v)))
        return p, v

    def apply(self):
        scale = 1.0
        if self._dist is not None:
            import torch
            self._dist.all_reduce(self._grad_tensor, op=self._dist.ReduceOp.SUM)      # NCCL over NVLink / NVSwitch
            torch.cuda.synchronize(self.device)
            scale = 1.0 / self._dist.get_world_size()
        self._check(self._lib.caz_train_apply(self._h, scale))

    def train_on_batch(self, planes, policy_target, value_target):
        """One optimizer step on this rank's shard of the batch (model.train_on_batch semantics).  Data parallel: the
        all-reduce of the upper half of the gradient vector (the layers the backward pass finishes first) runs under the
        backward pass of the lower half."""
        if self._dist is None:
            out = self.forward_backward(planes, policy_target, value_target)
            self.apply()
            return out
        import torch
        x = _lib.as_array(planes, np.float32)
        pt = _lib.as_array(policy_target, np.float32, (x.shape[0], self.arch.n_labels))
        vt = _lib.as_array(value_target, np.float32).reshape(-1)
        self._check(self._lib.caz_train_forward_backward_begin(self._h, _lib.ptr(x), _lib.ptr(pt), _lib.ptr(vt), x.shape[0]))
        split = int(self._lib.caz_train_gradient_split(self._h))
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._check(self._lib.caz_train_wait_gradients(self._h, 0, stream))
        works = [self._dist.all_reduce(self._grad_tensor[split:], op=self._dist.ReduceOp.SUM, async_op=True)]
        self._check(self._lib.caz_train_wait_gradients(self._h, 1, stream))
        if split > 0:
            works.append(self._dist.all_reduce(self._grad_tensor[:split], op=self._dist.ReduceOp.SUM, async_op=True))
        losses = np.zeros(4, np.float32)
        self._check(self._lib.caz_train_forward_backward_end(self._h, _lib.ptr(losses)))
        for wk in works:
            wk.wait()
        torch.cuda.synchronize(self.device)
        self._check(self._lib.caz_train_apply(self._h, 1.0 / self._dist.get_world_size()))
        return {"loss": float(losses[0]), "policy_out_loss": float(losses[1]), "value_out_loss": float(losses[2]), "l2": float(losses[3])}

    def fit(self, x, y, batch_size=384, epochs=1, shuffle=True, validation_split=0.0, seed=None):
        """model.fit(state_ary, [policy_ary, value_ary], batch_size, epochs, shuffle, validation_split) of
        optimize.py:88-93: the LAST validation_split of the samples is held out (before shuffling), every epoch walks
        the shuffled rest in batches (the last one may be short).  Returns {"loss": [...], "policy_out_loss": [...], ...}."""
        state, (policy, value) = np.asarray(x, np.float32), y
        policy, value = np.asarray(policy, np.float32), np.asarray(value, np.float32).reshape(-1)
        n_val = int(len(state) * validation_split)
        n = len(state) - n_val
        rng = np.random.default_rng(seed)
        hist = {"loss": [], "policy_out_loss": [], "value_out_loss": []}
        for _ in range(epochs):
            order = rng.permutation(n) if shuffle else np.arange(n)
            acc, seen = np.zeros(3), 0
            for b0 in range(0, n, batch_size):
                idx = order[b0:b0 + batch_size]
                out = self.train_on_batch(state[idx], policy[idx], value[idx])
                acc += len(idx) * np.array([out["loss"], out["policy_out_loss"], out["value_out_loss"]])
                seen += len(idx)
            for k, v in zip(hist, acc / max(seen, 1)):
                hist[k].append(float(v))
        return hist
