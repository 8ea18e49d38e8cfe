import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import _lib, encoder
cfg, w = bench.network(7)
ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=64)
pin_r = _lib.PinnedArray((16, 68), np.uint8); pin_p = _lib.PinnedArray((16, 1968), np.float32); pin_v = _lib.PinnedArray((16, 1), np.float32)
pin_r.array[:] = encoder.fens_to_records(encoder.synthetic_fens(16, seed=3))
for _ in range(200): ev.predict_records_into(pin_r.array, pin_p.array, pin_v.array)
ts = []
for _ in range(2000):
    t0 = time.perf_counter(); ev.predict_records_into(pin_r.array, pin_p.array, pin_v.array); ts.append(time.perf_counter() - t0)
print("records host call p50 us:", np.median(ts) * 1e6)
# raw ctypes call with prebuilt pointers
lib = _lib.load()
import ctypes
h = ev._handle
pr, pp, pv = ctypes.c_void_p(pin_r.array.ctypes.data), ctypes.c_void_p(pin_p.array.ctypes.data), ctypes.c_void_p(pin_v.array.ctypes.data)
ts = []
for _ in range(2000):
    t0 = time.perf_counter(); lib.caz_eval_records(h, pr, 16, pp, pv, None, 0); ts.append(time.perf_counter() - t0)
print("raw ctypes call p50 us:", np.median(ts) * 1e6)
ev.close()
