"""Builds the native library in-tree: csrc/caz_abi.cu -> _native/libcaz_b200.so (sm_100a only)."""
import contextlib
import fcntl
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_native")
LIB_PATH = os.path.join(OUT_DIR, "libcaz_b200.so")
MCTS_LIB_PATH = os.path.join(OUT_DIR, "libcaz_mcts.so")   # host-only search driver (include/caz_mcts.h), g++
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-I", INCLUDE]


@contextlib.contextmanager
def _build_lock():
    """One builder at a time (torchrun starts every rank at once); the output appears atomically (os.replace), so a
    rank that dlopens while another one builds sees either the old library or the new one, never a partial file."""
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def _compile(cmd, out_path, what):
    tmp = f"{out_path}.tmp.{os.getpid()}"
    res = subprocess.run(cmd + ["-o", tmp], capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(f"{what} failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, out_path)
    return res


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(INCLUDE, "caz_b200.h")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > built for s in _sources())


def build_native(force=False, verbose=False):
    """Compile with nvcc (cross-compiles without a GPU). Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcaz_b200.so (there is no CPU fallback)")
    with _build_lock():
        if not force and not is_stale():     # another rank built it while we waited for the lock
            return LIB_PATH
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "caz_abi.cu")]
        res = _compile(cmd, LIB_PATH, "nvcc")
    if verbose:
        print(res.stderr)
    return LIB_PATH


def _mcts_sources():
    chess = os.path.join(CSRC, "chess")
    return [os.path.join(CSRC, "caz_mcts.cpp"), os.path.join(INCLUDE, "caz_mcts.h")] + \
        sorted(os.path.join(chess, f) for f in os.listdir(chess))


def mcts_is_stale():
    if not os.path.exists(MCTS_LIB_PATH):
        return True
    built = os.path.getmtime(MCTS_LIB_PATH)
    return any(os.path.getmtime(s) > built for s in _mcts_sources())


def build_mcts(force=False):
    """Compile the host-side batched search (no CUDA in it). Returns the library path."""
    if not force and not mcts_is_stale():
        return MCTS_LIB_PATH
    gxx = shutil.which("g++")
    if not gxx:
        raise RuntimeError("g++ not found: cannot build libcaz_mcts.so")
    with _build_lock():
        if not force and not mcts_is_stale():
            return MCTS_LIB_PATH
        cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", INCLUDE, "-I", CSRC,
               os.path.join(CSRC, "caz_mcts.cpp")]
        _compile(cmd, MCTS_LIB_PATH, "g++")
    return MCTS_LIB_PATH


if __name__ == "__main__":
    print(build_mcts(force=True))
    print(build_native(force=True, verbose="-v" in __import__("sys").argv))
