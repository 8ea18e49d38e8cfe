"""N>1 host logic on CPU: world_size 2, gloo, launched exactly the way the driver launches bench.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port)] + script_args
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def test_sharding_and_timing_reduction_two_ranks():
    res = _torchrun([os.path.join("tests", "_dist_worker.py")], 29631)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "DIST_OK" in res.stdout


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    res = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], 29632)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "positions/s" and d["value"] > 0
