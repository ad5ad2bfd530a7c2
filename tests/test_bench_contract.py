"""bench.py's contract where it can be checked without a GPU: the reference arm answers with one JSON line and
exit code 0, the product arm refuses to run (no CPU fallback), the helpers behave."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def has_cuda():
    import torch
    return torch.cuda.is_available()


def test_reference_arm_without_gpu_prints_one_json_line():
    if has_cuda():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d


def test_product_arm_without_gpu_fails_loudly():
    if has_cuda():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CPU fallback" in out.stderr


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_recall_helper():
    sys.path.insert(0, ROOT)
    import bench
    ids = np.array([[1, 2, 3, 4], [5, 6, 7, 8]])
    gt = np.array([[1, 2, 9, 10], [5, 6, 7, 8]])
    assert bench.recall_at(ids, gt) == 0.75
