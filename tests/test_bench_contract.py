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


def test_cpu_arm_child_process(tmp_path):
    """oracle/cpu_arm.py (the process whose time bench.py reports as the CPU arm): loads a flat graph image,
    times bounded steps, returns the ids both summation orders give; it imports neither torch nor the product."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    from conftest import clustered
    src = open(os.path.join(ROOT, "oracle", "cpu_arm.py")).read()
    assert "import torch" not in src and "pgvector_hnsw_partitioning_b200" not in src and "libhnsw_b200" not in src
    x = clustered(3000, 32, 16, seed=4)
    q = clustered(200, 32, 16, seed=5)
    orc = O.Index(32, 16, 64, O.COSINE, O.F32, O.CANON, seed=3)
    orc.build(x)
    g = orc.export()
    d = str(tmp_path)
    for k in ("vecs", "level", "nbr0", "uoff", "nbru", "ntids", "tids"):
        np.save(os.path.join(d, k + ".npy"), getattr(g, k))
    json.dump({"dim": g.dim, "m": g.m, "efc": g.efc, "metric": g.metric, "dtype": g.dtype, "n": g.n, "upper_rows": g.upper_rows,
               "entry": g.entry}, open(os.path.join(d, "meta.json"), "w"))
    np.save(os.path.join(d, "queries.npy"), q)
    np.save(os.path.join(d, "rows.npy"), x)
    job = {"graph_dir": d, "queries": os.path.join(d, "queries.npy"), "ef": 40, "steps": 3, "warmup": 1, "budget_s": 1.0, "threads": 2,
           "parity": 100, "parity_out": os.path.join(d, "parity.npz"),
           "build": {"rows": os.path.join(d, "rows.npy"), "metric": 2, "dtype": 0, "n1": 500, "parts": 2, "n_part": 300}}
    res = bench.run_cpu_child(job, d)
    assert res["kind"] == "port" and res["queries_per_s"] > 0 and res["threads"] == 2 and 64 <= res["per_step"] <= 200
    assert res["build"]["single_thread"]["vectors_per_s"] > 0 and res["build"]["concurrent"]["partitions"] == 2
    par = dict(np.load(os.path.join(d, "parity.npz")))
    oe, od, oc, octr = orc.search_batch(q[:100], 40, threads=2)
    assert (par["canon_ids"] == oe).all() and (par["canon_dist"].view(np.uint32) == od.view(np.uint32)).all()
    assert res["canon_counters"]["n_dist"] == octr["n_dist"]
    # the parity record: identical ids -> everything identical; a perturbed id that is not a tie -> unexplained
    rec = bench.parity_record(par, oe.copy(), od.copy())
    assert rec["canonical_order"]["ids_identical"] == 100 and rec["canonical_order"]["distances_bit_identical"] == 100
    nat = rec["natural_order_top10"]
    assert nat["ids_identical"] + nat["within_1e-5_ties"] == 100 and nat["unexplained"] == 0
    bad = oe.copy()
    bad[0, 0] = oe[0, 20]
    bd = od.copy()
    bd[0, 0] = od[0, 20]
    rec2 = bench.parity_record(par, bad, bd)
    assert rec2["canonical_order"]["ids_identical"] == 99 and rec2["natural_order_top10"]["unexplained"] == 1
