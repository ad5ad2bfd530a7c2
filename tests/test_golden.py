"""Committed fixtures (tests/golden/hnsw_golden_v1.npz, made by tests/golden/make_golden.py).

They were produced by this repo's CPU oracle -- the reference mount has no golden vectors
(/root/reference/README.md:1) -- so they pin the oracle and the CUDA path against drift, not against
pgvector.  CPU part: the oracle still builds the same graphs and returns the same ids/distances.
GPU part: the CUDA scan on the fixture graph, and the CUDA sequential build, reproduce them bit for bit."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hnsw_golden_v1.npz")
CASES = ["l2_f32_sift", "cos_f32", "ip_f16"]
OPC = {(0, 0): "vector_l2_ops", (1, 0): "vector_ip_ops", (2, 0): "vector_cosine_ops",
       (0, 1): "halfvec_l2_ops", (1, 1): "halfvec_ip_ops", (2, 1): "halfvec_cosine_ops"}


def load(name):
    z = np.load(GOLD)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def graph_matches(g, f):
    n, entry, ur = (int(v) for v in f["n"])
    assert (g.n, g.entry, g.upper_rows) == (n, entry, ur)
    assert (g.level[:n] == f["level"]).all()
    assert (g.nbr0[:n] == f["nbr0"]).all()
    assert (g.uoff[:n] == f["uoff"]).all()
    assert (g.nbru[:ur] == f["nbru"]).all()
    assert (g.ntids[:n] == f["ntids"]).all()
    for e in range(n):
        assert (g.tids[e, :g.ntids[e]] == f["tids"][e, :f["ntids"][e]]).all()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle, name):
    f = load(name)
    metric, dtype, dim, m, efc, seed, ef = (int(v) for v in f["params"])
    ix = oracle.Index(dim, m, efc, metric, dtype, oracle.CANON, seed=seed)
    ix.build(f["x"])
    graph_matches(ix.export(), f)
    e, d, c, _ = ix.search_batch(f["q"], ef, threads=1)
    assert (c == f["res_cnt"]).all() and (e == f["res_elem"]).all()
    assert (d.view(np.uint32) == f["res_dist"].view(np.uint32)).all()


def test_golden_has_the_edge_cases():
    f = load("cos_f32")
    assert f["ntids"].max() >= 3          # folded duplicates
    assert int(f["n"][2]) > 0             # upper layers exist
    f = load("l2_f32_sift")
    d = f["res_dist"]
    assert (d[:, 1:] == d[:, :-1]).any()  # exact distance ties in the result lists


def test_oracle_reproduces_golden_resumable_scans_and_l1(oracle):
    """hnsw_golden_iter_l1_v1.npz: the first batches of resumable scans (with and without max_scan_tuples) and
    an L1 index on the cos_f32 rows"""
    f = load("cos_f32")
    z = np.load(os.path.join(os.path.dirname(GOLD), "hnsw_golden_iter_l1_v1.npz"))
    metric, dtype, dim, m, efc, seed, ef = (int(v) for v in f["params"])
    ix = oracle.Index(dim, m, efc, metric, dtype, oracle.CANON, seed=seed)
    ix.build(f["x"])
    for qi in range(4):
        for tag, mt in (("all", 10 ** 9), ("cap300", 300)):
            batches, _, _ = ix.iterate(f["q"][qi], 10, max_scan_tuples=mt, max_batches=12)
            assert [len(b[0]) for b in batches] == z["iter/%d/%s/sizes" % (qi, tag)].tolist()
            assert (np.concatenate([b[0] for b in batches]) == z["iter/%d/%s/elem" % (qi, tag)]).all()
            assert (np.concatenate([b[1] for b in batches]).view(np.uint32) == z["iter/%d/%s/dist" % (qi, tag)].view(np.uint32)).all()
    l1 = oracle.Index(dim, m, efc, oracle.L1, dtype, oracle.CANON, seed=seed)
    l1.build(f["x"])
    assert (l1.export().nbr0[:len(z["l1/nbr0"])] == z["l1/nbr0"]).all()
    e, d, _, _ = l1.search_batch(f["q"], ef, threads=1)
    assert (e == z["l1/res_elem"]).all() and (d.view(np.uint32) == z["l1/res_dist"].view(np.uint32)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_scan_reproduces_golden(oracle, pkg, name):
    f = load(name)
    metric, dtype, dim, m, efc, seed, ef = (int(v) for v in f["params"])
    n, entry, ur = (int(v) for v in f["n"])
    orc = oracle.Index(dim, m, efc, metric, dtype, oracle.CANON, seed=seed)
    orc.build(f["x"])
    ix = pkg.HnswIndex(dim, OPC[(metric, dtype)], m, efc, capacity=n + 8, seed=seed)
    ix.load_graph(orc.export())
    e, d, c = ix.search_elements(f["q"], ef)
    assert (c == f["res_cnt"]).all() and (e == f["res_elem"]).all()
    assert (d.view(np.uint32) == f["res_dist"].view(np.uint32)).all()
    ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("link_kernel", [0, 1, 2, 3])
def test_cuda_sequential_build_reproduces_golden(pkg, name, link_kernel):
    """every HnswUpdateConnection kernel (memoising, warp, pipelined) builds the fixture graph"""
    f = load(name)
    metric, dtype, dim, m, efc, seed, ef = (int(v) for v in f["params"])
    ix = pkg.HnswIndex(dim, OPC[(metric, dtype)], m, efc, capacity=len(f["x"]), seed=seed)
    ix.set_option("build_batch", 1)
    ix.set_option("link_kernel", link_kernel)
    ix.build(f["x"])
    graph_matches(ix.export_graph(), f)
    e, d, c = ix.search_elements(f["q"], ef)
    assert (e == f["res_elem"]).all() and (d.view(np.uint32) == f["res_dist"].view(np.uint32)).all()
    ix.close()
