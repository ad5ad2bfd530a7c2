"""Exact per-partition scan (config 5): bf16 tcgen05 GEMM candidate generation + fp32 re-rank,
against numpy (raw scores) and the oracle's double-precision brute force (final top-k)."""
import numpy as np
import pytest

from conftest import clustered, sift_like

pytestmark = pytest.mark.gpu


def bf16_round(a):
    """round-to-nearest-even fp32 -> bf16 -> fp32 in numpy"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("n,dim,nq", [(1000, 64, 37), (3000, 128, 300), (2500, 768, 130), (700, 100, 129)])
def test_gemm_scores_match_bf16_reference(pkg, n, dim, nq):
    """the tensor-core stage on its own: S = bf16(Q) . bf16(X)^T accumulated in fp32"""
    x = clustered(n, dim, 16, seed=n)
    q = clustered(nq, dim, 16, seed=n + 1)
    ix = pkg.HnswIndex(dim, "vector_ip_ops", 8, 32, capacity=n)
    g = type("G", (), {})()
    g.dim, g.m, g.efc, g.n, g.upper_rows, g.entry = dim, 8, 32, n, 0, 0
    g.vecs, g.level, g.nbr0 = x, np.zeros(n, np.uint8), np.full((n, 16), -1, np.int32)
    g.uoff, g.nbru, g.ntids, g.tids = np.full(n, -1, np.int32), np.full((1, 8), -1, np.int32), np.ones(n, np.uint8), np.zeros((n, 10), np.int64)
    ix.load_graph(g)
    elem, dist, scores, st = ix.bruteforce(q, 10, debug_scores=True, stats=True)
    want = bf16_round(q).astype(np.float64) @ bf16_round(x).astype(np.float64).T
    scale = np.abs(bf16_round(q)).astype(np.float64) @ np.abs(bf16_round(x)).astype(np.float64).T
    assert np.max(np.abs(scores - want) / (scale + 1e-6)) < 2e-6     # fp32 accumulation only
    assert st["certified"] + st["rescanned"] == nq
    ix.close()


@pytest.mark.parametrize("metric,opclass,dtype", [(0, "vector_l2_ops", 0), (1, "vector_ip_ops", 0), (2, "vector_cosine_ops", 0),
                                                   (1, "halfvec_ip_ops", 1), (0, "halfvec_l2_ops", 1)])
def test_exact_topk_matches_oracle(oracle, pkg, metric, opclass, dtype):
    n, dim, nq, k = 6000, 96, 400, 10
    dt = np.float16 if dtype else np.float32
    x = clustered(n, dim, 32, seed=3, dtype=dt)
    q = clustered(nq, dim, 32, seed=4, dtype=dt)
    orc = oracle.Index(dim, 8, 32, metric, dtype)
    orc.build(x[:50])                                   # any small graph: only the vectors matter below
    g = orc.export()
    xs = x if metric != 2 else np.stack([oracle.normalize(r, dtype)[0] for r in x])
    g.n, g.vecs = n, xs
    g.level, g.nbr0 = np.zeros(n, np.uint8), np.full((n, 16), -1, np.int32)
    g.uoff, g.nbru, g.upper_rows, g.entry = np.full(n, -1, np.int32), np.full((1, 8), -1, np.int32), 0, 0
    g.ntids, g.tids = np.ones(n, np.uint8), np.zeros((n, 10), np.int64)
    ix = pkg.HnswIndex(dim, opclass, 8, 32, capacity=n)
    ix.load_graph(g)
    full = oracle.Index.from_graph(g)
    gt, gd = full.bruteforce(q, k, threads=8)
    elem, dist, st = ix.bruteforce(q, k, stats=True)
    assert st["certified"] + st["rescanned"] == nq
    for i in range(nq):
        if list(elem[i]) == list(gt[i]):
            continue
        # any difference must be an fp32-vs-double near tie
        for j in range(k):
            assert abs(dist[i, j] - gd[i, j]) <= 1e-5 * max(abs(gd[i, j]), 1e-3), (i, j, elem[i], gt[i])
    same = np.mean([list(elem[i]) == list(gt[i]) for i in range(nq)])
    assert same > 0.98
    ix.close()


def test_uncertifiable_queries_are_rescanned(oracle, pkg):
    """nearly uniform data: bf16 cannot separate the candidates, the certificate fails and the fp32
    exhaustive path takes over; results stay exact."""
    rng = np.random.default_rng(0)
    n, dim, nq, k = 3000, 64, 20, 10
    x = (1.0 + 1e-3 * rng.standard_normal((n, dim))).astype(np.float32)
    q = (1.0 + 1e-3 * rng.standard_normal((nq, dim))).astype(np.float32)
    g = type("G", (), {})()
    g.dim, g.m, g.efc, g.metric, g.dtype, g.n, g.upper_rows, g.entry = dim, 8, 32, 1, 0, n, 0, 0
    g.vecs, g.level, g.nbr0 = x, np.zeros(n, np.uint8), np.full((n, 16), -1, np.int32)
    g.uoff, g.nbru, g.ntids, g.tids = np.full(n, -1, np.int32), np.full((1, 8), -1, np.int32), np.ones(n, np.uint8), np.zeros((n, 10), np.int64)
    ix = pkg.HnswIndex(dim, "vector_ip_ops", 8, 32, capacity=n)
    ix.load_graph(g)
    elem, dist, st = ix.bruteforce(q, k, stats=True)
    assert st["rescanned"] > 0
    full = oracle.Index.from_graph(g)
    for i in range(nq):
        want = sorted((oracle.distance(q[i], x[e], oracle.IP), e) for e in range(n))[:k]
        assert [e for _, e in want] == list(elem[i])
    ix.close()
