"""The GPU build path (hnswbuild / hnswinsert) against the oracle.

With a batch size of 1 the pipeline is the sequential algorithm, so the graph must be IDENTICAL to
the oracle's (every neighbour list, in order).  With real batches elements of one batch do not see
each other (as pgvector's parallel build workers do not), so the criterion is north_star's:
recall@10 within 0.5 pt of the oracle-built graph."""
import numpy as np
import pytest

from conftest import clustered, sift_like

pytestmark = pytest.mark.gpu

OPC = {(0, 0): "vector_l2_ops", (1, 0): "vector_ip_ops", (2, 0): "vector_cosine_ops",
       (0, 1): "halfvec_l2_ops", (1, 1): "halfvec_ip_ops", (2, 1): "halfvec_cosine_ops"}


def graphs_equal(a, b):
    assert a.n == b.n and a.entry == b.entry and a.upper_rows == b.upper_rows
    assert (a.level == b.level).all()
    assert (a.uoff == b.uoff).all()
    assert (a.vecs.view(np.uint8) == b.vecs.view(np.uint8)).all()
    bad = np.nonzero((a.nbr0 != b.nbr0).any(axis=1))[0]
    assert len(bad) == 0, "layer-0 lists differ for %d elements, first %s" % (len(bad), bad[:5])
    if a.upper_rows:
        assert (a.nbru[:a.upper_rows] == b.nbru[:b.upper_rows]).all()
    assert (a.ntids == b.ntids).all()
    for e in range(a.n):
        assert (a.tids[e, :a.ntids[e]] == b.tids[e, :b.ntids[e]]).all()


@pytest.mark.parametrize("metric,dtype,dim,m", [(0, 0, 32, 8), (2, 0, 100, 16), (1, 0, 24, 8), (1, 1, 64, 8), (0, 0, 128, 16)])
def test_sequential_build_is_identical_to_oracle(oracle, pkg, metric, dtype, dim, m):
    n = 1500
    x = sift_like(n, dim, seed=1) if (metric == 0 and dim == 128) else clustered(n, dim, 16, seed=dim, dtype=np.float16 if dtype else np.float32)
    efc = max(2 * m, 32)
    orc = oracle.Index(dim, m, efc, metric, dtype, oracle.CANON, seed=5)
    orc.build(x)
    ix = pkg.HnswIndex(dim, OPC[(metric, dtype)], m, efc, capacity=n, seed=5)
    ix.set_option("build_batch", 1)
    assert ix.build(x) == n
    graphs_equal(orc.export(), ix.export_graph())
    ix.close()


def test_sequential_build_with_duplicates_and_zero_vectors(oracle, pkg):
    x = clustered(600, 16, 4, seed=2)
    x[100:140] = x[7]          # 40 copies: more than HNSW_HEAPTIDS, several elements
    x[300] = 0
    for metric in (0, 2):
        orc = oracle.Index(16, 8, 32, metric, 0, oracle.CANON, seed=9)
        orc.build(x)
        ix = pkg.HnswIndex(16, OPC[(metric, 0)], 8, 32, capacity=600, seed=9)
        ix.set_option("build_batch", 1)
        got = ix.build(x)
        assert got == (599 if metric == 2 else 600)
        graphs_equal(orc.export(), ix.export_graph())
        t, d, c = ix.search(x[7:8], 60, 100)
        wt, wd = orc.search_tids(x[7], 100, 60)
        assert list(t[0][:len(wt)]) == list(wt)
        ix.close()


def test_insert_into_loaded_graph(oracle, pkg):
    """hnswinsert after the graph came from elsewhere: cached neighbour distances are recomputed."""
    x = clustered(1200, 20, 8, seed=3)
    orc = oracle.Index(20, 8, 32, 0, 0, oracle.CANON, seed=2)
    orc.build(x[:1000])
    ix = pkg.HnswIndex(20, "vector_l2_ops", 8, 32, capacity=1200, seed=2)
    ix.load_graph(orc.export())
    ix.set_option("build_batch", 1)
    assert ix.insert(x[1000:], np.arange(1000, 1200)) == 200
    for i in range(1000, 1200):
        orc.insert(x[i], i)
    graphs_equal(orc.export(), ix.export_graph())
    ix.close()


def test_trim_then_insert_again(oracle, pkg):
    """hb_index_trim frees the build-only state; inserting afterwards rebuilds it and still matches the oracle."""
    x = clustered(900, 20, 8, seed=13)
    orc = oracle.Index(20, 8, 32, 0, 0, oracle.CANON, seed=2)
    orc.build(x)
    ix = pkg.HnswIndex(20, "vector_l2_ops", 8, 32, capacity=900, seed=2)
    ix.set_option("build_batch", 1)
    assert ix.build(x[:600]) == 600
    ix.trim()
    e, d, c = ix.search_elements(x[:5], 20)
    assert (e[:, 0] == np.arange(5)).all()
    assert ix.insert(x[600:], np.arange(600, 900)) == 300
    graphs_equal(orc.export(), ix.export_graph())
    ix.close()


def recall(ids, gt):
    return float(np.mean([len(set(ids[i]) & set(gt[i])) / gt.shape[1] for i in range(len(gt))]))


@pytest.mark.parametrize("metric,dtype,dim", [(0, 0, 128), (2, 0, 96), (1, 1, 64)])
def test_batched_build_recall_within_half_point(oracle, pkg, metric, dtype, dim):
    n, nq = 20000, 500
    dt = np.float16 if dtype else np.float32
    x = sift_like(n, dim, seed=4) if metric == 0 else clustered(n, dim, 64, seed=4, dtype=dt)
    q = sift_like(nq, dim, seed=5) if metric == 0 else clustered(nq, dim, 64, seed=5, dtype=dt)
    orc = oracle.Index(dim, 16, 64, metric, dtype, oracle.CANON, seed=1)
    orc.build(x)
    gt, _ = orc.bruteforce(q, 10, threads=8)
    oe, _, _, _ = orc.search_batch(q, 40, threads=8)
    r_oracle = recall(oe[:, :10], gt)
    ix = pkg.HnswIndex(dim, OPC[(metric, dtype)], 16, 64, capacity=n, seed=1)
    assert ix.build(x) == n
    ge, gd, _ = ix.search_elements(q, 40)
    r_gpu = recall(ge[:, :10], gt)
    assert r_gpu >= r_oracle - 0.005, (r_gpu, r_oracle)
    # the GPU-built graph is a valid HNSW graph: oracle search on it gives the same ids
    g = ix.export_graph()
    assert ((g.nbr0 >= -1) & (g.nbr0 < n)).all() and not (g.nbr0 == np.arange(n)[:, None]).any()
    orc2 = oracle.Index.from_graph(g)
    oe2, od2, _, _ = orc2.search_batch(q, 40, threads=8)
    assert (oe2 == ge).all() and (od2 == gd).all()
    c = ix.counters()
    assert c["n_pair"] > 0
    ix.close()


def _oracle_built(oracle, x, metric, tag):
    """the sequentially built graph (pgvector's insert loop restated, natural summation order, single thread).  With
    HB_ORACLE_GRAPH_CACHE=<dir> the graph of a given (tag, shape) is kept there between runs: the 200k x 768 build
    takes 4-7 minutes of one core and the generators are deterministic."""
    import os
    n, dim = x.shape
    cache = os.environ.get("HB_ORACLE_GRAPH_CACHE")
    path = os.path.join(cache, "orc_%s_%d_%d.npz" % (tag, n, dim)) if cache else None
    if path and os.path.exists(path):
        z = np.load(path)
        vecs = x if metric != oracle.COSINE else np.stack([oracle.normalize(r, 0, oracle.NATURAL)[0] for r in x])
        g = oracle.Graph(dim=dim, m=16, efc=64, metric=metric, dtype=0, n=n, upper_rows=int(z["meta"][0]), entry=int(z["meta"][1]),
                         entry_level=int(z["meta"][2]), vecs=vecs, level=z["level"], nbr0=z["nbr0"], uoff=z["uoff"],
                         nbru=z["nbru"], ntids=z["ntids"], tids=z["tids"])
        return oracle.Index.from_graph(g, mode=oracle.NATURAL)
    orc = oracle.Index(dim, 16, 64, metric, 0, oracle.NATURAL, seed=1)
    orc.build(x)
    if path:
        g = orc.export()                # the rows are not kept: they are regenerated and normalised again on load
        np.savez_compressed(path, meta=np.array([g.upper_rows, g.entry, g.entry_level], np.int64), level=g.level, nbr0=g.nbr0,
                            uoff=g.uoff, nbru=g.nbru, ntids=g.ntids, tids=g.tids)
    return orc


def _recall_pair(oracle, pkg, x, q, metric, opclass, efs=(40,), tag="g"):
    """recall@10 at each ef_search of the oracle-built graph and of the GPU-batched-built graph, on the same rows, against
    the exact answer, per query: {ef: (oracle mean, gpu mean, standard error of the paired difference)}.  10 000 queries:
    with 1000 the paired difference of two equally good graphs scatters by +-0.6 pt, more than the criterion."""
    n, dim = x.shape
    orc = _oracle_built(oracle, x, metric, tag)
    ix = pkg.HnswIndex(dim, opclass, 16, 64, capacity=n, seed=1)
    assert ix.build(x) == n
    gt, _ = ix.bruteforce(q, 10)
    per_query = lambda ids: np.array([len(set(ids[i]) & set(gt[i])) / 10 for i in range(len(gt))])
    out = {}
    for ef in efs:
        oe, _, _, _ = orc.search_batch(q, ef, threads=8)
        ge, _, _ = ix.search_elements(q, ef)
        ro, rg = per_query(oe[:, :10]), per_query(ge[:, :10])
        out[ef] = (float(ro.mean()), float(rg.mean()), float((rg - ro).std() / np.sqrt(len(gt))))
    ix.close()
    return out


def test_build_recall_at_configs0_size(oracle, pkg):
    """north_star: "a GPU-built graph's recall@10 falls within 0.5 pt of the reference's" -- at the size of configs[0]
    (100k x 128 fp32 L2, m=16, ef_construction=64, ef_search=40).  The oracle build is the slow part (~40 s)."""
    x = sift_like(100000, 128, seed=31)
    q = sift_like(10000, 128, seed=32)
    r_o, r_g, se = _recall_pair(oracle, pkg, x, q, oracle.L2, "vector_l2_ops", tag="sift")[40]
    print("100k x 128: oracle-built %.4f, GPU-built %.4f (+- %.4f)" % (r_o, r_g, se))
    assert r_g >= r_o - 0.005, (r_g, r_o, se)


@pytest.mark.parametrize("n", [30000] + ([200000] if __import__("os").environ.get("HB_SLOW_TESTS") else []))
def test_build_recall_768_cosine(oracle, pkg, n):
    """the same criterion on configs[1]-shaped rows (768-d cosine).  30k rows in the default suite; HB_SLOW_TESTS=1 adds
    200k rows (a 4-7 minute single-threaded oracle build).  Measured at 200k over 10 000 queries
    (profiles/r2_build_recall_200k.txt): paired difference GPU-built minus oracle-built -0.12 +- 0.18 pt at ef_search=40
    and +0.08 +- 0.16 pt at 100 -- the 0.8 pt an earlier 1000-query run showed was sampling scatter."""
    x = clustered(n, 768, 256, seed=33)
    q = clustered(10000, 768, 256, seed=34)
    res = _recall_pair(oracle, pkg, x, q, oracle.COSINE, "vector_cosine_ops", efs=(40, 100), tag="clu")
    for ef, (r_o, r_g, se) in res.items():
        print("%d x 768 cosine, ef_search=%d: oracle-built %.4f, GPU-built %.4f (+- %.4f)" % (n, ef, r_o, r_g, se))
    for ef in (40, 100):
        assert res[ef][1] >= res[ef][0] - 0.005, res


def test_build_is_deterministic(pkg):
    x = clustered(5000, 48, 32, seed=6)
    gs = []
    for _ in range(2):
        ix = pkg.HnswIndex(48, "vector_cosine_ops", 16, 64, capacity=5000, seed=3)
        ix.build(x)
        gs.append(ix.export_graph())
        ix.close()
    graphs_equal(gs[0], gs[1])


def test_capacity_and_argument_errors(pkg):
    ix = pkg.HnswIndex(8, "vector_l2_ops", 8, 32, capacity=10)
    ix.set_option("auto_grow", 0)
    with pytest.raises(pkg.HnswError):
        ix.build(clustered(11, 8, 2, seed=1))
    ix.close()
    for bad in (dict(m=1), dict(m=101), dict(ef_construction=3), dict(m=40, ef_construction=64)):
        with pytest.raises(pkg.HnswError):
            pkg.HnswIndex(8, "vector_l2_ops", capacity=10, **bad)


def test_large_m_uses_the_warp_link_kernel(oracle, pkg):
    """2m > 63 candidates do not fit the 64-bit selection masks: the pair cache and the pipelined kernel are
    bypassed and the warp-per-list kernel must still reproduce the oracle."""
    n, dim, m = 700, 12, 36
    x = clustered(n, dim, 8, seed=21)
    orc = oracle.Index(dim, m, 2 * m, 0, 0, oracle.CANON, seed=4)
    orc.build(x)
    ix = pkg.HnswIndex(dim, "vector_l2_ops", m, 2 * m, capacity=n, seed=4)
    ix.set_option("build_batch", 1)
    assert ix.build(x) == n
    graphs_equal(orc.export(), ix.export_graph())
    ix.close()


def test_batched_build_folds_duplicates(oracle, pkg):
    """real batches + rows that duplicate vectors indexed by earlier batches: the flag path re-runs the tail
    with folded numbering; every heap TID must stay reachable and the graph must stay a valid HNSW graph."""
    n, dim = 6000, 24
    x = clustered(n, dim, 16, seed=31)
    x[3000:3040] = x[11]            # 40 copies of an early row: 10 TIDs per element, several elements
    x[4000:4005] = x[2500]
    x[5000] = x[4999]               # duplicate of a row of the same batch: may or may not fold
    ix = pkg.HnswIndex(dim, "vector_l2_ops", 16, 64, capacity=n, seed=2)
    assert ix.build(x) == n
    g = ix.export_graph()
    assert g.n < n and int(g.ntids.sum()) == n and g.ntids.max() == 10
    all_tids = np.concatenate([g.tids[e, :g.ntids[e]] for e in range(g.n)])
    assert sorted(all_tids.tolist()) == list(range(n))
    assert ((g.nbr0 >= -1) & (g.nbr0 < g.n)).all() and not (g.nbr0 == np.arange(g.n)[:, None]).any()
    t, d, c = ix.search(x[11:12], 60, 200)
    want = set([11] + list(range(3000, 3040)))
    assert want <= set(t[0, :c[0]].tolist()) and (d[0, :41] == 0).all()
    # same rows, no duplicates: recall of the batched build with folds stays at the oracle's level
    q = clustered(300, dim, 16, seed=32)
    orc = oracle.Index.from_graph(g)
    gt, _ = orc.bruteforce(q, 10, threads=8)
    e, _, _ = ix.search_elements(q, 40)
    rec = float(np.mean([len(set(e[i, :10]) & set(gt[i])) / 10 for i in range(len(q))]))
    assert rec > 0.97, rec
    ix.close()


def test_bulk_delete_removes_heap_tids(oracle, pkg):
    """ambulkdelete first pass (RemoveHeapTids): dead TIDs are never returned again, elements left without
    TIDs keep routing (results for the other rows unchanged apart from the dead ones dropping out), a later
    insert of the same vector reuses the emptied element."""
    n, dim = 3000, 20
    x = clustered(n, dim, 16, seed=41)
    x[1000:1004] = x[5]                                  # element of row 5 holds 5 TIDs
    ix = pkg.HnswIndex(dim, "vector_l2_ops", 8, 32, capacity=n + 10, seed=3)
    assert ix.build(x) == n
    q = clustered(60, dim, 16, seed=42)
    t0, d0, c0 = ix.search(q, 30, 100)
    dead = np.array([5, 1001, 1003, 17, 18, 19, 2999, 123456789], np.int64)       # the last one is not in the index
    assert ix.bulk_delete(dead) == 7
    assert ix.bulk_delete(dead) == 0
    t1, d1, c1 = ix.search(q, 30, 100)
    deadset = set(dead.tolist())
    for i in range(len(q)):
        want = [t for t in t0[i, :c0[i]].tolist() if t not in deadset]
        got = t1[i, :c1[i]].tolist()
        assert not (set(got) & deadset)
        assert got[:len(want)] == want and len(got) >= len(want)
    t, d, c = ix.search(x[5:6], 10, 40)
    assert sorted(t[0, :2].tolist()) == [1000, 1002] and d[0, 0] == 0 and d[0, 1] == 0 and d[0, 2] > 0
    g = ix.export_graph()
    assert int(g.ntids.sum()) == n - 7
    e17 = int(np.nonzero((g.tids[:, 0] == 0) & (g.ntids == 0))[0][0])            # an emptied element is still in the graph
    assert g.ntids[e17] == 0
    # re-inserting a deleted vector folds into its emptied element (FindDuplicateInMemory)
    before = ix.n
    assert ix.insert(x[17:18], np.array([777777], np.int64)) == 1
    assert ix.n == before
    t, d, c = ix.search(x[17:18], 1, 40)
    assert t[0, 0] == 777777 and d[0, 0] == 0
    ix.close()


@pytest.mark.parametrize("dim", [1600, 2000])
def test_wide_rows_single_stage_fill_and_in_place_fill(oracle, pkg, dim):
    """6.4 kB rows leave room for ONE stage of the pipelined kernel (the pair-cache fill pre-pass runs unpipelined);
    8 kB rows for none (the memoising kernel fills the matrix in place).  Both must reproduce the oracle."""
    n = 900
    x = clustered(n, dim, 8, seed=dim)
    orc = oracle.Index(dim, 16, 64, 0, 0, oracle.CANON, seed=7)
    orc.build(x)
    ix = pkg.HnswIndex(dim, "vector_l2_ops", 16, 64, capacity=n, seed=7)
    ix.set_option("build_batch", 1)
    assert ix.build(x) == n
    graphs_equal(orc.export(), ix.export_graph())
    ix.close()
    ix = pkg.HnswIndex(dim, "vector_l2_ops", 16, 64, capacity=n, seed=7)      # real batches: must terminate and search well
    assert ix.build(x) == n
    e, d, c = ix.search_elements(x[:50], 40)
    assert (e[:, 0] == np.arange(50)).all()
    ix.close()


def test_index_grows_past_its_capacity(oracle, pkg):
    """a pgvector index has no capacity: builds and inserts beyond the reservation grow the arrays (cached
    distances and pair caches move along) and the graph stays the oracle's."""
    n, dim = 2500, 16
    x = clustered(n, dim, 8, seed=51)
    orc = oracle.Index(dim, 8, 32, 0, 0, oracle.CANON, seed=9)
    orc.build(x)
    ix = pkg.HnswIndex(dim, "vector_l2_ops", 8, 32, capacity=64, seed=9)
    ix.set_option("build_batch", 1)
    assert ix.build(x[:1000]) == 1000                # 64 -> grows
    for lo in range(1000, n, 300):                   # repeated hb_insert calls keep growing it
        hi = min(n, lo + 300)
        assert ix.insert(x[lo:hi], np.arange(lo, hi)) == hi - lo
    graphs_equal(orc.export(), ix.export_graph())
    ix.close()
    ix = pkg.HnswIndex(dim, "vector_cosine_ops", 16, 64, capacity=100, seed=9)       # real batches, upload-ahead across a growth
    assert ix.build(x) == n
    e, d, c = ix.search_elements(x[:40], 40)
    assert (e[:, 0] == np.arange(40)).all()
    ix.close()
