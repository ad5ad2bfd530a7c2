"""The oracle's HNSW semantics: exhaustive equivalence on tiny inputs, recall against brute force,
pgvector's hnsw regression cases [RECALL], duplicates, determinism, image round trip."""
import numpy as np
import pytest

from conftest import clustered, sift_like


def test_hnsw_vector_regress_recall(oracle):
    """upstream test/sql/hnsw_vector.sql as remembered [RECALL]: rows [0,0,0],[1,2,3],[1,1,1] indexed,
    then [1,2,4] inserted, ORDER BY val <op> '[3,3,3]'."""
    O = oracle
    rows = np.array([[0, 0, 0], [1, 2, 3], [1, 1, 1], [1, 2, 4]], np.float32)
    q = np.array([3, 3, 3], np.float32)
    want = {O.L2: [1, 3, 2, 0], O.IP: [3, 1, 2, 0], O.COSINE: [2, 1, 3]}   # zero vector is not indexed for cosine
    for metric, order in want.items():
        ix = O.Index(3, 16, 64, metric)
        for i, r in enumerate(rows):
            ix.insert(r, i)
        t, d = ix.search_tids(q, 40, 10)
        assert list(t) == order, (metric, t)


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_exhaustive_when_n_le_ef(oracle, metric):
    """With N <= ef every element is reached on a connected graph: HNSW equals the exact scan."""
    O = oracle
    x = clustered(60, 16, 4, seed=metric)
    q = clustered(20, 16, 4, seed=100 + metric)
    ix = O.Index(16, 8, 64, metric)
    ix.build(x)
    gt, gd = ix.bruteforce(q, 10)
    for i in range(len(q)):
        e, d, _ = ix.search_elements(q[i], 64)
        assert len(e) == ix.n
        assert list(e[:10]) == list(gt[i])
        assert np.all(np.diff(d) >= 0)


@pytest.mark.parametrize("metric,dtype", [(0, 0), (2, 0), (1, 1)])
def test_recall_vs_bruteforce(oracle, metric, dtype):
    O = oracle
    x = clustered(4000, 32, 32, seed=1, dtype=np.float16 if dtype else np.float32)
    q = clustered(100, 32, 32, seed=2, dtype=np.float16 if dtype else np.float32)
    ix = O.Index(32, 16, 64, metric, dtype)
    ix.build(x)
    e, d, cnt, ctr = ix.search_batch(q, 40, threads=4)
    gt, _ = ix.bruteforce(q, 10, threads=4)
    rec = np.mean([len(set(e[i, :10]) & set(gt[i])) / 10 for i in range(len(q))])
    assert rec >= 0.95, rec                       # upstream's TAP recall tests use 0.95-0.99 thresholds
    assert ctr["n_dist"] > 0 and ctr["n_hop0"] >= len(q)


def test_deterministic_and_image_roundtrip(oracle):
    O = oracle
    x = sift_like(1500, 32, seed=3)
    q = sift_like(50, 32, seed=4)
    a = O.Index(32, 8, 32, O.L2, seed=5)
    b = O.Index(32, 8, 32, O.L2, seed=5)
    a.build(x)
    b.build(x)
    ga, gb = a.export(), b.export()
    assert (ga.nbr0 == gb.nbr0).all() and (ga.nbru == gb.nbru).all() and ga.entry == gb.entry
    c = O.Index.from_graph(ga)
    ea, da, _, _ = a.search_batch(q, 20)
    ec, dc, _, _ = c.search_batch(q, 20)
    assert (ea == ec).all() and (da == dc).all()
    # structure invariants: degrees bounded, no self loops, upper rows only for level >= 1
    assert ((ga.nbr0 >= -1) & (ga.nbr0 < ga.n)).all()
    assert not (ga.nbr0 == np.arange(ga.n)[:, None]).any()
    assert ((ga.uoff >= 0) == (ga.level > 0)).all()
    assert ga.level[ga.entry] == ga.entry_level == ga.level.max()


def test_duplicates_share_an_element(oracle):
    """FindDuplicateInMemory: identical vectors add heap TIDs to one element (up to HNSW_HEAPTIDS)."""
    O = oracle
    x = clustered(300, 8, 4, seed=6)
    ix = O.Index(8, 8, 32, O.L2)
    ix.build(x)
    n0 = ix.n
    e = ix.insert(x[17], 1000)
    assert e == 17 and ix.n == n0
    t, d = ix.search_tids(x[17], 40, 3)
    assert set(t[:2]) == {17, 1000} and d[0] == 0.0 and d[1] == 0.0
    assert t[0] == 1000          # heaptids are emitted last-added first
    for j in range(O.HEAPTIDS - 2):
        assert ix.insert(x[17], 2000 + j) == 17
    # the 11th copy no longer fits and becomes its own element
    assert ix.insert(x[17], 3000) == n0 and ix.n == n0 + 1


def test_cosine_skips_zero_norm(oracle):
    O = oracle
    ix = O.Index(4, 8, 32, O.COSINE)
    assert ix.insert(np.zeros(4, np.float32), 0) == -1
    assert ix.insert(np.array([1, 0, 0, 0], np.float32), 1) == 0
    assert ix.n == 1


def test_level_distribution(oracle):
    """HnswInitElement: P(level >= l) = m^-l."""
    O = oracle
    lv = np.array([O.level_for(1, i, 16) for i in range(200000)])
    assert abs((lv >= 1).mean() - 1 / 16) < 0.003
    assert abs((lv >= 2).mean() - 1 / 256) < 0.001
    assert lv.max() <= O.lib().orc_max_level(16)


def test_ties_are_ordered_by_id(oracle):
    """Equal distances are ordered by element id (the oracle's deterministic refinement of
    pgvector's pairing-heap order)."""
    O = oracle
    x = sift_like(800, 8, seed=8)
    ix = O.Index(8, 8, 32, O.L2)
    ix.build(x)
    q = sift_like(40, 8, seed=9)
    for i in range(len(q)):
        e, d, _ = ix.search_elements(q[i], 30)
        for j in range(len(e) - 1):
            assert d[j] < d[j + 1] or (d[j] == d[j + 1] and e[j] < e[j + 1])


def test_search_layer_upper(oracle):
    O = oracle
    x = clustered(3000, 16, 16, seed=10)
    ix = O.Index(16, 8, 32, O.L2, seed=3)
    ix.build(x)
    ent, lvl = ix.entry
    assert lvl >= 1
    e, d, c = ix.search_layer(x[5], [ent], 1, lvl)
    assert len(e) == 1 and c["n_hopu"] >= 1


def test_iterative_scan_properties(oracle):
    """hnsw.iterative_scan restatement (orc_iter_*): the first batch is the plain scan, later batches
    never repeat an element, a full iteration reaches everything the graph connects, and with
    max_scan_tuples reached the leftovers come one at a time in distance order."""
    O = oracle
    x = clustered(2500, 24, 16, seed=11)
    q = clustered(5, 24, 16, seed=12)
    ix = O.Index(24, 8, 32, O.L2, O.F32, O.CANON, seed=2)
    ix.build(x)
    for qi in q:
        e0, d0, _ = ix.search_elements(qi, 20)
        full, tuples, ctr = ix.iterate(qi, 20, max_scan_tuples=10 ** 9)
        assert (full[0][0] == e0).all() and (full[0][1] == d0).all()
        alle = np.concatenate([b[0] for b in full])
        assert len(set(alle.tolist())) == len(alle) >= 2400
        assert tuples == len(alle)                       # every visited element is returned exactly once
        for e, d in full:
            assert (np.diff(d) >= 0).all()               # each batch nearest-first
        capped, t2, _ = ix.iterate(qi, 20, max_scan_tuples=300)
        assert t2 >= 300
        tail = [b for b in capped if len(b[0]) == 1]
        assert len(tail) >= 1
        td = np.array([b[1][0] for b in capped[len(capped) - len(tail):]])
        assert (np.diff(td) >= 0).all()                  # leftovers in distance order
        assert sum(len(b[0]) for b in capped) == t2


def test_l1_recall_vs_bruteforce(oracle):
    O = oracle
    x = clustered(4000, 32, 32, seed=21)
    q = clustered(100, 32, 32, seed=22)
    ix = O.Index(32, 16, 64, O.L1, O.F32, O.CANON, seed=3)
    ix.build(x)
    gt, gd = ix.bruteforce(q, 10, threads=4)
    e, d, _, _ = ix.search_batch(q, 40, threads=4)
    rec = np.mean([len(set(e[i, :10]) & set(gt[i])) / 10 for i in range(len(q))])
    assert rec > 0.95, rec
    # distances returned are l1 distances of the returned elements
    i = 0
    want = np.abs(x[e[i, 0]].astype(np.float64) - q[i].astype(np.float64)).sum()
    assert abs(d[i, 0] - want) <= 1e-5 * want


def test_vacuum_repair_invariants(oracle):
    """hnswvacuum.c restated (orc_bulk_delete + orc_vacuum_repair): after RepairGraph + MarkDeleted no live element points at a
    deleted one, deleted elements are unlinked and zeroed, a deleted entry point moves to the highest live element,
    searches return live tuples only with good recall, and a second vacuum is a no-op."""
    n, dim = 3000, 24
    x = clustered(n, dim, 16, seed=1)
    ix = oracle.Index(dim, 8, 32, oracle.L2, oracle.F32, oracle.CANON, seed=3)
    ix.build(x)
    ent, lvl = ix.entry
    dead = np.unique(np.concatenate([np.arange(0, n, 3), [ent]])).astype(np.int64)
    assert ix.bulk_delete(dead) == len(dead) and ix.bulk_delete(dead) == 0
    marked, repaired = ix.vacuum_repair()
    assert marked == len(dead) and repaired > 0
    g = ix.export()
    alive = g.ntids > 0
    assert alive.sum() == n - len(dead)
    assert ix.entry[0] != ent and alive[ix.entry[0]]
    assert g.level[ix.entry[0]] == g.level[alive].max()
    for e in np.nonzero(alive)[0]:
        nb = g.nbr0[e][g.nbr0[e] >= 0]
        assert alive[nb].all()
        if g.uoff[e] >= 0:
            for r in range(g.level[e]):
                nb = g.nbru[g.uoff[e] + r]
                assert alive[nb[nb >= 0]].all()
    assert (g.nbr0[~alive] == -1).all() and not g.vecs[~alive].any()
    q = clustered(100, dim, 16, seed=2)
    e1, d1, c1, _ = ix.search_batch(q, 40)
    d2 = ((q[:, None, :] - x[None, alive, :]) ** 2).sum(-1)
    gt = np.nonzero(alive)[0][np.argsort(d2, axis=1)[:, :10]]
    assert all(alive[e1[i, :c1[i]]].all() for i in range(100))
    assert np.mean([len(set(e1[i, :10]) & set(gt[i])) / 10 for i in range(100)]) > 0.95
    assert ix.vacuum_repair() == (0, 0)
