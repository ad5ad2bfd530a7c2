"""hnsw.iterative_scan (pgvector 0.8 GetScanItems + ResumeScanItems) on the GPU against the oracle's
restatement (oracle/hnsw_oracle.c orc_iter_*): every batch must hold the same element ids and
distances, the `tuples` counter must agree, and the amgettuple stream must follow."""
import numpy as np
import pytest

from conftest import clustered, sift_like

pytestmark = pytest.mark.gpu

OPC = {(0, 0): "vector_l2_ops", (1, 0): "vector_ip_ops", (2, 0): "vector_cosine_ops", (1, 1): "halfvec_ip_ops"}


def make(oracle, pkg, metric, dtype, dim, n, m=8, efc=32, gen="clustered", seed=3):
    dt = np.float16 if dtype else np.float32
    x = sift_like(n, dim, seed=seed) if gen == "sift" else clustered(n, dim, 16, seed=seed, dtype=dt)
    q = sift_like(12, dim, seed=seed + 1) if gen == "sift" else clustered(12, dim, 16, seed=seed + 1, dtype=dt)
    orc = oracle.Index(dim, m, efc, metric, dtype, oracle.CANON, seed=4)
    orc.build(x)
    ix = pkg.HnswIndex(dim, OPC[(metric, dtype)], m, efc, capacity=n, seed=4)
    ix.load_graph(orc.export())
    return x, q, orc, ix


@pytest.mark.parametrize("metric,dtype,dim,gen", [(0, 0, 24, "clustered"), (2, 0, 40, "clustered"), (1, 1, 64, "clustered"), (0, 0, 16, "sift")])
def test_batches_equal_oracle(oracle, pkg, metric, dtype, dim, gen):
    n, ef = 3000, 20
    x, q, orc, ix = make(oracle, pkg, metric, dtype, dim, n, gen=gen)
    for max_tuples in (10 ** 9, 600):
        want = [orc.iterate(q[i], ef, max_scan_tuples=max_tuples) for i in range(len(q))]
        it = ix.iterate(q, ef, max_tuples)
        nb = 0
        while True:
            r = it.next()
            if r is None:
                break
            e, d, c = r
            for i in range(len(q)):
                batches = want[i][0]
                if nb < len(batches):
                    we, wd = batches[nb]
                    assert c[i] == len(we), (i, nb, c[i], len(we))
                    assert (e[i, :c[i]] == we).all(), (i, nb)
                    assert (d[i, :c[i]].view(np.uint32) == wd.view(np.uint32)).all(), (i, nb)
                else:
                    assert c[i] == 0
            nb += 1
            assert nb < 10000
        assert nb == max(len(w[0]) for w in want)
        assert (it.tuples() == np.array([w[1] for w in want])).all()
        it.close()
    ix.close()


def test_gettuple_streams_past_ef_search(oracle, pkg):
    """amgettuple with hnsw.iterative_scan = relaxed_order returns every reachable tuple once;
    strict_order drops the ones that would go backwards; off stops after ef_search elements."""
    n, ef = 1500, 10
    x, q, orc, ix = make(oracle, pkg, 0, 0, 20, n)
    batches, _, _ = orc.iterate(q[0], ef, max_scan_tuples=10 ** 9)
    want_e = np.concatenate([b[0] for b in batches])
    want_d = np.concatenate([b[1] for b in batches])
    g = orc.export()

    def stream(mode):
        sc = ix.beginscan()
        sc.set_iterative(mode, 10 ** 9)
        sc.rescan(q[0], ef)
        out = []
        while True:
            t = sc.gettuple()
            if t is None:
                break
            out.append(t)
        sc.endscan()
        return out

    off = stream("off")
    assert [t for t, _ in off] == [int(g.tids[e, 0]) for e in want_e[:ef]]
    rel = stream("relaxed_order")
    assert len(rel) == len(want_e) > 10 * ef
    assert [t for t, _ in rel] == [int(g.tids[e, 0]) for e in want_e]
    assert len(set(t for t, _ in rel)) == len(rel)
    strict = stream("strict_order")
    keep, prev = [], None
    for e, d in zip(want_e, want_d):
        if prev is not None and d < prev:
            continue
        prev = d
        keep.append(int(g.tids[e, 0]))
    assert [t for t, _ in strict] == keep
    ds = [d for _, d in strict]
    assert all(b >= a for a, b in zip(ds, ds[1:]))
    ix.close()


def test_rescan_resets_and_empty_index(pkg):
    ix = pkg.HnswIndex(8, "vector_l2_ops", 8, 32, capacity=16)
    sc = ix.beginscan()
    sc.set_iterative("relaxed_order", 100)
    sc.rescan(np.zeros(8, np.float32), 5)
    assert sc.gettuple() is None
    x = clustered(16, 8, 2, seed=1)
    ix.build(x)
    for _ in range(2):
        sc.rescan(x[3], 5)
        got = []
        while True:
            t = sc.gettuple()
            if t is None:
                break
            got.append(t[0])
        assert sorted(got) == list(range(16)) and got[0] == 3
    sc.endscan()
    ix.close()


def test_filtered_search_reaches_k(oracle, pkg):
    """WHERE + ORDER BY + LIMIT: with a 5 % filter a plain ef_search=20 scan finds ~1 qualifying row; the
    resumable scan keeps going until 10 are found, and they are the nearest qualifying rows the scan order yields."""
    n, ef, k = 4000, 20, 10
    x, q, orc, ix = make(oracle, pkg, 0, 0, 24, n)
    rng = np.random.default_rng(0)
    allowed = rng.random(n) < 0.05
    t, d, c = ix.search_filtered(q, k, ef, allowed, max_scan_tuples=10 ** 9)
    assert (c == k).all()
    assert allowed[t].all()
    for i in range(len(q)):
        batches, _, _ = orc.iterate(q[i], ef, max_scan_tuples=10 ** 9)
        stream = [int(e) for b in batches for e in b[0] if allowed[int(e)]]      # heap TID = element id here
        assert list(t[i]) == stream[:k]
    # the plain scan cannot satisfy the LIMIT
    pt, pd, pc = ix.search(q, ef, ef)
    assert np.mean([allowed[pt[i, :pc[i]]].sum() for i in range(len(q))]) < 3
    # max_scan_tuples bounds the work: fewer than k may come back
    t2, d2, c2 = ix.search_filtered(q, k, ef, np.zeros(n, bool), max_scan_tuples=500)
    assert (c2 == 0).all()
    ix.close()
