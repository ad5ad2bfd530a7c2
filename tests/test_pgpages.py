"""pgvector on-disk index pages -> flat GPU layout.  The page layout is recalled,
not verified (no PostgreSQL here): these tests pin the reader (csrc/pgpages.cu) to a writer of the same
layout (tests/pgpages_writer.py) and check that a graph that went through pages searches identically."""
import numpy as np
import pytest

from conftest import clustered
from pgpages_writer import write_pages

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric,dtype,dim,m,scatter", [(0, 0, 24, 8, False), (2, 0, 100, 16, True), (1, 1, 64, 8, False), (0, 0, 700, 16, True)])
def test_pages_round_trip_and_search(oracle, pkg, metric, dtype, dim, m, scatter):
    opc = {(0, 0): "vector_l2_ops", (2, 0): "vector_cosine_ops", (1, 1): "halfvec_ip_ops"}[(metric, dtype)]
    n = 900
    dt = np.float16 if dtype else np.float32
    x = clustered(n, dim, 16, seed=7, dtype=dt)
    x[40:44] = x[3]                                     # duplicates: several heap TIDs on one element
    q = clustered(20, dim, 16, seed=8, dtype=dt)
    tids = (np.arange(n, dtype=np.int64) // 50 << 16) | (np.arange(n, dtype=np.int64) % 50 + 1)   # (block, offset >= 1)
    efc = max(32, 2 * m)
    orc = oracle.Index(dim, m, efc, metric, dtype, oracle.CANON, seed=6)
    orc.build(x, tids)
    g = orc.export()
    blob = write_pages(g, m, efc, dim, half=bool(dtype), scatter=scatter)
    assert len(blob) % 8192 == 0 and len(blob) // 8192 > 3
    ix = pkg.HnswIndex(dim, opc, m, efc, capacity=n if scatter else 16, seed=6)      # 16: the load grows the index
    ix.load_pgvector_pages(blob)
    h = ix.export_graph()
    assert (h.n, h.entry, h.upper_rows) == (g.n, g.entry, g.upper_rows)
    assert (h.level == g.level[:g.n]).all() and (h.nbr0 == g.nbr0[:g.n]).all()
    assert (h.nbru[:h.upper_rows] == g.nbru[:g.upper_rows]).all()
    assert (h.ntids == g.ntids[:g.n]).all()
    for e in range(g.n):
        assert (h.tids[e, :h.ntids[e]] == g.tids[e, :g.ntids[e]]).all()
    assert (h.vecs.view(np.uint8) == g.vecs[:g.n].view(np.uint8)).all()
    t, d, c = ix.search(q, 10, 40)
    for i in range(len(q)):
        wt, wd = orc.search_tids(q[i], 40, 10)
        assert list(t[i, :c[i]]) == list(wt) and (d[i, :c[i]] == wd).all()
    ix.close()


def test_deleted_elements_and_bad_pages(oracle, pkg):
    n, dim, m = 300, 16, 8
    x = clustered(n, dim, 8, seed=1)
    tids = np.arange(n, dtype=np.int64) + (1 << 16) + 1
    orc = oracle.Index(dim, m, 32, 0, 0, oracle.CANON, seed=2)
    orc.build(x, tids)
    g = orc.export()
    dead = {5, 17, 100}
    blob = write_pages(g, m, 32, dim, deleted=dead)
    ix = pkg.HnswIndex(dim, "vector_l2_ops", m, 32, capacity=n)
    ix.load_pgvector_pages(blob)
    t, d, c = ix.search(x[5:6], 20, 40)
    dead_tids = {int(tids[e]) for e in dead}
    assert not (set(t[0, :c[0]].tolist()) & dead_tids)          # a deleted element returns no tuple
    assert int(tids[5]) not in t[0]
    with pytest.raises(pkg.HnswError):
        ix.load_pgvector_pages(b"\0" * 8192 * 2)                 # wrong magic
    ix2 = pkg.HnswIndex(dim + 1, "vector_l2_ops", m, 32, capacity=n)
    with pytest.raises(pkg.HnswError):
        ix2.load_pgvector_pages(blob)                            # dimension mismatch
    ix.close(); ix2.close()
