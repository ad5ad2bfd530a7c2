"""ambulkdelete on the GPU (hb_bulk_delete + hb_vacuum_repair) against the oracle's restatement of pgvector's
hnswvacuum.c (RemoveHeapTids, RepairGraph, MarkDeleted).  With vacuum_batch = 1 elements are repaired one after the
other and the graph must be IDENTICAL to the oracle's; with real batches the criterion is recall."""
import numpy as np
import pytest

from conftest import clustered, sift_like
from test_gpu_build import OPC, graphs_equal

pytestmark = pytest.mark.gpu


def both(oracle, pkg, x, metric, dtype, m, efc, seed=5):
    orc = oracle.Index(x.shape[1], m, efc, metric, dtype, oracle.CANON, seed=seed)
    orc.build(x)
    ix = pkg.HnswIndex(x.shape[1], OPC[(metric, dtype)], m, efc, capacity=x.shape[0], seed=seed)
    ix.load_graph(orc.export())
    return orc, ix


@pytest.mark.parametrize("metric,dtype,dim,m,frac", [(0, 0, 24, 8, 0.3), (2, 0, 64, 16, 0.1), (1, 1, 48, 8, 0.5), (0, 0, 128, 16, 0.2)])
def test_sequential_vacuum_is_identical_to_oracle(oracle, pkg, metric, dtype, dim, m, frac):
    n = 2000
    x = sift_like(n, dim, seed=3) if dim == 128 else clustered(n, dim, 16, seed=dim, dtype=np.float16 if dtype else np.float32)
    orc, ix = both(oracle, pkg, x, metric, dtype, m, max(2 * m, 32))
    rng = np.random.default_rng(dim)
    dead = rng.choice(n, int(frac * n), replace=False).astype(np.int64)
    ent, _ = orc.entry
    dead[0] = ent                                     # the entry point is among the deleted: it must move
    assert ix.bulk_delete(dead) == orc.bulk_delete(dead) == len(set(dead.tolist()))
    ix.set_option("vacuum_batch", 1)
    got, want = ix.vacuum_repair(), orc.vacuum_repair()
    assert got == want and got[0] == len(set(dead.tolist())) and got[1] > 0
    assert ix.entry == orc.entry and ix.entry[0] != ent
    go, gg = orc.export(), ix.export_graph()
    graphs_equal(go, gg)
    # nothing points at a deleted element, deleted elements have no lists and a zero vector
    alive = gg.ntids > 0
    for e in np.nonzero(alive)[0]:
        nb = gg.nbr0[e][gg.nbr0[e] >= 0]
        assert alive[nb].all()
    assert (gg.nbr0[~alive] == -1).all() and not gg.vecs[~alive].any()
    # searches agree with the oracle on the repaired graph and return live tuples only
    q = clustered(100, dim, 16, seed=dim + 1, dtype=np.float16 if dtype else np.float32) if dim != 128 else sift_like(100, dim, seed=4)
    e1, d1, c1 = ix.search_elements(q, 40)
    oe, od, oc, _ = orc.search_batch(q, 40, threads=2)
    assert (c1 == oc).all() and (e1 == oe).all() and (d1.view(np.uint32) == od.view(np.uint32)).all()
    t, _, c = ix.search(q, 10, 40)
    deadset = set(dead.tolist())
    assert not (set(t[t >= 0].tolist()) & deadset)
    # a second vacuum finds nothing to delete
    assert ix.vacuum_repair()[0] == 0 and orc.vacuum_repair()[0] == 0
    graphs_equal(orc.export(), ix.export_graph())
    # the index keeps working: inserts after a vacuum (sequential) still match the oracle
    y = clustered(50, dim, 16, seed=dim + 2, dtype=np.float16 if dtype else np.float32) if dim != 128 else sift_like(50, dim, seed=5)
    ix.set_option("build_batch", 1)
    tids = np.arange(10 ** 6, 10 ** 6 + 50, dtype=np.int64)
    assert ix.insert(y, tids) == 50
    ix.close()


def test_batched_vacuum_keeps_recall(oracle, pkg):
    n, dim = 20000, 64
    x = clustered(n, dim, 64, seed=11)
    q = clustered(300, dim, 64, seed=12)
    ix = pkg.HnswIndex(dim, "vector_l2_ops", 16, 64, capacity=n, seed=3)
    assert ix.build(x) == n
    dead = np.random.default_rng(1).choice(n, n // 4, replace=False).astype(np.int64)
    assert ix.bulk_delete(dead) == len(dead)
    marked, repaired = ix.vacuum_repair()
    assert marked == len(dead) and repaired > 0
    alive = np.ones(n, bool)
    alive[dead] = False
    g = ix.export_graph()
    for e in np.nonzero(alive)[0][::37]:
        nb = g.nbr0[e][g.nbr0[e] >= 0]
        assert alive[nb].all()
    d2 = ((q[:, None, :] - x[None, alive, :]) ** 2).sum(-1)
    gt = np.nonzero(alive)[0][np.argsort(d2, axis=1)[:, :10]]
    t, _, _ = ix.search(q, 10, 64)
    rec = np.mean([len(set(t[i]) & set(gt[i])) / 10 for i in range(len(q))])
    assert rec >= 0.95, rec
    ix.close()


def test_vacuum_of_everything_and_of_nothing(oracle, pkg):
    x = clustered(300, 16, 4, seed=1)
    orc, ix = both(oracle, pkg, x, 0, 0, 8, 32)
    assert ix.vacuum_repair()[0] == 0 == orc.vacuum_repair()[0]            # nothing deleted: only not-full lists are re-linked
    ix.set_option("vacuum_batch", 1)
    all_t = np.arange(300, dtype=np.int64)
    assert ix.bulk_delete(all_t) == 300 == orc.bulk_delete(all_t)
    assert ix.vacuum_repair()[0] == 300 == orc.vacuum_repair()[0]
    assert ix.entry == orc.entry == (-1, -1)
    t, d, c = ix.search(x[:5], 5, 40)
    assert (c == 0).all()
    ix.close()
