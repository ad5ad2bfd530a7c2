"""vector_l1_ops / halfvec_l1_ops (pgvector 0.7+) through every stage of the path:
opclass FUNCTION 1 (l1_distance), the scan, the build with each reverse-link kernel, resumable scans.
Bar as for the other operator classes: ids and distances bit-identical to the oracle."""
import numpy as np
import pytest

from conftest import clustered, sift_like

pytestmark = pytest.mark.gpu


def test_known_answers(pkg, oracle):
    # pgvector regression expectations [RECALL]: l1_distance('[0,0]','[3,4]') = 7, '[1,2,3]' <+> '[3,4,5]' = 6
    assert oracle.distance([0, 0], [3, 4], oracle.L1) == 7.0
    assert oracle.distance([1, 2, 3], [3, 4, 5], oracle.L1) == 6.0
    ix = pkg.HnswIndex(2, "vector_l1_ops", 4, 8, capacity=4)
    ix.build(np.array([[3, 4], [0, 0]], np.float32))
    t, d, c = ix.search(np.array([[0, 0]], np.float32), 2, 8)
    assert list(t[0]) == [1, 0] and list(d[0]) == [0.0, 7.0]
    with pytest.raises(pkg.HnswError):
        ix.bruteforce(np.zeros((1, 2), np.float32), 1)       # the exact scan is a contraction: no l1
    ix.close()


@pytest.mark.parametrize("dim,dtype", [(3, 0), (100, 0), (128, 0), (768, 0), (1100, 0), (64, 1), (1536, 1)])
def test_distance_kernel_bit_exact(oracle, pkg, dim, dtype):
    dt = np.float16 if dtype else np.float32
    x = clustered(200, dim, 8, seed=dim, dtype=dt)
    q = clustered(5, dim, 8, seed=dim + 1, dtype=dt)
    ix = pkg.HnswIndex(dim, "halfvec_l1_ops" if dtype else "vector_l1_ops", 8, 32, capacity=200)
    g = oracle.Graph(dim=dim, m=8, efc=32, metric=oracle.L1, dtype=dtype, n=200, upper_rows=0, entry=0, vecs=x,
                     level=np.zeros(200, np.uint8), nbr0=np.full((200, 16), -1, np.int32), uoff=np.full(200, -1, np.int32),
                     nbru=np.full((1, 8), -1, np.int32), ntids=np.ones(200, np.uint8), tids=np.zeros((200, 10), np.int64))
    ix.load_graph(g)
    cand = np.random.default_rng(0).integers(0, 200, (5, 40)).astype(np.int32)
    got = ix.distance(q, cand)
    for i in range(5):
        for j in range(40):
            want = oracle.distance(q[i], x[cand[i, j]], oracle.L1, dtype, oracle.CANON)
            assert got[i, j] == np.float32(want)
            nat = oracle.distance(q[i], x[cand[i, j]], oracle.L1, dtype, oracle.NATURAL)
            assert abs(got[i, j] - nat) <= 1e-5 * abs(nat) + 1e-30
    ix.close()


@pytest.mark.parametrize("dim,dtype,gen", [(128, 0, "sift"), (96, 0, "clustered"), (64, 1, "clustered")])
def test_scan_and_resumable_scan(oracle, pkg, dim, dtype, gen):
    dt = np.float16 if dtype else np.float32
    n = 4000
    x = sift_like(n, dim, seed=1) if gen == "sift" else clustered(n, dim, 32, seed=1, dtype=dt)
    q = sift_like(50, dim, seed=2) if gen == "sift" else clustered(50, dim, 32, seed=2, dtype=dt)
    orc = oracle.Index(dim, 16, 64, oracle.L1, dtype, oracle.CANON, seed=1)
    orc.build(x)
    ix = pkg.HnswIndex(dim, "halfvec_l1_ops" if dtype else "vector_l1_ops", 16, 64, capacity=n, seed=1)
    ix.load_graph(orc.export())
    e, d, c = ix.search_elements(q, 40)
    oe, od, oc, octr = orc.search_batch(q, 40, threads=4)
    assert (c == oc).all() and (e == oe).all() and (d.view(np.uint32) == od.view(np.uint32)).all()
    ctr = ix.counters(reset=True)
    assert ctr["n_dist"] == octr["n_dist"] and ctr["n_hop0"] == octr["n_hop0"]
    it = ix.iterate(q[:6], 20, 800)
    want = [orc.iterate(q[i], 20, max_scan_tuples=800)[0] for i in range(6)]
    nb = 0
    while True:
        r = it.next()
        if r is None:
            break
        for i in range(6):
            if nb < len(want[i]):
                assert (r[0][i, :r[2][i]] == want[i][nb][0]).all()
            else:
                assert r[2][i] == 0
        nb += 1
    assert nb == max(len(w) for w in want)
    it.close()
    ix.close()


@pytest.mark.parametrize("link_kernel", [0, 1, 2, 3])
@pytest.mark.parametrize("dim,dtype", [(24, 0), (40, 1)])
def test_sequential_build_is_identical_to_oracle(oracle, pkg, dim, dtype, link_kernel):
    dt = np.float16 if dtype else np.float32
    n = 1200
    x = clustered(n, dim, 16, seed=3, dtype=dt)
    orc = oracle.Index(dim, 8, 32, oracle.L1, dtype, oracle.CANON, seed=5)
    orc.build(x)
    ix = pkg.HnswIndex(dim, "halfvec_l1_ops" if dtype else "vector_l1_ops", 8, 32, capacity=n, seed=5)
    ix.set_option("build_batch", 1)
    ix.set_option("link_kernel", link_kernel)
    assert ix.build(x) == n
    a, b = orc.export(), ix.export_graph()
    assert a.n == b.n and a.entry == b.entry
    assert (a.nbr0[:a.n] == b.nbr0[:b.n]).all()
    assert (a.nbru[:a.upper_rows] == b.nbru[:b.upper_rows]).all()
    ix.close()


def test_batched_build_recall(oracle, pkg):
    n, nq, dim = 20000, 300, 64
    x, q = clustered(n, dim, 64, seed=4), clustered(nq, dim, 64, seed=5)
    orc = oracle.Index(dim, 16, 64, oracle.L1, 0, oracle.CANON, seed=1)
    orc.build(x)
    gt, _ = orc.bruteforce(q, 10, threads=8)
    oe, _, _, _ = orc.search_batch(q, 40, threads=8)
    ix = pkg.HnswIndex(dim, "vector_l1_ops", 16, 64, capacity=n, seed=1)
    assert ix.build(x) == n
    ge, _, _ = ix.search_elements(q, 40)
    rec = lambda ids: float(np.mean([len(set(ids[i, :10]) & set(gt[i])) / 10 for i in range(nq)]))
    assert rec(ge) >= rec(oe) - 0.005, (rec(ge), rec(oe))
    ix.close()


@pytest.mark.parametrize("dim,dtype", [(128, 0), (768, 0), (512, 1)])
def test_cta_per_query_scan_l1(oracle, pkg, dim, dtype):
    """the four-warps-per-query kernels (csrc/scan_cta.cuh) under the l1 operator classes: forced for every row length
    (variant 7), each sub-form, against the oracle and against the warp-per-query kernel."""
    dt = np.float16 if dtype else np.float32
    n = 3000
    x = clustered(n, dim, 24, seed=dim + 3, dtype=dt)
    q = clustered(80, dim, 24, seed=dim + 4, dtype=dt)
    orc = oracle.Index(dim, 16, 64, oracle.L1, dtype, oracle.CANON, seed=2)
    orc.build(x)
    ix = pkg.HnswIndex(dim, "halfvec_l1_ops" if dtype else "vector_l1_ops", 16, 64, capacity=n, seed=2)
    ix.load_graph(orc.export())
    oe, od, oc, _ = orc.search_batch(q, 40, threads=4)
    for v in (7, 17, 27, 6):
        ix.set_option("variant", v)
        e, d, c = ix.search_elements(q, 40)
        assert (c == oc).all() and (e == oe).all() and (d.view(np.uint32) == od.view(np.uint32)).all(), v
    ix.close()
