"""Oracle distance functions against known answers and float64 numpy.

Known answers marked [RECALL] are the literal-vector results of upstream pgvector's regression
tests (test/sql/functions.sql / vector_type.sql) as remembered; the reference mount has no tests
(/root/reference/README.md:1), so they cannot be diffed against a file here.
"""
import numpy as np
import pytest


def test_known_answers_recall(oracle):
    O = oracle
    # l2_distance('[0,0]','[3,4]') = 5  -> index uses the squared form: 25
    assert O.distance([0, 0], [3, 4], O.L2) == 25.0
    # '[1,2,3]' <-> '[3,4,5]' = sqrt(12)
    assert O.distance([1, 2, 3], [3, 4, 5], O.L2) == 12.0
    # inner_product('[1,2]','[3,4]') = 11 -> negative inner product -11
    assert O.distance([1, 2], [3, 4], O.IP) == -11.0
    # cosine_distance('[1,2]','[2,4]') = 0 ; '[1,0]','[0,2]' = 1 ; '[1,1]','[-1,-1]' = 2
    for a, b, want in (([1, 2], [2, 4], 0.0), ([1, 0], [0, 2], 1.0), ([1, 1], [-1, -1], 2.0)):
        an, ok1 = O.normalize(np.array(a, np.float32))
        bn, ok2 = O.normalize(np.array(b, np.float32))
        assert ok1 and ok2
        assert abs(1.0 + O.distance(an, bn, O.IP) - want) < 1e-6
    # zero vector has no direction: HnswCheckNorm fails
    _, ok = O.normalize(np.zeros(3, np.float32))
    assert not ok


@pytest.mark.parametrize("dim", [1, 3, 4, 5, 31, 100, 128, 129, 768, 1536, 2000])
@pytest.mark.parametrize("metric", [0, 1])
def test_canonical_and_natural_vs_float64(oracle, dim, metric):
    O = oracle
    rng = np.random.default_rng(dim * 7 + metric)
    a = rng.standard_normal(dim).astype(np.float32)
    b = rng.standard_normal(dim).astype(np.float32)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    ref = ((a64 - b64) ** 2).sum() if metric == O.L2 else -(a64 * b64).sum()
    scale = ((a64 - b64) ** 2).sum() if metric == O.L2 else np.abs(a64 * b64).sum()
    for mode in (O.CANON, O.NATURAL):
        got = O.distance(a, b, metric, O.F32, mode)
        assert abs(got - ref) <= 1e-5 * scale + 1e-30, (mode, got, ref)


@pytest.mark.parametrize("dim", [7, 8, 64, 1536])
def test_halfvec(oracle, dim):
    O = oracle
    rng = np.random.default_rng(dim)
    a = rng.standard_normal(dim).astype(np.float16)
    b = rng.standard_normal(dim).astype(np.float16)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    for metric, ref, scale in ((O.L2, ((a64 - b64) ** 2).sum(), ((a64 - b64) ** 2).sum()),
                               (O.IP, -(a64 * b64).sum(), np.abs(a64 * b64).sum())):
        for mode in (O.CANON, O.NATURAL):
            got = O.distance(a, b, metric, O.F16, mode)
            assert abs(got - ref) <= 1e-5 * scale, (metric, mode, got, ref)


def test_canonical_order_is_what_the_header_says(oracle):
    """Re-derive the canonical order in plain numpy fp32 and compare bit-for-bit."""
    O = oracle
    rng = np.random.default_rng(5)
    for dim in (3, 128, 300, 768):
        a = rng.standard_normal(dim).astype(np.float32)
        b = rng.standard_normal(dim).astype(np.float32)
        acc = np.zeros(128, np.float32)
        for e in range(dim):
            # fmaf(a, b, acc): emulate with float64 (product exact, one rounding)
            acc[e % 128] = np.float32(np.float64(a[e]) * np.float64(b[e]) + np.float64(acc[e % 128]))
        p = [np.float32(np.float32(acc[4 * l] + acc[4 * l + 1]) + np.float32(acc[4 * l + 2] + acc[4 * l + 3])) for l in range(32)]
        s = 16
        while s >= 1:
            for l in range(s):
                p[l] = np.float32(p[l] + p[l + s])
            s //= 2
        assert O.distance(a, b, O.IP, O.F32, O.CANON) == -float(p[0])


def test_symmetry_bitwise(oracle):
    O = oracle
    rng = np.random.default_rng(9)
    a = rng.standard_normal(200).astype(np.float32)
    b = rng.standard_normal(200).astype(np.float32)
    for metric in (O.L2, O.IP):
        assert O.distance(a, b, metric) == O.distance(b, a, metric)


def test_normalize_matches_float64(oracle):
    O = oracle
    rng = np.random.default_rng(2)
    for dim in (3, 768):
        a = rng.standard_normal(dim).astype(np.float32)
        out, ok = O.normalize(a)
        assert ok
        want = (a.astype(np.float64) / np.sqrt((a.astype(np.float64) ** 2).sum())).astype(np.float32)
        assert np.max(np.abs(out - want)) <= 1e-7
    h = rng.standard_normal(64).astype(np.float16)
    out, ok = O.normalize(h, O.F16)
    assert ok and out.dtype == np.float16
    assert abs(float((out.astype(np.float64) ** 2).sum()) - 1.0) < 5e-3


@pytest.mark.parametrize("dim", [1, 3, 100, 128, 769, 1536])
@pytest.mark.parametrize("dtype", [0, 1])
def test_l1_vs_float64(oracle, dim, dtype):
    """vector_l1_ops / halfvec_l1_ops FUNCTION 1 (pgvector 0.7 l1_distance): known answers [RECALL] and float64"""
    O = oracle
    assert O.distance([0, 0], [3, 4], O.L1) == 7.0 and O.distance([1, 2, 3], [3, 4, 5], O.L1) == 6.0
    rng = np.random.default_rng(dim + 31 * dtype)
    dt = np.float16 if dtype else np.float32
    a, b = rng.standard_normal(dim).astype(dt), rng.standard_normal(dim).astype(dt)
    ref = np.abs(a.astype(np.float64) - b.astype(np.float64)).sum()
    for mode in (O.CANON, O.NATURAL):
        assert abs(O.distance(a, b, O.L1, dtype, mode) - ref) <= 1e-5 * ref + 1e-30
    assert O.distance(a, a, O.L1, dtype) == 0.0
