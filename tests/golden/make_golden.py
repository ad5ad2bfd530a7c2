"""Generates tests/golden/hnsw_golden_v1.npz.

The reference mount holds no source, tests or golden vectors (/root/reference/README.md:1), so
these fixtures cannot come from the reference: they are produced by THIS repo's CPU oracle
(oracle/, canonical summation order) and pin it -- and with it the CUDA path -- against silent
drift between rounds.  Parity with real pgvector stays unpinned.

    python tests/golden/make_golden.py        # rewrites the .npz next to this file
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import clustered, sift_like          # noqa: E402
from oracle import oracle as O                     # noqa: E402

CASES = [  # name, metric, dtype, dim, m, efc, n, generator
    ("l2_f32_sift", O.L2, O.F32, 16, 8, 32, 600, "sift"),
    ("cos_f32", O.COSINE, O.F32, 24, 8, 32, 500, "clustered"),
    ("ip_f16", O.IP, O.F16, 40, 6, 24, 400, "clustered"),
]


def make():
    out = {}
    for name, metric, dtype, dim, m, efc, n, gen in CASES:
        npdt = np.float16 if dtype == O.F16 else np.float32
        x = sift_like(n, dim, seed=41) if gen == "sift" else clustered(n, dim, 12, seed=41, dtype=npdt)
        q = sift_like(16, dim, seed=42) if gen == "sift" else clustered(16, dim, 12, seed=42, dtype=npdt)
        if gen == "clustered":
            x[50:53] = x[7]                      # duplicates fold into one element (heap TIDs)
        ix = O.Index(dim, m, efc, metric, dtype, O.CANON, seed=17)
        ix.build(x)
        g = ix.export()
        ef = 20
        e, d, c, _ = ix.search_batch(q, ef, threads=1)
        out[name + "/x"] = x
        out[name + "/q"] = q
        out[name + "/params"] = np.array([metric, dtype, dim, m, efc, 17, ef], np.int64)
        out[name + "/n"] = np.array([g.n, g.entry, g.upper_rows], np.int64)
        out[name + "/level"] = g.level[:g.n]
        out[name + "/nbr0"] = g.nbr0[:g.n]
        out[name + "/uoff"] = g.uoff[:g.n]
        out[name + "/nbru"] = g.nbru[:g.upper_rows]
        out[name + "/ntids"] = g.ntids[:g.n]
        t = g.tids[:g.n].copy()
        t[np.arange(10)[None, :] >= g.ntids[:g.n, None]] = 0      # slots past ntids are not part of the fixture
        out[name + "/tids"] = t
        out[name + "/res_elem"] = e
        out[name + "/res_dist"] = d
        out[name + "/res_cnt"] = c
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hnsw_golden_v1.npz"), **out)
    make_iter(out)
    return out


def make_iter(base):
    """resumable scans (hnsw.iterative_scan) and L1 on the "cos_f32" rows: the first batches of four queries with
    and without a max_scan_tuples bound, and an L1 index's results -> hnsw_golden_iter_l1_v1.npz"""
    x, q = base["cos_f32/x"], base["cos_f32/q"]
    metric, dtype, dim, m, efc, seed, ef = (int(v) for v in base["cos_f32/params"])
    out = {}
    ix = O.Index(dim, m, efc, metric, dtype, O.CANON, seed=seed)
    ix.build(x)
    for qi in range(4):
        for tag, mt in (("all", 10 ** 9), ("cap300", 300)):
            batches, tuples, _ = ix.iterate(q[qi], 10, max_scan_tuples=mt, max_batches=12)
            out["iter/%d/%s/elem" % (qi, tag)] = np.concatenate([b[0] for b in batches])
            out["iter/%d/%s/dist" % (qi, tag)] = np.concatenate([b[1] for b in batches])
            out["iter/%d/%s/sizes" % (qi, tag)] = np.array([len(b[0]) for b in batches], np.int32)
    l1 = O.Index(dim, m, efc, O.L1, dtype, O.CANON, seed=seed)
    l1.build(x)
    g = l1.export()
    e, d, c, _ = l1.search_batch(q, ef, threads=1)
    out["l1/nbr0"] = g.nbr0[:g.n]
    out["l1/res_elem"], out["l1/res_dist"] = e, d
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hnsw_golden_iter_l1_v1.npz"), **out)


if __name__ == "__main__":
    o = make()
    print("wrote %d arrays" % len(o))
