"""north_star's build criterion at size: recall@10 of the oracle-built graph (sequential pgvector insert loop, natural
summation order) vs GPU-built graphs under different batch fractions, same rows, same queries, exact ground truth.
NOTE: 1000 queries scatter by +-0.6 pt in the paired difference; tools/exp_build_recall2.py is the 10 000-query form.
usage: python tools/exp_build_recall.py n dim [ef]   (the oracle build is single-threaded: ~4 min at 200000 x 768)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pgvector_hnsw_partitioning_b200 as pkg
from oracle import oracle as O
from conftest import clustered, sift_like

n, dim = int(sys.argv[1]), int(sys.argv[2])
efs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "40,100").split(",")]
l2 = dim == 128
x = sift_like(n, dim, seed=31) if l2 else clustered(n, dim, 256, seed=33)
q = sift_like(1000, dim, seed=32) if l2 else clustered(1000, dim, 256, seed=34)
metric, opclass = (O.L2, "vector_l2_ops") if l2 else (O.COSINE, "vector_cosine_ops")


def recall(ids, gt):
    return float(np.mean([len(set(ids[i]) & set(gt[i])) / gt.shape[1] for i in range(len(gt))]))


gt = None
for name, opts in (("default", {}), ("small=16", {"build_fraction_small": 16}), ("small=8", {"build_fraction_small": 8}),
                   ("all=32", {"build_fraction_small": 32, "build_fraction": 32}), ("all=64", {"build_fraction_small": 64, "build_fraction": 64})):
    ix = pkg.HnswIndex(dim, opclass, 16, 64, capacity=n, seed=1)
    for k, v in opts.items():
        ix.set_option(k, v)
    t0 = time.time(); ix.build(x); dt = time.time() - t0
    if gt is None:
        gt, _ = ix.bruteforce(q, 10)
    r = []
    for ef in efs:
        e, _, _ = ix.search_elements(q, ef)
        r.append("ef=%d %.4f" % (ef, recall(e[:, :10], gt)))
    print("GPU %-9s build %.2fs  recall@10: %s" % (name, dt, "  ".join(r)), flush=True)
    ix.close()
t0 = time.time()
orc = O.Index(dim, 16, 64, metric, 0, O.NATURAL, seed=1)
orc.build(x)
r = []
for ef in efs:
    oe, _, _, _ = orc.search_batch(q, ef, threads=8)
    r.append("ef=%d %.4f" % (ef, recall(oe[:, :10], gt)))
print("oracle (sequential, %d s)  recall@10: %s" % (time.time() - t0, "  ".join(r)), flush=True)
