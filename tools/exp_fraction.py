import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import pgvector_hnsw_partitioning_b200 as pkg
from oracle import oracle as O
from conftest import clustered, sift_like
def recall(ids, gt): return float(np.mean([len(set(ids[i]) & set(gt[i])) / gt.shape[1] for i in range(len(gt))]))
for metric, dtype, dim, opc in ((0, 0, 128, "vector_l2_ops"), (2, 0, 96, "vector_cosine_ops"), (1, 1, 64, "halfvec_ip_ops")):
    n, nq = 20000, 500
    dt = np.float16 if dtype else np.float32
    x = sift_like(n, dim, seed=4) if metric == 0 else clustered(n, dim, 64, seed=4, dtype=dt)
    q = sift_like(nq, dim, seed=5) if metric == 0 else clustered(nq, dim, 64, seed=5, dtype=dt)
    orc = O.Index(dim, 16, 64, metric, dtype, O.CANON, seed=1); orc.build(x)
    gt, _ = orc.bruteforce(q, 10, threads=8)
    oe, _, _, _ = orc.search_batch(q, 40, threads=8)
    line = "%s oracle %.4f" % (opc, recall(oe[:, :10], gt))
    for frac in (16, 8, 4):
        ix = pkg.HnswIndex(dim, opc, 16, 64, capacity=n, seed=1); ix.set_option("build_fraction", frac)
        t0 = time.time(); ix.build(x); dt_ = time.time() - t0
        ge, _, _ = ix.search_elements(q, 40)
        line += " | 1/%d: %.4f (%.2fs)" % (frac, recall(ge[:, :10], gt), dt_)
        ix.close()
    print(line, flush=True)
