"""Build-quality A/B against a sequentially built graph prepared off the GPU box: _cache/orc_clu_200000_768.npz (the file
tests/test_gpu_build.py writes under HB_ORACLE_GRAPH_CACHE=_cache) holds the oracle-built graph of clustered(200000, 768, 256, seed=33); this script regenerates the rows,
loads that graph into an index, builds GPU graphs under the option sets given, and compares recall@10 over 10 000 queries.
usage: HB_RECALL_ROWS=200000 python tools/exp_build_recall2.py [ef,ef,...] [name:opt=v,opt=v ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pgvector_hnsw_partitioning_b200 as pkg
from conftest import clustered

n, dim, nq = int(os.environ.get("HB_RECALL_ROWS", "200000")), 768, 10000
efs = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "40,100").split(",")]
x = clustered(n, dim, 256, seed=33)
q = clustered(nq, dim, 256, seed=34)


def recall_rows(ids, gt):
    return np.array([len(set(ids[i]) & set(gt[i])) / gt.shape[1] for i in range(len(gt))])


z = np.load(os.path.join(ROOT, "_cache", "orc_clu_%d_768.npz" % n))


class G:
    pass


g = G()
g.n = n
g.upper_rows, g.entry, g.entry_level = [int(v) for v in z["meta"]]
xn = np.empty_like(x)
for i0 in range(0, n, 65536):
    c = x[i0:i0 + 65536].astype(np.float64)
    xn[i0:i0 + 65536] = (c / np.sqrt((c * c).sum(1, keepdims=True))).astype(np.float32)
g.vecs, g.level, g.nbr0, g.uoff, g.nbru, g.ntids = xn, z["level"], z["nbr0"], z["uoff"], z["nbru"], z["ntids"]
if "tids" in z:
    g.tids = z["tids"]
else:
    g.tids = np.zeros((n, 10), np.int64); g.tids[:, 0] = z["tids0"]
ixo = pkg.HnswIndex(dim, "vector_cosine_ops", 16, 64, capacity=n, seed=1)
ixo.load_graph(g)
gt, _ = ixo.bruteforce(q, 10)
base = {}
for ef in efs:
    e, _, c = ixo.search_elements(q, ef)
    base[ef] = recall_rows(e[:, :10], gt)
print("oracle-built: " + "  ".join("ef=%d %.4f" % (ef, base[ef].mean()) for ef in efs) + "  mean degree %.2f" % (g.nbr0 >= 0).sum(1).mean(), flush=True)
ixo.close()
sets = [("default", {})]
for a in sys.argv[2:]:
    name, _, rest = a.partition(":")
    sets.append((name, {kv.split("=")[0]: int(kv.split("=")[1]) for kv in rest.split(",") if kv}))
for name, opts in sets:
    ix = pkg.HnswIndex(dim, "vector_cosine_ops", 16, 64, capacity=n, seed=1)
    for k, v in opts.items():
        ix.set_option(k, v)
    t0 = time.time(); ix.build(x); dt = time.time() - t0
    r = []
    for ef in efs:
        e, _, _ = ix.search_elements(q, ef)
        rr = recall_rows(e[:, :10], gt)
        d = rr - base[ef]
        r.append("ef=%d %.4f (paired diff %+.4f +- %.4f)" % (ef, rr.mean(), d.mean(), d.std() / np.sqrt(len(d))))
    deg = (ix.export_graph().nbr0 >= 0).sum(1).mean()
    print("GPU %-12s build %.2fs  %s  mean degree %.2f" % (name, dt, "  ".join(r), deg), flush=True)
    ix.close()
