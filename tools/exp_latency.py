"""amgettuple latency as one backend sees it: hb_rescan + first hb_gettuple (host query in, first TID out),
and hb_search_batch at small batch sizes.  usage: python tools/exp_latency.py [n] [ef]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgvector_hnsw_partitioning_b200 as pkg
from bench import gen_set

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 90
dev = torch.device("cuda", 0)
x = gen_set(n, 768, 20260102, dev)
ix = pkg.HnswIndex(768, "vector_cosine_ops", 16, 64, capacity=n, seed=1)
ix.build(x.cpu().numpy())
q = gen_set(2048, 768, 20260102 + 1000, dev).cpu().numpy()
del x
sc = ix.beginscan()
for i in range(20):
    sc.rescan(q[i], ef); sc.gettuple()
t0 = time.perf_counter()
for i in range(200):
    sc.rescan(q[20 + i], ef); sc.gettuple()
dt = (time.perf_counter() - t0) / 200
print("hb_rescan + first hb_gettuple: %.1f us per query (ef_search=%d)" % (dt * 1e6, ef))
sc.endscan()
for nq in (1, 8, 64, 512):
    ix.search(q[:nq], 10, ef)
    t0 = time.perf_counter()
    reps = 50
    for r in range(reps):
        ix.search(q[(r * nq) % 1024:(r * nq) % 1024 + nq], 10, ef)
    dt = (time.perf_counter() - t0) / reps
    print("hb_search_batch nq=%d: %.1f us per call, %.0f queries/s" % (nq, dt * 1e6, nq / dt))
