"""Experiment harness: build the C2 index once, then time the scan under several option settings.
usage: python tools/exp_scan.py [n] [ef] ; settings are (variant, slots, grid) triples below."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pgvector_hnsw_partitioning_b200 as pkg
from bench import gen_set, exact_topk, recall_at

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dim, nq = 768, int(os.environ.get("NQ", "10000"))
dev = torch.device("cuda", 0)
x = gen_set(n, dim, 20260102, dev)
ix = pkg.HnswIndex(dim, "vector_cosine_ops", 16, 64, capacity=n, seed=1)
t0 = time.time(); ix.build(x.cpu().numpy()); print("build %.1fs" % (time.time() - t0), flush=True)
qe = gen_set(1000, dim, 20260102 + 1000, dev)
gt = exact_topk(x, qe, 10)
del x
qs = gen_set(nq * 6, dim, 20260102 + 2000, dev).view(6, nq, dim)
stream = torch.cuda.current_stream().cuda_stream
elem = torch.empty((nq, ef), dtype=torch.int32, device=dev); dist = torch.empty((nq, ef), dtype=torch.float32, device=dev)
cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
settings = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(0, 0, 0)]
for variant, slots, grid in settings:
    ix.set_option("variant", variant); ix.set_option("slots", slots); ix.set_option("grid", grid)
    for w in range(2):
        ix.search_dev(qs[w].data_ptr(), nq, ef, elem.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
    torch.cuda.synchronize(); ix.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(2, 6):
        ix.search_dev(qs[s].data_ptr(), nq, ef, elem.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    c = ix.counters(reset=True)
    gb = (c["n_dist"] * 3072 + c["n_hop0"] * 128 + c["n_hopu"] * 64) / 4 / 1e9
    print("variant=%d slots=%d grid=%d: %.3f ms/step  %.0f QPS  %.0f GB/s alg  slow=%d  last_kernel_ms=%.3f" %
          (variant, slots, grid, ms, nq / ms * 1e3, gb / ms * 1e3, c["n_slow"], ix.last_search_ms()), flush=True)
