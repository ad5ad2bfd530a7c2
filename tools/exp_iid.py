"""the iid-Gaussian variant of configs[1], reported once to show how data-dependent the headline is.
1M x 768 iid N(0,1) rows, cosine, m=16, ef_construction=64; recall@10 against the exact scan and scan throughput
per ef_search.  usage: python tools/exp_iid.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgvector_hnsw_partitioning_b200 as pkg
from bench import recall_at

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dim, nq = 768, 10000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(20260199)
x = torch.randn((n, dim), generator=g, device=dev)
q = torch.randn((nq * 4, dim), generator=g, device=dev).view(4, nq, dim)
ix = pkg.HnswIndex(dim, "vector_cosine_ops", 16, 64, capacity=n, seed=1)
t0 = time.time(); ix.build(x.cpu().numpy()); bt = time.time() - t0
c = ix.counters(reset=True)
print("build %d x %d iid: %.2f s = %.0f vectors/s (n_dist %.0f, n_pair %.0f per insert)" % (n, dim, bt, n / bt, c["n_dist"] / n, c["n_pair"] / n), flush=True)
del x
qe = q[0][:1000].contiguous()
gt, _, st = ix.bruteforce(qe.cpu().numpy(), 10, stats=True)
print("exact scan: certified %d rescanned %d" % (st["certified"], st["rescanned"]), flush=True)
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
main = torch.cuda.current_stream(dev)
for ef in (100, 200, 400, 800):
    e = torch.empty((nq, ef), dtype=torch.int32, device=dev); d = torch.empty((nq, ef), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    ix.search_dev(qe.data_ptr(), 1000, ef, e.data_ptr(), d.data_ptr(), cnt.data_ptr(), main.cuda_stream)
    torch.cuda.synchronize()
    rec = recall_at(e[:1000, :10].cpu().numpy(), gt)
    outs = [(torch.empty((nq, ef), dtype=torch.int32, device=dev), torch.empty((nq, ef), dtype=torch.float32, device=dev),
             torch.empty((nq,), dtype=torch.int32, device=dev)) for _ in range(3)]
    ix.counters(reset=True)
    for st_ in streams: st_.wait_stream(main)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st_ in streams: st_.wait_stream(main)
    for s_ in range(6):
        st_, (a, b, c_) = streams[s_ % 3], outs[s_ % 3]
        ix.search_dev(q[1 + s_ % 3].data_ptr(), nq, ef, a.data_ptr(), b.data_ptr(), c_.data_ptr(), st_.cuda_stream)
    for st_ in streams: main.wait_stream(st_)
    e1.record(main); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    cc = ix.counters(reset=True)
    gb = (cc["n_dist"] * dim * 4 + cc["n_hop0"] * 128) / 6 / 1e9
    print("ef_search=%d recall@10=%.4f  %.2f ms/step %.0f queries/s  %.0f GB/s algorithmic  n_dist/query %.0f  slow-path queries %d" %
          (ef, rec, ms, nq / ms * 1e3, gb / ms * 1e3, cc["n_dist"] / (6 * nq), cc["n_slow"]), flush=True)
