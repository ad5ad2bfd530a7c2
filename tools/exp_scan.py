"""Experiment harness: build the C2 index once, then time the scan under several option settings.
usage: python tools/exp_scan.py [n] [ef] ; settings are (variant, slots, grid) triples below."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pgvector_hnsw_partitioning_b200 as pkg
from bench import gen_set, exact_topk, recall_at

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dim, nq = int(os.environ.get("DIM", "768")), int(os.environ.get("NQ", "10000"))
OPC = os.environ.get("OPC", "vector_cosine_ops")
dev = torch.device("cuda", 0)
x = gen_set(n, dim, 20260102, dev)
ix = pkg.HnswIndex(dim, OPC, 16, 64, capacity=n, seed=1)
t0 = time.time(); ix.build(x.cpu().numpy()); print("build %.1fs" % (time.time() - t0), flush=True)
qe = gen_set(1000, dim, 20260102 + 1000, dev)
del x
NS = int(os.environ.get("NSTREAM", "2"))
STEPS = int(os.environ.get("STEPS", "8"))
qs = gen_set(nq * (STEPS + 2), dim, 20260102 + 2000, dev).view(STEPS + 2, nq, dim)
streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
outs = [(torch.empty((nq, ef), dtype=torch.int32, device=dev), torch.empty((nq, ef), dtype=torch.float32, device=dev),
         torch.empty((nq,), dtype=torch.int32, device=dev)) for _ in range(NS)]
main = torch.cuda.current_stream(dev)


def run(first, count):
    for st in streams:
        st.wait_stream(main)
    for s_ in range(count):
        st, (e_, d_, c_) = streams[s_ % NS], outs[s_ % NS]
        ix.search_dev(qs[first + s_].data_ptr(), nq, ef, e_.data_ptr(), d_.data_ptr(), c_.data_ptr(), st.cuda_stream)
    for st in streams:
        main.wait_stream(st)


settings = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(0, 0, 0)]
for variant, slots, grid in settings:
    ix.set_option("variant", variant); ix.set_option("slots", slots); ix.set_option("grid", grid)
    run(0, 2)
    torch.cuda.synchronize(); ix.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    run(2, STEPS)
    e1.record(main); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    c = ix.counters(reset=True)
    gb = (c["n_dist"] * dim * 4 + c["n_hop0"] * 128 + c["n_hopu"] * 64) / STEPS / 1e9
    print("variant=%d slots=%d grid=%d streams=%d: %.3f ms/step  %.0f QPS  %.0f GB/s alg  slow=%d" %
          (variant, slots, grid, NS, ms, nq / ms * 1e3, gb / ms * 1e3, c["n_slow"]), flush=True)
