"""Random-row gather ceiling: the batched distance kernel (opclass FUNCTION 1) over uniformly random
candidate lists -- no graph dependency chain, so this is what the memory system gives for random
3 KB row reads with this access pattern.  usage: python tools/exp_gather.py [n] [dim] [opclass]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgvector_hnsw_partitioning_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
opc = sys.argv[3] if len(sys.argv) > 3 else "vector_ip_ops"
half = opc.startswith("halfvec")
dev = torch.device("cuda", 0)
ix = pkg.HnswIndex(dim, opc, 16, 64, capacity=n)
g = type("G", (), {})()
x = torch.randn((n, dim), device=dev, dtype=torch.float16 if half else torch.float32)
g.dim, g.m, g.efc, g.n, g.upper_rows, g.entry = dim, 16, 64, n, 0, 0
g.vecs = x.cpu().numpy(); g.level = np.zeros(n, np.uint8); g.nbr0 = np.full((n, 32), -1, np.int32)
g.uoff = np.full(n, -1, np.int32); g.nbru = np.full((1, 16), -1, np.int32); g.ntids = np.ones(n, np.uint8)
g.tids = np.zeros((n, 10), np.int64)
ix.load_graph(g)
nq = 10000
q = torch.randn((nq, dim), device=dev, dtype=x.dtype)
stream = torch.cuda.current_stream().cuda_stream
for nc in (32, 128, 512):
    cand = torch.randint(0, n, (nq, nc), device=dev, dtype=torch.int32)
    out = torch.empty((nq, nc), device=dev, dtype=torch.float32)
    for _ in range(2):
        ix.distance_dev(q.data_ptr(), nq, cand.data_ptr(), nc, out.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ix.distance_dev(q.data_ptr(), nq, cand.data_ptr(), nc, out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    row = dim * (2 if half else 4)
    print("nc=%d: %.3f ms  %.0f GB/s (rows %d B, random over %d rows = %.1f GB)" % (nc, ms, nq * nc * row / ms / 1e6, row, n, n * row / 1e9), flush=True)
