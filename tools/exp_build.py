"""Build-path experiment: time hb_build at a given shape and report recall of the result.
usage: python tools/exp_build.py n dim opclass [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgvector_hnsw_partitioning_b200 as pkg
from bench import gen_set, recall_at

n = int(sys.argv[1]); dim = int(sys.argv[2]); opc = sys.argv[3]
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
half = opc.startswith("halfvec")
dev = torch.device("cuda", 0)
x = gen_set(n, dim, 20260104, dev)
if "ip" in opc:     # C4: not normalised, norms ~ lognormal(0.1)
    x = x * torch.exp(0.1 * torch.randn((n, 1), device=dev, generator=torch.Generator(device=dev).manual_seed(5)))
xh = (x.half() if half else x).cpu().numpy()
ix = pkg.HnswIndex(dim, opc, 16, 64, capacity=n, seed=1)
if batch:
    ix.set_option("build_batch", batch)
for kv in os.environ.get("HB_OPTS", "").split(","):
    if "=" in kv:
        ix.set_option(kv.split("=")[0], int(kv.split("=")[1]))
if os.environ.get("LINK"):
    ix.set_option("link_kernel", int(os.environ["LINK"]))
t0 = time.time(); ix.build(xh); dt = time.time() - t0
c = ix.counters(reset=True)
row = dim * (2 if half else 4)
print("build %d x %d %s: %.2f s = %.0f vectors/s; n_dist %.0f/insert n_pair %.0f/insert; algorithmic %.0f GB/s" %
      (n, dim, opc, dt, n / dt, c["n_dist"] / n, c["n_pair"] / n, (c["n_dist"] + c["n_pair"]) * row / dt / 1e9), flush=True)
if os.environ.get("HB_BUILD_ONLY"):
    sys.exit(0)
q = gen_set(1000, dim, 20260104 + 1000, dev)
qh = (q.half() if half else q).cpu().numpy()
gt, _, st = ix.bruteforce(qh, 10, stats=True)
for ef in [int(v) for v in os.environ.get("EFS", "40,100").split(",")]:
    e, d, _ = ix.search_elements(qh, ef)
    print("  ef_search=%d recall@10=%.4f (exact scan: certified %d rescanned %d)" % (ef, recall_at(e[:, :10], gt), st["certified"], st["rescanned"]), flush=True)
