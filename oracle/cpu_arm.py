#!/usr/bin/env python
"""The CPU arm of bench.py: times the oracle (CPU restatement of pgvector's HNSW path) on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE.  bench.py runs this file as a CHILD PROCESS for its `cpu_baseline`
leg and for `--impl reference`; the child imports numpy and oracle/ only -- never the product package,
torch or CUDA -- so nothing of the GPU library is mapped into the process whose time is reported.
The graph it searches arrives as a flat image (.npy files in a directory, normally under /dev/shm)
written by the parent; who built that graph is recorded by the parent (`graph_built_by`).

usage: python oracle/cpu_arm.py <job.json>      -> prints one JSON object on stdout
job keys:
  graph_dir     directory with meta.json + vecs/level/nbr0/uoff/nbru/ntids/tids .npy (search jobs)
  queries       .npy, nq x dim
  ef            hnsw.ef_search
  steps, warmup timed / untimed steps; each step searches `per_step` queries (0 = sized for budget_s)
  budget_s      CPU seconds the whole search job may take (bounds per_step)
  threads       0 = all host cores
  parity        number of leading queries searched once more in BOTH summation orders, ids and
                distances saved to `parity_out` (.npz: canon_ids, canon_dist, nat_ids, nat_dist)
  build         optional {"rows": .npy, "metric", "dtype", "m", "efc", "n1", "parts", "n_part"}:
                oracle build rate, one thread over n1 rows, and `parts` independent indexes of n_part
                rows built concurrently (one per thread)
"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def load_graph(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    arr = {k: np.load(os.path.join(d, k + ".npy"), mmap_mode="r") for k in ("vecs", "level", "nbr0", "uoff", "nbru", "ntids", "tids")}
    return O.Graph(**meta, **arr)


def search_job(job, out):
    cores = os.cpu_count() or 1
    threads = job.get("threads") or cores
    g = load_graph(job["graph_dir"])
    orc = O.Index.from_graph(g, O.NATURAL)        # pgvector's scalar loops with its compiler flags
    q = np.load(job["queries"])
    ef, steps, warmup = int(job["ef"]), int(job["steps"]), int(job["warmup"])
    probe = q[:min(64, len(q))]
    t0 = time.perf_counter()
    orc.search_batch(probe, ef, threads=threads)
    per_q = (time.perf_counter() - t0) / len(probe)          # wall seconds per query with all threads busy
    per_step = int(job.get("per_step") or 0)
    if per_step <= 0:
        per_step = int(max(64, min(len(q), float(job.get("budget_s", 20.0)) / max(per_q, 1e-7) / max(steps + warmup, 1))))
    per_step = min(per_step, len(q))
    times = []
    for s in range(warmup + steps):
        lo = (s * per_step) % max(1, len(q) - per_step + 1)
        t0 = time.perf_counter()
        orc.search_batch(q[lo:lo + per_step], ef, threads=threads)
        times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    one = q[:min(len(q), max(32, int(2.0 / max(per_q * threads, 1e-7))))]
    t0 = time.perf_counter()
    orc.search_batch(one, ef, threads=1)
    dt1 = time.perf_counter() - t0
    out.update({"queries_per_s": per_step * len(timed) / sum(timed), "ms_per_step": 1e3 * sum(timed) / len(timed),
                "per_step": per_step, "threads": threads, "cores": cores, "single_thread_queries_per_s": len(one) / dt1,
                "mode": "natural (pgvector's -ftree-vectorize -fassociative-math loops)"})
    npar = int(job.get("parity") or 0)
    if npar > 0:
        qs = q[:npar]
        ne, nd, ncnt, _ = orc.search_batch(qs, ef, threads=threads)
        orc.set_mode(O.CANON)
        ce, cd, ccnt, cctr = orc.search_batch(qs, ef, threads=threads)
        np.savez(job["parity_out"], canon_ids=ce, canon_dist=cd, canon_cnt=ccnt, nat_ids=ne, nat_dist=nd, nat_cnt=ncnt)
        out["parity_queries"] = int(npar)
        out["canon_counters"] = cctr


def build_job(b, out):
    rows = np.load(b["rows"], mmap_mode="r")
    dim = rows.shape[1]
    metric, dtype, m, efc = int(b["metric"]), int(b["dtype"]), int(b.get("m", 16)), int(b.get("efc", 64))
    n1 = min(int(b["n1"]), rows.shape[0])
    ix = O.Index(dim, m, efc, metric, dtype, O.NATURAL, seed=1)
    t0 = time.perf_counter()
    ix.build(np.ascontiguousarray(rows[:n1]))
    dt1 = time.perf_counter() - t0
    ctr = ix.build_counters()
    parts, n_part = int(b.get("parts") or (os.cpu_count() or 1)), int(b["n_part"])
    n_part = min(n_part, rows.shape[0] // max(parts, 1))
    res = {}
    if parts > 1 and n_part > 0:
        idx = [O.Index(dim, m, efc, metric, dtype, O.NATURAL, seed=2 + p) for p in range(parts)]
        chunks = [np.ascontiguousarray(rows[p * n_part:(p + 1) * n_part]) for p in range(parts)]
        th = [threading.Thread(target=idx[p].build, args=(chunks[p],)) for p in range(parts)]     # ctypes drops the GIL
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dtp = time.perf_counter() - t0
        res = {"partitions": parts, "rows_per_partition": n_part, "vectors_per_s": parts * n_part / dtp}
    out["build"] = {"single_thread": {"rows": n1, "vectors_per_s": n1 / dt1, "n_dist_per_insert": ctr["n_dist"] / n1,
                                      "n_pair_per_insert": ctr["n_pair"] / n1},
                    "concurrent": res, "cores": os.cpu_count() or 1,
                    "note": "oracle insert loop (HnswFindElementNeighbors + HnswUpdateConnection restated), in-memory, "
                            "no WAL/buffer manager; small samples flatter the CPU (search cost grows with log n)"}


def main():
    job = json.load(open(sys.argv[1]))
    out = {"kind": "port", "what": "CPU restatement of pgvector HNSW semantics (oracle/), not pgvector"}
    if job.get("graph_dir"):
        search_job(job, out)
    if job.get("build"):
        build_job(job["build"], out)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
