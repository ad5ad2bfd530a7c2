"""Host-side logic of the partitioned path with world_size 2 on CPU (gloo): routing, ownership,
the all-gather exchange and the merge contract.  The kernels themselves are covered by -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pgvector_hnsw_partitioning_b200 as pkg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, n, nq, k = 8, 4000, 37, 10
    tids = np.arange(n, dtype=np.int64) * 7 + 3
    rows = pkg.split_rows(tids, P, rank, world)
    owned = pkg.owned_partitions(P, rank, world)
    assert sorted(rows) == owned == [p for p in range(P) if p % world == rank]
    # every rank computes the same routing; the union over ranks covers each row exactly once
    mine = np.zeros(n, np.int32)
    for p, idx in rows.items():
        assert (pkg.partition_route(tids[idx], P) == p).all()
        mine[idx] += 1
    cover = torch.tensor(mine)
    dist.all_reduce(cover)
    assert (cover.numpy() == 1).all()
    # per-rank "search": exact top-k of a synthetic score over the rows this rank owns
    rng = np.random.default_rng(5)                     # same on both ranks = broadcast queries
    score = rng.random((nq, n)).astype(np.float32)
    local_idx = np.sort(np.concatenate([rows[p] for p in owned]))
    order = np.argsort(score[:, local_idx], axis=1, kind="stable")[:, :k]
    lt = torch.tensor(tids[local_idx][order])
    ld = torch.tensor(np.take_along_axis(score[:, local_idx], order, axis=1))
    all_t, all_d = pkg.exchange_topk(lt, ld, world)
    assert all_t.shape == (world, nq, k) and all_d.shape == (world, nq, k)
    assert (all_t[rank] == lt).all()
    # merge contract (hb_merge_topk_dev does this on the GPU): k best of the union, ties by rank
    merged = []
    for i in range(nq):
        c = sorted((float(all_d[r, i, j]), r, int(all_t[r, i, j])) for r in range(world) for j in range(k))[:k]
        merged.append([t for _, _, t in c])
    want = tids[np.argsort(score, axis=1, kind="stable")[:, :k]]
    assert (np.array(merged) == want).all()
    out.put((rank, int(len(local_idx))))
    dist.destroy_process_group()


def test_partitioned_host_logic_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = dict(out.get() for _ in range(2))
    assert got[0] + got[1] == 4000


def test_owned_partitions_cover():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pgvector_hnsw_partitioning_b200 as pkg
    for world in (1, 2, 4, 8):
        alln = sorted(p for r in range(world) for p in pkg.owned_partitions(8, r, world))
        assert alln == list(range(8))
