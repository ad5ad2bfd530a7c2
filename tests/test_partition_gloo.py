"""Host-side logic of the partitioned path with world_size 2 on CPU (gloo): routing, ownership,
the shape of the one all-gather and the merge contract (ordered by (distance, tid), independent of how
partitions are grouped into ranks).  The kernels and the NCCL path are covered by -m gpu tests
(tests/test_gpu_partition.py, which spawns two NCCL ranks when two devices are visible)."""
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
    # the one exchange of the data path: every rank's packed block [tids | dist], gathered (hb_part does this
    # with ncclAllGather on the device; here gloo carries the same bytes)
    block = torch.cat([lt.contiguous().view(torch.uint8).flatten(), ld.contiguous().view(torch.uint8).flatten()])
    assert block.numel() == nq * k * 12
    gathered = torch.empty(world * block.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, block)
    blocks = gathered.view(world, -1)
    all_t = [blocks[r, :nq * k * 8].view(torch.int64).view(nq, k).numpy() for r in range(world)]
    all_d = [blocks[r, nq * k * 8:].view(torch.float32).view(nq, k).numpy() for r in range(world)]
    assert (all_t[rank] == lt.numpy()).all()
    # merge contract (part_merge_kernel): greedy head merge ordered by (distance, tid)
    merged = np.array([pkg.merge_rule([(all_t[r][i], all_d[r][i]) for r in range(world)], k)[0] for i in range(nq)])
    want = tids[np.argsort(score, axis=1, kind="stable")[:, :k]]
    assert (merged == want).all()
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


def test_merge_rule_is_grouping_independent():
    """Merging per rank and then across ranks equals one flat merge, exact ties included (the key
    (distance, tid) is a total order): the answer does not depend on the world size."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pgvector_hnsw_partitioning_b200 as pkg
    rng = np.random.default_rng(9)
    P, k = 8, 10
    for trial in range(50):
        lists = []
        tid = rng.permutation(1000)
        for p in range(P):
            c = int(rng.integers(0, k + 1))
            d = np.sort(rng.integers(0, 6, c)).astype(np.float32)          # few distinct values: many ties
            t = np.full(k, -1, np.int64)
            t[:c] = tid[p * k:p * k + c]                                    # within-list tie order is NOT by tid
            dd = np.full(k, np.inf, np.float32)
            dd[:c] = d
            lists.append((t, dd))
        flat = pkg.merge_rule(lists, k)
        for world in (2, 4, 8):
            per_rank = []
            for r in range(world):
                t, d = pkg.merge_rule([lists[p] for p in range(P) if p % world == r], k)
                per_rank.append((np.array(t), np.array(d, np.float32)))
            assert pkg.merge_rule(per_rank, k) == flat


def test_owned_partitions_cover():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pgvector_hnsw_partitioning_b200 as pkg
    for world in (1, 2, 4, 8):
        alln = sorted(p for r in range(world) for p in pkg.owned_partitions(8, r, world))
        assert alln == list(range(8))
